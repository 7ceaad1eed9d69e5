"""Host-side recorder: the Python mirror of the reference's record path.

Recording stays on the CPU exactly as in the reference (it is a one-off); only the recorded
instruction list is replayed on the GPU.  This module mirrors, name for name:

* ``TrackedArray`` (``src/tracked.jl:70-87``) — here a handle on a *buffer id* of the tape program
  plus the record-time value;
* ``Adjoint`` wrappers whose ``capture`` materialises index-linked ``TrackedReal`` arrays
  (``src/tracked.jl:223-225,348-351``; ``src/tape.jl:24-25,67-69``) — operand class IDX;
* the instruction recorders for ``*`` (``linalg/arithmetic.jl:171-177``), array ``+``/``-``
  (``:36-41,111-123``), ``sum``/``mean`` (``linalg/reductions.jl:8-13,66-71``), ``∇broadcast``
  (``broadcast.jl:142-162``), ``map``/``broadcast`` of ``@forward`` closures
  (``elementwise.jl:230-310``), infix broadcasts (``elementwise.jl:449-517``) and range
  ``getindex`` (``tracked.jl:308-316``).

Julia spellings map to Python as ``a' -> a.T``, ``a*b -> a @ b``, ``f.(x, y) -> bcast(f, x, y)``,
``map(@forward(f), x, y) -> map_(forward(f), x, y)``, ``x[1:2:end] -> x[0::2]``.
"""
from __future__ import annotations

import numpy as np

from . import program as P
from . import expr as E


COMPACT_TRANSPOSE_THRESHOLD = 1 << 20


class InstructionTape:
    """``InstructionTape`` (``src/tape.jl:5-7``) being recorded into a ``TapeProgram``."""

    def __init__(self, dtype=np.float64):
        dtype = np.dtype(dtype)
        if dtype == np.float64:
            code = P.F64
        elif dtype == np.float32:
            code = P.F32
        else:
            raise TypeError("only Float64 and Float32 tapes are supported")
        self.dtype = dtype
        self.program = P.TapeProgram(code)
        # shapes_only: the tape is being recorded from DEVICE arrays (api.AbstractTape): no array element exists on the host, values
        # are zero-stride placeholders of the right shape and every record_* below infers the output shape instead of computing the
        # output (the program does not depend on values; the first forward sweep on the device produces them).  Anything that
        # would need a value - comparisons of TrackedReals, i.e. value-dependent control flow - raises.
        self.shapes_only = False

    def __len__(self):
        return len(self.program.instrs)

    def record(self, opcode, inputs, out_buffer, expr=None, n_expr_args=0):
        """``record!`` (``src/tape.jl:9-12``)."""
        self.program.instrs.append(P.Instr(opcode, inputs, out_buffer, expr, n_expr_args))

    def scalar_block(self) -> "P.ScalarBlock":
        """The block that the next ``ScalarInstruction`` (``src/tape.jl:30-45``) is appended to: consecutive scalar
        instructions share one tape entry, any array instruction in between starts a new one (tape order is kept)."""
        ins = self.program.instrs
        if ins and ins[-1].scalar is not None:
            return ins[-1].scalar
        token = self.program.add_buffer(P.BUF_TR, ())
        blk = P.ScalarBlock(token)
        ins.append(P.Instr(P.OP_SCALAR_BLOCK, [], token, scalar=blk))
        return blk


class ForwardOptimize:
    """``@forward`` (``src/macros.jl:40-74``): the wrapped scalar closure is differentiated in
    forward mode and recorded as ONE instruction."""

    def __init__(self, f):
        self.f = f

    def __call__(self, *a):
        # (self::ForwardOptimize)(t::TrackedReal) / (a, b) with at least one TrackedReal (macros.jl:84-131): ONE instruction
        if len(a) in (1, 2) and any(isinstance(x, TrackedReal) for x in a) and \
                all(isinstance(x, TrackedReal) or np.ndim(x) == 0 for x in a):
            # only TrackedReal arguments are dualised; plain Reals stay Reals (macros.jl:101-124)
            return record_scalar(E.trace(self.f, len(a), dual=[isinstance(x, TrackedReal) for x in a]), *a)
        return self.f(*a)


def forward(f):
    return f if isinstance(f, ForwardOptimize) else ForwardOptimize(f)


class SkipOptimize:
    """``@skip`` (``src/macros.jl:140-168``): evaluate on plain values, record nothing."""

    def __init__(self, f):
        self.f = f

    def __call__(self, *a):
        return self.f(*[value(x) for x in a])


def skip(f):
    return SkipOptimize(f)


class TrackedArray:
    __slots__ = ("value", "tape", "buffer")
    __array_ufunc__ = None      # numpy defers `ndarray @ tracked`, `ndarray + tracked` to the reflected methods

    def __init__(self, value_, tape: InstructionTape, buffer: int):
        self.value = value_
        self.tape = tape
        self.buffer = buffer

    # -- AbstractArray interface ---------------------------------------------------------
    @property
    def shape(self):
        return self.value.shape

    @property
    def ndim(self):
        return self.value.ndim

    @property
    def size(self):
        return self.value.size

    def __len__(self):
        return self.value.shape[0]

    @property
    def T(self):
        return Adjoint(self)

    def adjoint(self):
        return Adjoint(self)

    def transpose(self):
        return Adjoint(self)

    # -- recorded operations -------------------------------------------------------------
    def __matmul__(self, other):
        return record_mul(self, other)

    def __rmatmul__(self, other):
        return record_mul(other, self)

    def __add__(self, other):
        return record_plusminus(P.OP_ADD, self, other)

    def __radd__(self, other):
        return record_plusminus(P.OP_ADD, other, self)

    def __sub__(self, other):
        return record_plusminus(P.OP_SUB, self, other)

    def __rsub__(self, other):
        return record_plusminus(P.OP_SUB, other, self)

    def __neg__(self):
        tp = self.tape
        out = _track_new(lambda: -self.value, tp, shape=self.shape)
        tp.record(P.OP_NEG, [_capture(self, tp)], out.buffer)
        return out

    def __mul__(self, other):
        raise TypeError("`*` between arrays is matrix multiplication in the reference: use `@`; "
                        "for elementwise products use broadcast('*', a, b) or bcast(lambda x, y: x*y, a, b)")

    def __getitem__(self, key):
        return record_getindex(self, key)

    def __setitem__(self, *a):
        raise TypeError("TrackedArrays do not support setindex!")   # tracked.jl:390

    def reshape(self, *dims):
        """``reshape`` aliases value AND deriv (``src/tracked.jl:402-408``)."""
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        v = self.value.reshape(dims, order="F")
        root = self.buffer
        prog = self.tape.program
        while prog.buffers[root].alias_of >= 0:
            root = prog.buffers[root].alias_of
        b = prog.add_buffer(P.BUF_TA, v.shape, alias_of=root)
        return TrackedArray(v, self.tape, b)


class TrackedReal:
    """``TrackedReal`` (``src/tracked.jl:47-57``).  Three flavours, as in the reference:

    * output of an array instruction (``sum``, ``dot``...): owns the size-1 buffer ``buffer``;
    * scalar ``getindex`` of a TrackedArray (``tracked.jl:348-351``): NO instruction, value/deriv live in
      ``origin[index]`` — here ``(buffer, index)`` with ``index`` the 0-based linear element;
    * output of a ``ScalarInstruction`` (``macros.jl:84-131``): slot ``slot`` of scalar block ``block``.
    """

    __slots__ = ("value", "tape", "buffer", "index", "block", "slot")
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, value_, tape, buffer=-1, index=0, block=None, slot=-1):
        self.value, self.tape, self.buffer, self.index, self.block, self.slot = value_, tape, buffer, index, block, slot

    shape = ()
    ndim = 0
    size = 1

    # DiffRules binary functions (derivatives/scalars.jl:5-22): one ScalarInstruction each
    def __add__(self, o): return _scalar_binary(P.E_ADD, self, o)
    def __radd__(self, o): return _scalar_binary(P.E_ADD, o, self)
    def __sub__(self, o): return _scalar_binary(P.E_SUB, self, o)
    def __rsub__(self, o): return _scalar_binary(P.E_SUB, o, self)
    def __mul__(self, o): return _scalar_binary(P.E_MUL, self, o)
    def __rmul__(self, o): return _scalar_binary(P.E_MUL, o, self)
    def __truediv__(self, o): return _scalar_binary(P.E_DIV, self, o)
    def __rtruediv__(self, o): return _scalar_binary(P.E_DIV, o, self)
    def __neg__(self): return self._scalar_unary(P.E_NEG)
    def __pos__(self): return self

    def __pow__(self, o):
        if isinstance(o, (int, np.integer)) and not isinstance(o, bool):
            # x^2 lowers to literal_pow -> ^(t::TrackedReal, 2::Int) (binary ForwardOptimize with a Real exponent)
            ops = np.asarray([(P.E_POWI, 0, 0, int(o))], np.int32)
            return record_scalar((ops, np.zeros(0)), self)
        return _scalar_binary(P.E_POW, self, o)

    def __rpow__(self, o): return _scalar_binary(P.E_POW, o, self)

    def _scalar_unary(self, code):
        return record_scalar((np.asarray([(code, 0, 0, 0)], np.int32), np.zeros(0)), self)

    @staticmethod
    def _scalar_binary_code(code, a, b):
        return _scalar_binary(code, a, b)

    # SkipOptimize'd predicates (derivatives/scalars.jl:24-46): evaluated on values, nothing recorded — which is why a
    # compiled tape freezes the branch taken at record time (docs/src/api.md:39-44)
    def _cmp_value(self):
        if self.tape is not None and getattr(self.tape, "shapes_only", False):
            raise TypeError("value-dependent control flow on a tape recorded from device arrays: no values exist on the host "
                            "(record from host arrays instead)")
        return self.value

    def __lt__(self, o): return self._cmp_value() < value(o)
    def __le__(self, o): return self._cmp_value() <= value(o)
    def __gt__(self, o): return self._cmp_value() > value(o)
    def __ge__(self, o): return self._cmp_value() >= value(o)
    def __eq__(self, o): return self._cmp_value() == value(o)
    def __ne__(self, o): return self._cmp_value() != value(o)
    __hash__ = object.__hash__

    def __float__(self):
        return float(self.value)


def _root_buffer(prog, b):
    while prog.buffers[b].alias_of >= 0:
        b = prog.buffers[b].alias_of
    return b


def _slot_buffer(t: "TrackedReal", want_whole=False):
    """``(buffer, elem)`` through which readers outside ``t.block`` see the slot; exports it on first use.  ``want_whole``
    asks for a size-1 buffer of its own (operand class TR of an array instruction, or the tape output)."""
    blk, prog = t.block, t.tape.program
    hit = blk.export_of.get(t.slot)
    if hit is not None and (not want_whole or prog.buffers[hit[0]].kind == P.BUF_TR):
        return hit
    if want_whole:
        b = prog.add_buffer(P.BUF_TR, ())
        blk.export(t.slot, b, 0)
        if hit is None:
            return b, 0
        blk.export_of[t.slot] = hit       # keep the first export as the canonical one for scalar readers
        return b, 0
    if blk.spill_buffer < 0:
        blk.spill_buffer = prog.add_buffer(P.BUF_TA, (1,))
    k = blk.n_spill
    blk.n_spill += 1
    blk.export(t.slot, blk.spill_buffer, k)
    return blk.spill_buffer, k


def _scalar_operand(blk, x, tp):
    if isinstance(x, TrackedReal):
        if x.tape is not tp:
            raise ValueError("TrackedReal from a different tape")
        if x.block is blk:
            return (P.SP_POOL, x.slot)
        if x.block is not None:
            b, k = _slot_buffer(x)
            return (blk.ref(b), k)
        return (blk.ref(_root_buffer(tp.program, x.buffer)), int(x.index))
    if isinstance(x, (TrackedArray, Adjoint)) or np.ndim(x) != 0:
        raise TypeError("scalar instructions take TrackedReal / Real arguments")
    return (P.SP_LITERAL, blk.literal(float(x)))            # capture(::Real) (tape.jl:24): frozen by value


def record_scalar(expr, a, b=None):
    """One ``ScalarInstruction`` (``macros.jl:84-131``): ``expr`` is the traced scalar function of 1 or 2 arguments; value and
    partials are re-evaluated from it on every forward replay (``scalars.jl:80-123``)."""
    args = [a] if b is None else [a, b]
    tp = _tape_of(*[x for x in args if isinstance(x, TrackedReal)])
    blk = tp.scalar_block()
    ops, fc = expr
    eid = blk.expr(ops, fc, len(args))
    oa = _scalar_operand(blk, a, tp)
    ob = (P.SP_NONE, 0) if b is None else _scalar_operand(blk, b, tp)
    if tp.shapes_only:
        v = np.zeros((), tp.dtype)                 # no values on the host (device recording): the block evaluates on the device
    else:
        v = E.eval_value((np.asarray(ops, np.int32).reshape(-1, 4), fc), [np.asarray(value(x), tp.dtype) for x in args], tp.dtype)
    slot = blk.push(eid, oa, ob)
    return TrackedReal(v[()], tp, -1, 0, blk, slot)


def _scalar_binary(code, a, b):
    # ^ is three ForwardDiff methods (Dual^Dual, Dual^Real, Real^Dual): which one ran is part of the instruction (expr.pow_imm)
    imm = E.pow_imm(0, isinstance(a, TrackedReal), isinstance(b, TrackedReal)) if code == P.E_POW else 0
    return record_scalar((np.asarray([(code, 0, 1, imm)], np.int32), np.zeros(0)), a, b)


def _scalar_const(tp, c):
    """A block slot holding a literal (not an instruction of the reference: used to materialise plain Reals in ``collect``)."""
    blk = tp.scalar_block()
    eid = blk.expr(np.asarray([(P.E_CONST, 0, 0, 0)], np.int32), np.asarray([float(c)]), 0)
    slot = blk.push(eid, (P.SP_NONE, 0), (P.SP_NONE, 0))
    return TrackedReal(np.asarray(c, tp.dtype)[()], tp, -1, 0, blk, slot)


def _scalar_identity(t: "TrackedReal"):
    return record_scalar((np.asarray([(P.E_ARG, 0, 0, 0)], np.int32), np.zeros(0)), t)


def convert(t):
    """``convert(::Type{T1<:TrackedReal}, t::TrackedReal)`` (``src/tracked.jl:237-257``): a new TrackedReal of another value /
    deriv type that keeps the line of references back to the origin's deriv.  Forward: ``value!(output, value(input))``; reverse:
    ``increment_deriv!(input, deriv(output)); unseed!(output)`` - i.e. the identity scalar instruction (partial 1), and that is how
    it travels: a slot of the current scalar block with the one-op program ``E_ARG``.  (All values of one tape share the tape's
    element type here, so the conversion of the value itself is a no-op.)  A plain Real converts to an origin-less TrackedReal
    that no instruction refers to (``tracked.jl:259-265``): returned as is."""
    if not isinstance(t, TrackedReal):
        return t
    return _scalar_identity(t)


def scalar_buffer(t: "TrackedReal") -> int:
    """Size-1 buffer that carries ``t`` (operand class TR of array instructions; ``output_hook`` of a GradientTape)."""
    prog = t.tape.program
    if t.block is None:
        if prog.buffers[t.buffer].kind == P.BUF_TR and prog.buffers[t.buffer].size == 1:
            return t.buffer
        t = _scalar_identity(t)       # element of an array: route it through a slot of its own
    return _slot_buffer(t, want_whole=True)[0]


def collect(xs, tape=None):
    """``[t1, t2, ...]`` — an ``Array{TrackedReal}`` used as a tape OUTPUT (``seed!(output, i)`` works on any array of
    tracked scalars, ``api/utils.jl:44-58``): element ``k`` of a fresh buffer mirrors ``xs[k]``."""
    xs = list(np.asarray(xs, dtype=object).reshape(-1, order="F")) if not isinstance(xs, (list, tuple)) else list(xs)
    tp = tape if tape is not None else _tape_of(*[x for x in xs if isinstance(x, TrackedReal)])
    prog = tp.program
    buf = prog.add_buffer(P.BUF_TA, (len(xs),))
    vals = np.zeros(len(xs), tp.dtype)
    for k, x in enumerate(xs):
        if not isinstance(x, TrackedReal):
            x = _scalar_const(tp, x)
        elif x.block is None:
            x = _scalar_identity(x)
        x.block.export(x.slot, buf, k)
        vals[k] = x.value
    return TrackedArray(vals, tp, buf)


class Adjoint:
    """Lazy ``Adjoint``/``Transpose`` of a tracked or plain array (real eltype: the two coincide)."""

    __slots__ = ("parent",)
    __array_ufunc__ = None

    def __init__(self, parent):
        self.parent = parent

    @property
    def shape(self):
        s = self.parent.shape
        return (1, s[0]) if len(s) == 1 else (s[1], s[0])

    @property
    def T(self):
        return self.parent

    def __matmul__(self, other):
        return record_mul(self, other)

    def __rmatmul__(self, other):
        return record_mul(other, self)


def value(x):
    """``value`` (``src/tracked.jl:97-103``)."""
    if isinstance(x, (TrackedArray, TrackedReal)):
        return x.value
    if isinstance(x, Adjoint):
        v = value(x.parent)
        return v.reshape(1, -1) if v.ndim == 1 else v.T
    return x


def istracked(x):
    if isinstance(x, Adjoint):
        return istracked(x.parent)
    return isinstance(x, (TrackedArray, TrackedReal))


def _tape_of(*xs):
    for x in xs:
        if isinstance(x, Adjoint):
            x = x.parent
        if isinstance(x, (TrackedArray, TrackedReal)):
            return x.tape
    raise TypeError("no tracked argument")


def placeholder(shape, dtype):
    """A read-only array of the given shape that owns ONE element (all strides zero): what a shapes-only tape holds as values."""
    shape = tuple(int(d) for d in shape)
    return np.lib.stride_tricks.as_strided(np.zeros(1, dtype), shape=shape, strides=(0,) * len(shape), writeable=False)


def _track_new(v, tp: InstructionTape, shape=None):
    """Track a freshly computed value.  ``v`` may be a thunk (called only when the tape records values); ``shape`` is the output
    shape a shapes-only tape needs instead."""
    if tp.shapes_only:
        if shape is None:
            raise TypeError("internal: this instruction has no shape rule for shapes-only recording")
        shape = tuple(int(d) for d in shape)
        if len(shape) == 0:
            return TrackedReal(tp.dtype.type(0), tp, tp.program.add_buffer(P.BUF_TR, ()))
        return TrackedArray(placeholder(shape, tp.dtype), tp, tp.program.add_buffer(P.BUF_TA, shape))
    if callable(v):
        v = v()
    v = np.asarray(v, tp.dtype)
    if v.ndim == 0:
        return TrackedReal(v[()], tp, tp.program.add_buffer(P.BUF_TR, ()))
    return TrackedArray(np.asfortranarray(v), tp, tp.program.add_buffer(P.BUF_TA, v.shape))


def _capture(x, tp: InstructionTape, bound_to=None) -> P.Operand:
    """``capture`` (``src/tracked.jl:223-225``): TrackedArray → itself; an array whose elements are
    TrackedReals (the ``Adjoint`` of a tracked array) → dense index-linked copy (IDX); plain arrays →
    frozen copy (CONST)."""
    bound = None
    if bound_to is not None:
        shp = tuple(x.shape)
        nd = len(bound_to)
        # index_bound(x, out) = CartesianIndex(size(x,1..N_out))   (propagation.jl:25-27)
        bound = tuple(shp[i] if i < len(shp) else 1 for i in range(nd))
    if isinstance(x, TrackedArray):
        return P.Operand(x.buffer, P.CLS_TA, True, x.shape, bound=bound)
    if isinstance(x, TrackedReal):
        return P.Operand(scalar_buffer(x), P.CLS_TR, True, (), bound=bound)
    if isinstance(x, Adjoint):
        par = x.parent
        if isinstance(par, TrackedArray):
            if par.ndim == 1:
                m = np.arange(par.size, dtype=np.int64)          # row vector: identity map
            elif par.size > COMPACT_TRANSPOSE_THRESHOLD:
                # same map, shipped as a tag instead of size(a) int64 entries; loader and oracle expand it themselves
                return P.Operand(par.buffer, P.CLS_IDX, True, x.shape, P.IDX_TRANSPOSE, None, bound=bound)
            else:
                m = P.transpose_map(par.shape[0], par.shape[1])
            return P.Operand(par.buffer, P.CLS_IDX, True, x.shape, P.IDX_EXPLICIT, m, bound=bound)
        x = value(x)                                             # plain array: frozen dense copy
    v = np.asarray(x, tp.dtype)
    if v.ndim == 0:
        raise TypeError("plain Real arguments select the :tracker broadcast implementation "
                        "(broadcast.jl:68-85), which is not lowered yet; close over the number instead")
    b = tp.program.add_buffer(P.BUF_CONST, v.shape, const_data=np.array(v, order="F", copy=True))
    return P.Operand(b, P.CLS_CONST, False, v.shape, bound=bound)


# ---------------------------------------------------------------------------------------- *
def record_mul(x, y):
    """``record_mul`` (``linalg/arithmetic.jl:171-177``; dispatch combos ``:186-240``)."""
    tp = _tape_of(x, y)
    vx, vy = np.asarray(value(x)), np.asarray(value(y))
    if vx.ndim not in (1, 2) or vy.ndim not in (1, 2):
        raise TypeError("`*` is defined for vectors and matrices")
    if vx.shape[-1] != vy.shape[0]:
        raise ValueError(f"`*`: inner dimensions do not match: {vx.shape} x {vy.shape}")
    oshape = (vx.shape[:-1] if vx.ndim == 2 else ()) + (vy.shape[1:] if vy.ndim == 2 else ())
    out = _track_new(lambda: vx @ vy, tp, shape=oshape)
    if isinstance(out, TrackedReal):
        raise TypeError("scalar-valued `*` (row vector times vector) is not lowered; use dot/sum")
    tp.record(P.OP_MUL, [_capture(x, tp), _capture(y, tp)], out.buffer)
    return out


# ------------------------------------------------------------------------------------- + / -
def record_plusminus(opcode, x, y):
    """``record_plus``/``record_minus`` (``linalg/arithmetic.jl:36-41,118-123``)."""
    if isinstance(x, Adjoint) or isinstance(y, Adjoint):
        raise TypeError("array +/- on lazy adjoints is not lowered")
    tp = _tape_of(x, y)
    vx, vy = np.asarray(value(x)), np.asarray(value(y))
    if vx.shape != vy.shape:
        raise ValueError("array +/- needs equal shapes (use broadcast('+', a, b) otherwise)")
    out = _track_new(lambda: vx + vy if opcode == P.OP_ADD else vx - vy, tp, shape=vx.shape)
    tp.record(opcode, [_capture(x, tp), _capture(y, tp)], out.buffer)
    return out


# --------------------------------------------------------------------------------- reductions
def sum(x, dims=None):  # noqa: A001 - mirrors Base.sum
    """``Base.sum(::TrackedArray)`` (``linalg/reductions.jl:8-13``); ``sum(x, dims)`` records a ``sum!`` instruction
    (``:39-46``).  ``dims`` is 1-based like Julia's."""
    if hasattr(x, "_inner_sum") and dims is None:
        return x._inner_sum()          # Array{TrackedReal{TrackedReal}}: generic scalar fold (nested.py)
    if not isinstance(x, TrackedArray):
        return np.sum(x) if dims is None else np.sum(x, axis=tuple(d - 1 for d in np.atleast_1d(dims)), keepdims=True)
    if dims is not None:
        return record_sum_dims(x, dims)
    tp = x.tape
    out = _track_new(lambda: np.sum(x.value), tp, shape=())
    tp.record(P.OP_SUM, [_capture(x, tp)], out.buffer)
    return out


def mean(x):
    """``Statistics.mean(::TrackedArray)`` (``linalg/reductions.jl:66-71``)."""
    if not isinstance(x, TrackedArray):
        return np.mean(x)
    tp = x.tape
    out = _track_new(lambda: np.mean(x.value), tp, shape=())
    tp.record(P.OP_MEAN, [_capture(x, tp)], out.buffer)
    return out


def record_sum_dims(x, dims):
    dims = tuple(int(d) for d in np.atleast_1d(dims))
    if any(d < 1 or d > x.ndim for d in dims):
        raise ValueError("sum: dims out of range")
    tp = x.tape
    out = _track_new(lambda: np.sum(x.value, axis=tuple(d - 1 for d in dims), keepdims=True), tp,
                     shape=tuple(1 if (i + 1) in dims else n for i, n in enumerate(x.shape)))
    # cache = index_bound(out, x)   (reductions.jl:43)
    tp.record(P.OP_SUMDIMS, [P.Operand(x.buffer, P.CLS_TA, True, x.shape, bound=out.shape)], out.buffer)
    return out


def dot(x, y):
    """``LinearAlgebra.dot`` (``linalg/reductions.jl:91-104``)."""
    if isinstance(x, Adjoint) or isinstance(y, Adjoint):
        raise TypeError("dot of lazy adjoints is not lowered")
    tp = _tape_of(x, y)
    vx, vy = np.asarray(value(x)), np.asarray(value(y))
    if vx.size != vy.size:
        raise ValueError("dot needs equal lengths")
    out = _track_new(lambda: np.dot(vx.reshape(-1, order="F"), vy.reshape(-1, order="F")), tp, shape=())
    tp.record(P.OP_DOT, [_capture(x, tp), _capture(y, tp)], out.buffer)
    return out


# ------------------------------------------------------------------------- elementwise closures
def _bshape(shapes):
    return np.broadcast_shapes(*shapes)


def _record_elementwise(opcode, f, args, require_same_shape=False):
    tp = _tape_of(*args)
    for a in args:
        if isinstance(a, Adjoint):
            raise TypeError("elementwise operations on lazy adjoints are not lowered")
        if (isinstance(a, TrackedReal) or np.ndim(value(a)) == 0) and opcode != P.OP_TRACKER_BCAST:
            raise TypeError("scalar (Real/TrackedReal) arguments are only valid in dot-broadcasts (the :tracker implementation, "
                            "broadcast.jl:68-85)")
    shapes = [tuple(np.shape(value(a))) for a in args]
    if require_same_shape and any(s != shapes[0] for s in shapes):
        raise ValueError("map needs equal shapes")
    nd = max(len(s) for s in shapes)
    # Julia broadcasting aligns LEADING dimensions (column-major), numpy aligns trailing ones
    padded = [s + (1,) * (nd - len(s)) for s in shapes]
    oshape = tuple(int(max(p[i] for p in padded)) for i in range(nd))
    for p in padded:
        for i in range(nd):
            if p[i] not in (1, oshape[i]):
                raise ValueError(f"arrays could not be broadcast to a common size: {shapes}")
    # which arguments the closure sees as Duals: every Real/array argument of ∇broadcast (broadcast.jl:142-153) and of
    # tracker_∇broadcast (dual(x::Real, p), broadcast.jl:218-224); only the tracked ones for @forward closures
    # (elementwise.jl:234-301); (broadcast,^) uses its own cached formulas (elementwise.jl:633-643)
    if opcode in (P.OP_MAP, P.OP_BCAST):
        dual = [istracked(a) for a in args]
    else:
        dual = [True] * len(args)
    flavor = 1 if opcode == P.OP_TRACKER_BCAST else (2 if opcode == P.OP_BCAST_POW else 0)
    ex = E.trace(f, len(args), dual=dual, pow_flavor=flavor)
    def values():
        vals = [np.asarray(value(a), tp.dtype).reshape(p, order="F") for a, p in zip(args, padded)]
        return np.broadcast_to(E.eval_value(ex, vals, tp.dtype), oshape)
    out = _track_new(values, tp, shape=oshape)
    ops = [_capture(a, tp, bound_to=oshape) for a in args]
    tp.record(opcode, ops, out.buffer, ex, len(args))
    return out


def bcast(f, *args):
    """Fused dot-broadcast ``f.(args...)`` → ONE ``∇broadcast`` instruction
    (``broadcast.jl:86-94,142-162``).  Partials are taken w.r.t. every array argument."""
    if isinstance(f, ForwardOptimize):
        f = f.f
    # get_implementation (broadcast.jl:68-85): any plain Real / TrackedReal argument selects tracker_∇broadcast
    if any(isinstance(a, TrackedReal) or np.ndim(value(a)) == 0 for a in args):
        return _record_tracker_bcast(f, args)
    return _record_elementwise(P.OP_NBCAST, f, args)


def _record_tracker_bcast(f, args):
    """``tracker_∇broadcast`` (``broadcast.jl:229-260``).  Plain ``Real`` arguments carry no adjoint (``dual(x, p) = x`` only for
    non-Reals... a plain Real IS dualised but its adjoint is discarded by ``_add_to_deriv!``), so they are closed over as literals;
    ``TrackedReal`` arguments stay operands (class TR) whose adjoint is the full reduction ``unbroadcast(::Number, Δ) = sum(Δ)``."""
    live, slots = [], []
    for a in args:
        if isinstance(a, (TrackedArray, TrackedReal)) or np.ndim(value(a)) > 0:
            slots.append(len(live)); live.append(a)
        else:
            slots.append(float(a))

    def g(*xs):
        return f(*[xs[s] if isinstance(s, int) else s for s in slots])
    return _record_elementwise(P.OP_TRACKER_BCAST, g, live)


def map_(f, *args):
    """``map(f::ForwardOptimize, x[, y])`` (``elementwise.jl:230-310``)."""
    if not isinstance(f, ForwardOptimize):
        raise TypeError("map_ lowers only @forward closures (wrap with forward(f))")
    if len(args) not in (1, 2):
        raise TypeError("map of a @forward closure takes 1 or 2 arrays")
    return _record_elementwise(P.OP_MAP, f.f, args, require_same_shape=True)


_INFIX = {"+": (P.OP_BCAST_ADD, lambda a, b: a + b), "-": (P.OP_BCAST_SUB, lambda a, b: a - b),
          "*": (P.OP_BCAST_MUL, lambda a, b: a * b), "/": (P.OP_BCAST_DIV, lambda a, b: a / b),
          "\\": (P.OP_BCAST_LDIV, lambda a, b: b / a), "^": (P.OP_BCAST_POW, lambda a, b: a ** b)}


def broadcast(f, *args):
    """``broadcast(f::ForwardOptimize, x[, y])`` (``elementwise.jl:230-310``) and the built-in
    infix instructions ``broadcast(+|-|*|/|\\|^, x, y)`` (``elementwise.jl:449-685``)."""
    if isinstance(f, str):
        if f not in _INFIX or len(args) != 2:
            raise TypeError(f"unsupported infix broadcast {f!r}")
        op, fn = _INFIX[f]
        return _record_elementwise(op, fn, args)
    if not isinstance(f, ForwardOptimize):
        raise TypeError("broadcast lowers only @forward closures or '+', '-', '*', '/', '\\\\', '^'")
    if len(args) not in (1, 2):
        raise TypeError("broadcast of a @forward closure takes 1 or 2 arrays")
    return _record_elementwise(P.OP_BCAST, f.f, args)


# ----------------------------------------------------------------------------------- getindex
def record_getindex(t: TrackedArray, key):
    """Range/colon ``getindex`` (``src/tracked.jl:308-316``); the index iterable
    (``index_iterable``, ``:291-306``) is shipped as an explicit 0-based linear map."""
    if isinstance(key, list) and key and all(isinstance(k, tuple) for k in key):
        # getindex(t, inds::AbstractArray{<:CartesianIndex}) (tracked.jl:342-347): 0-based index tuples here
        lin = np.arange(t.size, dtype=np.int64).reshape(t.shape, order="F")
        m = np.asarray([lin[k] for k in key], dtype=np.int64)
        tp = t.tape
        out = _track_new(lambda: t.value.reshape(-1, order="F")[m], tp, shape=m.shape)
        tp.record(P.OP_GETINDEX, [P.Operand(t.buffer, P.CLS_TA, True, t.shape, P.IDX_EXPLICIT, m)], out.buffer)
        return out
    if not isinstance(key, tuple):
        key = (key,)
    if all(isinstance(k, (int, np.integer)) for k in key):
        # scalar getindex (tracked.jl:348-351): TrackedReal(value[i], deriv[i], tape, index, origin) — no instruction
        if len(key) == 1:
            lin = int(key[0]) + (t.size if key[0] < 0 else 0)
            if not 0 <= lin < t.size:
                raise IndexError("index out of range")
        else:
            if len(key) != t.ndim:
                raise IndexError("wrong number of indices")
            lin = int(np.ravel_multi_index(tuple(int(k) % d for k, d in zip(key, t.shape)), t.shape, order="F"))
        return TrackedReal(t.tape.dtype.type(0) if t.tape.shapes_only else t.value.reshape(-1, order="F")[lin], t.tape, t.buffer, lin)
    generic = not all(isinstance(k, slice) for k in key)
    if generic:
        # (getindex, Val(:generic)): Integer / Colon / integer-vector per dimension (tracked.jl:352-362)
        key = tuple(np.asarray(k, dtype=np.int64) if isinstance(k, (list, np.ndarray)) else k for k in key)
        if not all(isinstance(k, (slice, int, np.integer, np.ndarray)) for k in key):
            raise TypeError("unsupported index type")
    lin = np.arange(t.size, dtype=np.int64).reshape(t.shape, order="F")
    if len(key) == 1 and t.ndim > 1:
        sub = lin.reshape(-1, order="F")[key[0]]       # linear indexing x[a:b]
    else:
        if len(key) != t.ndim:
            raise IndexError("wrong number of indices")
        if generic:                                    # Julia indexes the outer product of the per-dimension index sets
            sub = lin[np.ix_(*[np.arange(t.shape[d])[k] if isinstance(k, slice) else np.atleast_1d(k) for d, k in enumerate(key)])]
            sub = sub.reshape([n for n, k in zip(sub.shape, key) if not isinstance(k, (int, np.integer))], order="F")
        else:
            sub = lin[key]
    m = np.ascontiguousarray(sub.reshape(-1, order="F"))
    tp = t.tape
    out = _track_new(lambda: t.value.reshape(-1, order="F")[m].reshape(sub.shape, order="F"), tp, shape=sub.shape)
    op = P.Operand(t.buffer, P.CLS_TA, True, t.shape, P.IDX_EXPLICIT, m)
    tp.record(P.OP_GETINDEX_GENERIC if generic else P.OP_GETINDEX, [op], out.buffer)
    return out
