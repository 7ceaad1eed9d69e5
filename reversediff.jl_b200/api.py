"""Tape objects and the replay API — host-side mirror of the reference's L4/L5 layers for the
compiled-tape path only.

==========================================  ===========================================================
reference (Julia)                           here (Python spelling; Julia shim in julia/ReverseDiffB200.jl)
==========================================  ===========================================================
``GradientTape(f, input)``                  ``GradientTape(f, input)``            api/tape.jl:197-209
``JacobianTape(f, input)``                  ``JacobianTape(f, input)``            api/tape.jl:226-238
``HessianTape(f, input)``                   ``HessianTape(f, input)``             api/tape.jl:285-293
``compile(t)``                              ``compile(t)``                        api/tape.jl:146-149
``gradient!(result, ct, input)``            ``gradient_(result, ct, input)``      api/gradients.jl:78-82
``jacobian!(result, ct, input)``            ``jacobian_(result, ct, input)``      api/jacobians.jl:120-124
``hessian!(result, ct, input)``             ``hessian_(result, ct, input)``       api/hessians.jl:86-97
``DiffResults.GradientResult`` etc.         ``DiffResult``                        api/utils.jl:91-111
==========================================  ===========================================================

Uncompiled tapes cannot be replayed here: there is deliberately no CPU path — ``compile`` lowers
to the B200 engine or raises.
"""
from __future__ import annotations

import numpy as np

from . import program as P
from .tracked import InstructionTape, TrackedArray, TrackedReal, value, scalar_buffer, collect
from . import engine as _engine


class DiffResult:
    """Minimal ``DiffResults.MutableDiffResult``: ``value`` plus derivative arrays."""

    def __init__(self, value_, *derivs):
        self.value = value_
        self.derivs = list(derivs)

    @property
    def gradient(self):
        return self.derivs[0]

    @property
    def jacobian(self):
        return self.derivs[0]

    @property
    def hessian(self):
        return self.derivs[1]


class AbstractTape:
    kind = "abstract"

    def __init__(self, func, input_, dtype=None):
        self.func = func
        self.is_tuple = isinstance(input_, (tuple, list))
        xs = list(input_) if self.is_tuple else [input_]
        # Record-time inputs that already live on the device (torch CUDA tensors, same shape conventions as at replay: an N-D
        # tensor of the recorded shape).  When EVERY input is a device array the tape is recorded SHAPES-ONLY (SURVEY section 8 f.4,
        # TrackedArray{V,D,N,VA,DA} with device storage, src/tracked.jl:70-87): no element of any input crosses to the host, the
        # operator-overloading pass runs on zero-stride placeholders and infers shapes (tracked.InstructionTape.shapes_only), and
        # the values the reference's recording pass would have computed come out of the first forward sweep on the device.  The
        # program is the same byte for byte (it never depended on values); a function whose control flow depends on values cannot
        # be recorded this way and says so.
        shapes_only = any(_engine._is_device(x) for x in xs)      # (host inputs of such a tape are not read either: replay supplies every value)
        if shapes_only and dtype is None:
            names = {str(x.dtype).replace("torch.", "") for x in xs}
            dtype = np.float32 if names == {"float32"} else np.float64     # (np.result_type of the reference's promote rules: any Float64 wins)
        if shapes_only:
            from .tracked import placeholder
            xs = [placeholder(tuple(x.shape), np.dtype(dtype)) for x in xs]
        else:
            xs = [np.asarray(x) for x in xs]
        if dtype is None:
            dtype = np.result_type(*[np.asarray(x).dtype for x in xs])
            if dtype not in (np.float32, np.float64):
                dtype = np.float64
        self.tape = InstructionTape(dtype)
        self.tape.shapes_only = shapes_only
        prog = self.tape.program
        # GradientConfig(input): track(similar(x), D, tp) then track!(cfg.input, input)
        # (api/Config.jl:37,45-47 ; tracked.jl:476-478)
        self.input = []
        for x in xs:
            v = x if shapes_only else np.array(x, dtype=self.tape.dtype, order="F", copy=True)
            if v.ndim == 0:
                raise TypeError("tape inputs must be arrays")
            b = prog.add_buffer(P.BUF_TA, v.shape)
            prog.inputs.append(b)
            self.input.append(TrackedArray(v, self.tape, b))
        self.output = func(*self.input) if self.is_tuple else func(self.input[0])
        if isinstance(self.output, (list, tuple)) or (isinstance(self.output, np.ndarray) and self.output.dtype == object):
            self.output = collect(self.output, self.tape)       # Array{TrackedReal} output (api/utils.jl:44-58 seeds it elementwise)
        self._check_output()
        prog.output_buffer = scalar_buffer(self.output) if isinstance(self.output, TrackedReal) else self.output.buffer

    def _check_output(self):
        raise NotImplementedError

    def __len__(self):
        return len(self.tape)

    # hooks (api/tape.jl:28-32)
    def func_hook(self): return self.func
    def input_hook(self): return tuple(self.input) if self.is_tuple else self.input[0]
    def output_hook(self): return self.output

    def program_bytes(self) -> bytes:
        return self.tape.program.to_bytes()


class GradientTape(AbstractTape):
    kind = "gradient"

    def _check_output(self):
        if not isinstance(self.output, TrackedReal):
            raise TypeError("GradientTape: f must return a tracked scalar (e.g. end in sum/mean)")


class JacobianTape(AbstractTape):
    kind = "jacobian"

    def _check_output(self):
        if not isinstance(self.output, TrackedArray):
            raise TypeError("JacobianTape: f must return a tracked array")


class HessianTape(AbstractTape):
    """``HessianTape(f, input)`` (``api/tape.jl:285-293``).  Two recordings of the same Hessian:

    ``mode="forward"`` (default) records the *inner array tape* of ``f``; the engine replays it forward-over-reverse in Dual
    chunks (BASELINE.json north star) — NOT the reference's mechanism, same result.

    ``mode="reverse"`` records what the reference records: the OUTER scalar tape of a reverse-over-reverse sweep
    (``api/Config.jl:152-156``; see ``nested.py``), i.e. the program ``x -> grad f(x)``; ``hessian!`` is then literally
    ``seeded_reverse_pass!(result, output::Vector, input, tape)`` (``api/utils.jl:44-58``): one one-hot seeded reverse sweep of
    that scalar tape per gradient component, R of them per device sweep."""
    kind = "hessian"

    def __init__(self, func, input_, dtype=None, mode="forward"):
        if isinstance(input_, (tuple, list)):
            # api/hessians.jl:103-108
            raise TypeError("Taking the Hessian of a function with multiple arguments is not yet supported")
        if mode not in ("forward", "reverse"):
            raise ValueError("mode must be 'forward' (forward-over-reverse on the array tape) or 'reverse' (the reference's "
                             "reverse-over-reverse scalar tape)")
        self.mode = mode
        if mode == "forward":
            super().__init__(func, input_, dtype)
            return
        from .nested import record_gradient_program
        self.func = func
        self.is_tuple = False
        dev_in = _engine._is_device(input_)
        if dev_in:                                               # device input: shapes-only, as in AbstractTape.__init__
            dt = np.dtype(dtype) if dtype is not None else np.dtype(np.float32 if str(input_.dtype).endswith("float32") else np.float64)
            v = np.zeros(tuple(input_.shape), dt, order="F")     # (the nested recorder indexes element by element: a real array of zeros)
        else:
            v = np.array(input_, dtype=np.float64 if dtype is None else dtype, order="F", copy=True)
        if v.dtype not in (np.float32, np.float64):
            v = v.astype(np.float64)
        self.tape = InstructionTape(v.dtype)                     # jtp: the tape that is kept (api/tape.jl:288)
        self.tape.shapes_only = dev_in
        prog = self.tape.program
        b = prog.add_buffer(P.BUF_TA, v.shape)
        prog.inputs.append(b)
        self.input = [TrackedArray(v, self.tape, b)]             # jcfg.input
        y, grad = record_gradient_program(func, self.input[0])   # GradientTape(f, ht.input, gcfg) + seeded_reverse_pass!
        self.primal = y.value                                    # f(x) as an outer-tape scalar
        # value!(result, ...) of the DiffResult method (hessians.jl:92-97) reads it back from the engine
        self.primal_buffer = scalar_buffer(self.primal) if isinstance(self.primal, TrackedReal) else -1
        self.output = collect(grad, self.tape)                   # ht.output: the gradient, an array of outer-tape scalars
        prog.output_buffer = self.output.buffer

    def _check_output(self):
        if not isinstance(self.output, TrackedReal):
            raise TypeError("HessianTape: f must return a tracked scalar")


class CompiledTape:
    """``CompiledTape`` (``api/tape.jl:55-59``): here a handle on an engine-resident tape."""

    def __init__(self, tape: AbstractTape, ctx=None, **options):
        self.tape = tape
        self.kind = tape.kind
        self.ctx = ctx if ctx is not None else _engine.default_context()
        self.program = tape.program_bytes()
        self.handle = _engine.EngineTape(self.ctx, self.program, **options)
        self.dtype = tape.tape.dtype

    def __len__(self):
        return len(self.tape)

    def func_hook(self): return self.tape.func_hook()
    def input_hook(self): return self.tape.input_hook()
    def output_hook(self): return self.tape.output_hook()

    # forward_pass!/reverse_pass! (api/tape.jl:122-134)
    def forward_pass(self): self.handle.forward()
    def reverse_pass(self): self.handle.reverse_scalar_seed()


def compile(t: AbstractTape, ctx=None, **options) -> CompiledTape:  # noqa: A001 - mirrors ReverseDiff.compile
    if isinstance(t, CompiledTape):
        return t
    return CompiledTape(t, ctx, **options)


# ---------------------------------------------------------------------------------------------
def _as_list(x, is_tuple):
    return list(x) if is_tuple else [x]


def _is_device(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _need_compiled(ct, kind):
    if not isinstance(ct, CompiledTape):
        raise TypeError(f"{kind}! replays only compiled tapes here: call compile(tape) first "
                        "(uncompiled replay would be a CPU path; this engine has none)")
    if ct.kind != kind:
        raise TypeError(f"{kind}! needs a compiled {kind.capitalize()}Tape, got a {ct.kind} tape")


def seeded_forward_pass(ct: CompiledTape, input_):
    """``seeded_forward_pass!`` (``api/tape.jl:40-44``)."""
    xs = _as_list(input_, ct.tape.is_tuple)
    if len(xs) != len(ct.tape.input):
        raise ValueError("wrong number of inputs for this tape")
    for i, x in enumerate(xs):
        ct.handle.set_input(i, x, tuple(ct.tape.input[i].shape))
    return None


def _result_arrays(result, n):
    """Unpack result containers the way ``extract_result!`` dispatches (``api/utils.jl:77-111``)."""
    if isinstance(result, DiffResult):
        return [result.derivs[0]], [result]
    if isinstance(result, (tuple, list)):
        arrs, drs = [], []
        for r in result:
            if isinstance(r, DiffResult):
                arrs.append(r.derivs[0]); drs.append(r)
            else:
                arrs.append(r)
        return arrs, drs
    return [result], []


def gradient_(result, ct: CompiledTape, input_):
    """``gradient!(result, tape::CompiledGradient, input)`` (``api/gradients.jl:78-82``)."""
    _need_compiled(ct, "gradient")
    seeded_forward_pass(ct, input_)
    arrs, drs = _result_arrays(result, len(ct.tape.input))
    if len(arrs) != len(ct.tape.input):
        raise ValueError("result does not match the tape's inputs")
    val = ct.handle.gradient(arrs, [tuple(t.shape) for t in ct.tape.input], want_value=bool(drs))
    for d in drs:
        d.value = val
    return result


class GradientPipeline:
    """Consecutive ``gradient!`` evaluations of one tape, pipelined over ``depth`` engine copies of it (each on a context = stream
    set of its own): ``submit`` enqueues an evaluation (``rdb_tape_set_input`` + ``rdb_gradient_async``) and returns at once, so the
    upload of evaluation k+1 and the download of k-1 run under the sweeps of k; ``submit`` on a busy slot and ``drain`` complete
    the oldest call first (``rdb_tape_wait``).  Results land where ``gradient!`` would put them and are valid once the call has
    been completed; inputs and results must stay untouched until then.  The reference has no counterpart (its replay is
    synchronous CPU code); what is kept is the per-call contract of ``gradient!(result, tape, input)`` (``api/gradients.jl:78-82``)."""

    def __init__(self, tape, depth=2, device=None, **options):
        if isinstance(tape, CompiledTape):
            device = tape.ctx.device_id if device is None else device
            tape = tape.tape
        dev = 0 if device is None else device
        self.slots = [CompiledTape(tape, _engine.Context(dev), **options) for _ in range(depth)]
        self._inflight = [None] * depth
        self._next = 0

    def submit(self, result, input_):
        k = self._next
        self._next = (k + 1) % len(self.slots)
        self._complete(k)
        ct = self.slots[k]
        _need_compiled(ct, "gradient")
        seeded_forward_pass(ct, input_)
        arrs, drs = _result_arrays(result, len(ct.tape.input))
        if len(arrs) != len(ct.tape.input):
            raise ValueError("result does not match the tape's inputs")
        ct.handle.gradient_async(arrs, [tuple(t.shape) for t in ct.tape.input], want_value=bool(drs))
        self._inflight[k] = (drs, input_, result)
        return result

    def _complete(self, k):
        if self._inflight[k] is None:
            return
        drs, _inp, _res = self._inflight[k]
        self._inflight[k] = None
        val = self.slots[k].handle.wait()
        for d in drs:
            d.value = val

    def drain(self):
        """Complete every evaluation in flight, oldest first."""
        n = len(self.slots)
        for j in range(n):
            self._complete((self._next + j) % n)

    def launch_count(self):
        return sum(ct.handle.launch_count() for ct in self.slots)


def gradient(ct: CompiledTape, input_):
    """``gradient!(tape, input)`` (``api/gradients.jl:61-65``): allocates via construct_result."""
    res = [np.zeros(t.shape, ct.dtype, order="F") for t in ct.tape.input]
    out = tuple(res) if ct.tape.is_tuple else res[0]
    return gradient_(out, ct, input_)


def jacobian_(result, ct: CompiledTape, input_, rows=None):
    """``jacobian!(result, tape::CompiledJacobian, input)`` (``api/jacobians.jl:120-124``).  ``rows=(b,e)``
    restricts to a block of output-seed rows (the multi-GPU shard); ``result`` then holds ``e-b`` rows."""
    _need_compiled(ct, "jacobian")
    seeded_forward_pass(ct, input_)
    m = ct.tape.output.size
    b, e = (0, m) if rows is None else rows
    arrs, drs = _result_arrays(result, len(ct.tape.input))
    if len(arrs) != len(ct.tape.input):
        raise ValueError("result does not match the tape's inputs")
    val = None
    for slot, arr in enumerate(arrs):    # tuple inputs: the sweep loop is repeated per input (api/utils.jl:66-71)
        n = ct.tape.input[slot].size
        val = ct.handle.jacobian(slot, b, e, arr, (e - b, n), want_value=bool(drs), out_shape=ct.tape.output.shape)
    for d in drs:
        d.value = val
    return result


def jacobian(ct: CompiledTape, input_):
    m = ct.tape.output.size
    res = [np.zeros((m, t.size), ct.dtype, order="F") for t in ct.tape.input]
    out = tuple(res) if ct.tape.is_tuple else res[0]
    return jacobian_(out, ct, input_)


def hessian_(result, ct: CompiledTape, input_, cols=None, rows=None):
    """``hessian!(result, tape::CompiledHessian, input)`` (``api/hessians.jl:86-97``).  ``cols=(b,e)`` restricts a forward-mode
    tape to a block of columns, ``rows=(b,e)`` a reverse-mode (reference-style) tape to a block of gradient components — the
    multi-GPU shards; ``result`` then holds ``n x (e-b)`` resp. ``(e-b) x n``."""
    _need_compiled(ct, "hessian")
    seeded_forward_pass(ct, input_)
    n = ct.tape.input[0].size
    if getattr(ct.tape, "mode", "forward") == "reverse":
        if cols is not None:
            raise ValueError("a reverse-mode HessianTape shards by rows (gradient components): pass rows=(b, e)")
        b, e = (0, n) if rows is None else rows
        if isinstance(result, DiffResult):
            # seeded_reverse_pass!(DiffResult(gradient(result), hessian(result)), tape); value!(result, f(input))  (hessians.jl:92-97)
            if _is_device(result.derivs[1]) or _is_device(result.derivs[0]):
                raise TypeError("hessian! of a reverse-mode HessianTape into a DiffResult needs host arrays (the gradient is the "
                                "tape's output value, extracted on the host); pass a plain device matrix for the Hessian alone")
            g = ct.handle.jacobian(0, b, e, result.derivs[1], (e - b, n), want_value=True, out_shape=(n,))
            _store(result.derivs[0], g)
            # the reference re-evaluates f(input) on the CPU here; the primal is already on the tape (recorded by
            # record_gradient_program), so it is read back from the engine instead - no host evaluation of f
            result.value = ct.handle.get_value(ct.tape.primal_buffer)[0]
        else:
            ct.handle.jacobian(0, b, e, result, (e - b, n), want_value=False, out_shape=(n,))
        return result
    if rows is not None:
        raise ValueError("a forward-mode HessianTape shards by columns: pass cols=(b, e)")
    b, e = (0, n) if cols is None else cols
    if isinstance(result, DiffResult):
        dev_g = result.derivs[0] if _is_device(result.derivs[0]) else None
        g, val = ct.handle.hessian(b, e, result.derivs[1], (n, e - b), want_grad=True, grad_out=dev_g)
        if dev_g is None:
            _store(result.derivs[0], g)
        result.value = val
    else:
        ct.handle.hessian(b, e, result, (n, e - b), want_grad=False)
    return result


def _store(dst, flat):
    """Write a flat column-major vector into a host result array of any layout (a reshape of a C-ordered array is a copy, so
    the assignment goes through the array itself)."""
    dst[...] = np.asarray(flat).reshape(dst.shape, order="F")


def shard_range(n: int, world: int, rank: int):
    """Contiguous block of output-seed rows (jacobian!) or Hessian columns owned by `rank` of `world`."""
    per = n // world
    return rank * per, ((rank + 1) * per if rank < world - 1 else n)


def hessian(ct: CompiledTape, input_):
    n = ct.tape.input[0].size
    return hessian_(np.zeros((n, n), ct.dtype, order="F"), ct, input_)


# --------------------------------------------------------------------------------------------- multi-GPU
def shard_bounds(n: int, world: int):
    """``world + 1`` offsets of the partition ``shard_range`` induces (identical on every rank)."""
    return [shard_range(n, world, r)[0] for r in range(world)] + [n]


def make_comm(ctx=None):
    """One ``rdb_comm`` per process.  The engine has no bootstrap of its own: rank 0's NCCL id travels through the host's
    ``torch.distributed`` group (any backend; Julia would use MPI.jl), after that all traffic is NCCL inside librdb200."""
    import torch.distributed as dist
    ctx = ctx if ctx is not None else _engine.default_context()
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [_engine.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return _engine.Comm(ctx, world, rank, box[0])


def gather_jacobian_(result, block, ct: CompiledTape, comm, slot=0, root=-1):
    """Place every rank's ``jacobian_(block, ct, x, rows=shard_range(m, world, rank))`` block into the reference's result: the
    column-major ``length(output) x length(input)`` matrix (``result_matrix``, ``api/utils.jl:44-58``), on every rank
    (``root=-1``) or on ``root``.  ``block`` and ``result`` are flat device tensors."""
    m, n = ct.tape.output.size, ct.tape.input[slot].size
    comm.gather(ct.dtype, 0, block, shard_bounds(m, comm.world), n, result, root)
    return result


def gather_hessian_(result, block, ct: CompiledTape, comm, root=-1):
    """Same for ``hessian_(block, ct, x, cols=shard_range(n, world, rank))`` column blocks (contiguous in the column-major
    Hessian, ``api/hessians.jl:86-90``); a reverse-mode (reference-style) tape shards by rows like ``jacobian!``."""
    n = ct.tape.input[0].size
    kind = 0 if getattr(ct.tape, "mode", "forward") == "reverse" else 1
    comm.gather(ct.dtype, kind, block, shard_bounds(n, comm.world), n, result, root)
    return result


_ = value
