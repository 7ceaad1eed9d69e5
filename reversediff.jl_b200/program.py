"""Flat, serialisable *tape program*: the object that crosses the C ABI.

A ReverseDiff tape is a ``Vector{AbstractInstruction}`` whose entries hold
``(func, input, output, cache)`` (reference ``src/tape.jl:52-57``) and whose
operands are ``TrackedArray`` buffer pairs (``src/tracked.jl:70-87``),
index-linked ``TrackedReal`` arrays (``src/tracked.jl:47-57,348-351``) or frozen
plain arrays (``capture``, ``src/tracked.jl:223-225``).  None of that can cross a
``ccall``, so ``compile`` lowers it to the byte layout below.  The same bytes are
parsed by the C++ loader (``csrc/program.h``) and by the numpy oracle
(``oracle/replay.py``), which is what makes "same instruction order and index
maps" checkable.

Layout (little-endian, every section 8-byte aligned)::

    Header        128 bytes   (HEADER_DTYPE)
    BufferRec[]   n_buffers   (BUFFER_DTYPE)      value/deriv buffer table
    int32[]       n_inputs    buffer ids in input_hook order (api/tape.jl:30)
    InstrRec[]    n_instr     (INSTR_DTYPE)       tape[1..end] order (tape.jl:75-80)
    int64[]       n_idx       index maps, 0-based linear indices into the origin
    int32[]       n_expr      scalar-expression bytecode, 4 int32 per op
    float64[]     n_fconst    literal constants used by the bytecode
    bytes         n_const     frozen CONST array data, column-major
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x3030324244525052  # "RPRDB200" little-endian-ish tag
VERSION = 2

# dtypes ---------------------------------------------------------------------------------
F64, F32 = 0, 1
NP_DTYPE = {F64: np.float64, F32: np.float32}

# buffer kinds ---------------------------------------------------------------------------
BUF_TA, BUF_TR, BUF_CONST = 0, 1, 2

# operand classes -----------------------------------------------------
CLS_TA, CLS_IDX, CLS_CONST, CLS_TR = 0, 1, 2, 3

# IDX map encodings
IDX_NONE, IDX_EXPLICIT, IDX_TRANSPOSE = 0, 1, 2

# instruction opcodes -------------------------------------------------------------------
OP_MUL = 1        # `*`            linalg/arithmetic.jl:171-177
OP_ADD = 2        # array `+`      linalg/arithmetic.jl:36-41
OP_SUB = 3        # array `-`      linalg/arithmetic.jl:118-123
OP_NEG = 4        # unary `-`      linalg/arithmetic.jl:111-116
OP_SUM = 5        # sum            linalg/reductions.jl:8-13
OP_MEAN = 6       # mean           linalg/reductions.jl:66-71
OP_NBCAST = 7     # ∇broadcast     broadcast.jl:142-162
OP_MAP = 8        # map(@forward f, ...)        elementwise.jl:230-310
OP_BCAST = 9      # broadcast(@forward f, ...)  elementwise.jl:230-310
OP_BCAST_ADD = 10  # (broadcast,+)  elementwise.jl:449-455
OP_BCAST_SUB = 11  # (broadcast,-)  elementwise.jl:480-486
OP_BCAST_MUL = 12  # (broadcast,*)  elementwise.jl:511-517
OP_BCAST_DIV = 13  # (broadcast,/)  elementwise.jl:555-563
OP_BCAST_LDIV = 14  # (broadcast,\\) elementwise.jl:594-602
OP_BCAST_POW = 15  # (broadcast,^)  elementwise.jl:651-659
OP_GETINDEX = 16  # getindex       tracked.jl:308-316
OP_DOT = 17       # dot            linalg/reductions.jl:91-104
OP_SUMDIMS = 18   # sum!/sum(x,dims) linalg/reductions.jl:32-46
OP_TRACKER_BCAST = 19  # tracker_∇broadcast (a plain Real or TrackedReal among the args)  broadcast.jl:229-238
OP_GETINDEX_GENERIC = 20  # (getindex, Val(:generic)): integer-vector indices  tracked.jl:352-362
OP_SCALAR_BLOCK = 21  # a run of consecutive ScalarInstructions (tape.jl:30-45; derivatives/scalars.jl:52-123), see ScalarBlock
OP_MORE_OPERANDS = 22  # not an instruction: operands 5.. of the record before it (`∇broadcast` is N-ary, broadcast.jl:142-146)
OP_NAMES = {
    OP_MUL: "*", OP_ADD: "+", OP_SUB: "-", OP_NEG: "neg", OP_SUM: "sum", OP_MEAN: "mean",
    OP_NBCAST: "∇broadcast", OP_MAP: "map", OP_BCAST: "broadcast",
    OP_BCAST_ADD: "(broadcast,+)", OP_BCAST_SUB: "(broadcast,-)", OP_BCAST_MUL: "(broadcast,*)",
    OP_GETINDEX: "getindex", OP_BCAST_DIV: "(broadcast,/)", OP_BCAST_LDIV: "(broadcast,\\)", OP_BCAST_POW: "(broadcast,^)",
    OP_DOT: "dot", OP_SUMDIMS: "sum!", OP_TRACKER_BCAST: "tracker_∇broadcast", OP_GETINDEX_GENERIC: "(getindex,:generic)",
    OP_SCALAR_BLOCK: "scalar-block",
}

# operand spaces of a scalar instruction (ScalarBlock payload)
SP_POOL = -1      # TrackedReal created by an earlier instruction of the SAME block: idx = its instruction number
SP_LITERAL = -2   # untracked Real (capture(::Real), tape.jl:24): idx = index into the fconst pool
SP_NONE = -3      # absent second input of a unary instruction
# space >= 0: ordinal into the block's referenced-buffer list; idx = 0-based linear element of that buffer.  This is a
# TrackedReal WITH an origin (tracked.jl:348-351: scalar getindex records nothing, value/deriv are pulled from / pushed to
# origin[index], tracked.jl:178-197), or a TrackedReal produced by an array instruction / exported by another block.

# scalar-expression bytecode: SSA, 4 x int32 per op = (code, a, b, imm) -----------------
# registers 0..N-1 are the N arguments; op j writes register N+j; the last op is the result.
E_CONST = 1    # r = fconst[imm]
E_ADD, E_SUB, E_MUL, E_DIV = 2, 3, 4, 5
E_NEG = 6
E_POWI = 7     # r = a ** imm (integer literal power, by repeated product like Base.literal_pow)
E_POW = 8
E_TANH, E_EXP, E_LOG, E_SQRT, E_SIN, E_COS = 9, 10, 11, 12, 13, 14
E_ABS2, E_ABS = 15, 16
E_MAX, E_MIN = 17, 18
E_INV = 19
E_SIGMOID = 20  # logistic (LogExpFunctions.logistic)
E_ARG = 21      # r = arg[imm] (identity; used when the closure returns an argument)
# second tranche of DiffRules.diffrules() functions (src/derivatives/scalars.jl:5-22, elementwise.jl:70-76)
E_LOG1P, E_EXPM1, E_ASIN, E_ACOS, E_ATAN, E_SINH, E_COSH, E_TAN = 22, 23, 24, 25, 26, 27, 28, 29
E_ATAN2 = 30    # atan(y, x)
E_HYPOT = 31
# third tranche (exp2 exp10 log2 log10 cbrt asinh acosh atanh sec csc cot sech csch coth), all unary, all in DiffRules.diffrules()
(E_EXP2, E_EXP10, E_LOG2, E_LOG10, E_CBRT, E_ASINH, E_ACOSH, E_ATANH, E_SEC, E_CSC, E_COT, E_SECH, E_CSCH, E_COTH) = range(32, 46)
E_LAST = E_COTH
E_NAMES = {
    E_CONST: "const", E_ADD: "add", E_SUB: "sub", E_MUL: "mul", E_DIV: "div", E_NEG: "neg",
    E_POWI: "powi", E_POW: "pow", E_TANH: "tanh", E_EXP: "exp", E_LOG: "log", E_SQRT: "sqrt",
    E_SIN: "sin", E_COS: "cos", E_ABS2: "abs2", E_ABS: "abs", E_MAX: "max", E_MIN: "min",
    E_INV: "inv", E_SIGMOID: "logistic", E_ARG: "arg",
    E_LOG1P: "log1p", E_EXPM1: "expm1", E_ASIN: "asin", E_ACOS: "acos", E_ATAN: "atan", E_SINH: "sinh", E_COSH: "cosh",
    E_TAN: "tan", E_ATAN2: "atan2", E_HYPOT: "hypot",
    E_EXP2: "exp2", E_EXP10: "exp10", E_LOG2: "log2", E_LOG10: "log10", E_CBRT: "cbrt", E_ASINH: "asinh", E_ACOSH: "acosh",
    E_ATANH: "atanh", E_SEC: "sec", E_CSC: "csc", E_COT: "cot", E_SECH: "sech", E_CSCH: "csch", E_COTH: "coth",
}
WIRE_ARGS = 4     # operand slots of one instruction record
MAX_ARGS = 8      # operands of one instruction; records 5.. travel in OP_MORE_OPERANDS records right behind it
MAX_DIMS = 4
MAX_EXPR_REGS = 32

HEADER_DTYPE = np.dtype([
    ("magic", "<u8"), ("version", "<u4"), ("dtype", "<u4"),
    ("n_buffers", "<u4"), ("n_instr", "<u4"), ("n_inputs", "<u4"), ("output_buffer", "<i4"),
    ("off_buffers", "<u8"), ("off_inputs", "<u8"), ("off_instr", "<u8"),
    ("off_idx", "<u8"), ("n_idx", "<u8"),
    ("off_expr", "<u8"), ("n_expr", "<u8"),
    ("off_fconst", "<u8"), ("n_fconst", "<u8"),
    ("off_const", "<u8"), ("n_const", "<u8"),
    ("total_bytes", "<u8"),
])
assert HEADER_DTYPE.itemsize == 128

BUFFER_DTYPE = np.dtype([
    ("kind", "<i4"), ("ndims", "<i4"), ("dims", "<i8", (MAX_DIMS,)),
    ("alias_of", "<i4"), ("pad", "<i4"), ("const_offset", "<i8"),
])
assert BUFFER_DTYPE.itemsize == 56

OPERAND_DTYPE = np.dtype([
    ("buffer", "<i4"), ("cls", "<i4"), ("tracked", "<i4"), ("idx_kind", "<i4"),
    ("idx_offset", "<i8"), ("idx_len", "<i8"),
    ("ndims", "<i4"), ("pad", "<i4"), ("dims", "<i8", (MAX_DIMS,)),
    ("bound", "<i8", (MAX_DIMS,)),
])
assert OPERAND_DTYPE.itemsize == 104

INSTR_DTYPE = np.dtype([
    ("opcode", "<i4"), ("n_in", "<i4"), ("out", "<i4"), ("n_expr_args", "<i4"),
    ("expr_offset", "<i8"), ("expr_len", "<i8"),
    ("in", OPERAND_DTYPE, (WIRE_ARGS,)),
])
assert INSTR_DTYPE.itemsize == 32 + 4 * 104


def _align8(n: int) -> int:
    return (n + 7) & ~7


class Operand:
    """One captured instruction input (``capture``: tape.jl:24-25,67-69; tracked.jl:223-225)."""

    __slots__ = ("buffer", "cls", "tracked", "idx_kind", "idx_map", "dims", "bound")

    def __init__(self, buffer, cls, tracked, dims, idx_kind=IDX_NONE, idx_map=None, bound=None):
        self.buffer = int(buffer)
        self.cls = int(cls)
        self.tracked = bool(tracked)
        self.dims = tuple(int(d) for d in dims)
        self.idx_kind = int(idx_kind)
        self.idx_map = idx_map            # np.int64 array (0-based) when IDX_EXPLICIT
        self.bound = tuple(int(b) for b in bound) if bound is not None else None


class Instr:
    __slots__ = ("opcode", "inputs", "out", "expr", "n_expr_args", "scalar")

    def __init__(self, opcode, inputs, out, expr=None, n_expr_args=0, scalar=None):
        self.opcode = int(opcode)
        self.inputs = list(inputs)
        self.out = int(out)
        self.expr = expr                  # (ops: int32[n,4], fconsts: float64[]) or None
        self.n_expr_args = int(n_expr_args)
        self.scalar = scalar              # ScalarBlock for OP_SCALAR_BLOCK


class ScalarBlock:
    """A maximal run of consecutive ``ScalarInstruction``s (``src/tape.jl:30-45``), shipped as ONE tape entry.

    Instruction ``i`` of the block is ``(func = exprs[expr_id[i]], input = (a, b), output = pool slot i)``; every scalar
    instruction creates exactly one new ``TrackedReal`` (``macros.jl:84-131``), so slot == instruction number.  A slot that
    anything OUTSIDE the block reads (the tape output, an operand of a later array instruction or of a later block) is
    *exported*: mirrored into element ``elem`` of a tape buffer, whose deriv carries the adjoint back into the block.

    Wire form (int64 words in the idx section, referenced by ``in[0].idx_offset/idx_len``)::

        n_instr, n_refbufs, n_exports, n_exprs,
        refbufs[n_refbufs],                                  buffer ids; operand spaces / export targets index this table
        exprs[n_exprs]     x (expr_offset, expr_len, n_args),
        exports[n_exports] x (slot, ref_ordinal, elem),
        instrs[n_instr]    x (expr_id, a_space, a_idx, b_space, b_idx)
    """

    def __init__(self, token_buffer=-1):
        self.token_buffer = int(token_buffer)     # the tape entry's nominal `out` (a block has many outputs: its exports)
        self.spill_buffer = -1                    # BUF_TA that collects slots read by LATER blocks; sized by finalize()
        self.n_spill = 0
        self.expr_id, self.a_space, self.a_idx, self.b_space, self.b_idx = [], [], [], [], []
        self.exprs = []           # (ops int32[n,4], fconsts float64[], n_args)
        self._expr_key = {}
        self.refbufs, self._ref_ord = [], {}
        self.literals, self._lit_ord = [], {}     # block-local literal pool (re-based into fconst on the wire)
        self.exports = []                         # (slot, ref_ordinal, elem)
        self.export_of = {}                       # slot -> (buffer, elem) of its first export

    def __len__(self):
        return len(self.expr_id)

    def expr(self, ops, fconsts, n_args) -> int:
        ops = np.ascontiguousarray(ops, np.int32).reshape(-1, 4)
        fconsts = np.ascontiguousarray(fconsts, np.float64)
        key = (ops.tobytes(), fconsts.tobytes(), int(n_args))
        k = self._expr_key.get(key)
        if k is None:
            k = self._expr_key[key] = len(self.exprs)
            self.exprs.append((ops, fconsts, int(n_args)))
        return k

    def ref(self, buffer: int) -> int:
        k = self._ref_ord.get(buffer)
        if k is None:
            k = self._ref_ord[buffer] = len(self.refbufs)
            self.refbufs.append(int(buffer))
        return k

    def literal(self, v: float) -> int:
        v = float(v)
        key = np.float64(v).tobytes()        # bit pattern: keeps -0.0 and NaN payloads apart
        k = self._lit_ord.get(key)
        if k is None:
            k = self._lit_ord[key] = len(self.literals)
            self.literals.append(v)
        return k

    def export(self, slot: int, buffer: int, elem: int):
        self.exports.append((int(slot), self.ref(buffer), int(elem)))
        self.export_of.setdefault(int(slot), (int(buffer), int(elem)))

    def push(self, expr_id, a, b) -> int:
        self.expr_id.append(expr_id)
        self.a_space.append(a[0]); self.a_idx.append(a[1])
        self.b_space.append(b[0]); self.b_idx.append(b[1])
        return len(self.expr_id) - 1

    @staticmethod
    def from_payload(w, expr, fconst, token_buffer) -> "ScalarBlock":
        """Inverse of the wire form; literals come back as indices into a block-local pool again."""
        w = np.asarray(w, np.int64)
        n, nref, nexp, nex = (int(v) for v in w[:4])
        blk = ScalarBlock(token_buffer)
        o = 4
        blk.refbufs = [int(v) for v in w[o:o + nref]]; o += nref
        blk._ref_ord = {b: k for k, b in enumerate(blk.refbufs)}
        table = w[o:o + 3 * nex].reshape(-1, 3); o += 3 * nex
        for eo, el, na in table:
            ops = np.asarray(expr[int(eo): int(eo) + 4 * int(el)], np.int32).reshape(-1, 4).copy()
            local = []
            for row in ops:
                if row[0] == E_CONST:
                    local.append(float(fconst[row[3]]))
                    row[3] = len(local) - 1
            blk.exprs.append((ops, np.asarray(local, np.float64), int(na)))
        for s_, r_, e_ in w[o:o + 3 * nexp].reshape(-1, 3):
            blk.exports.append((int(s_), int(r_), int(e_)))
            blk.export_of.setdefault(int(s_), (blk.refbufs[int(r_)], int(e_)))
        o += 3 * nexp
        body = w[o:o + 5 * n].reshape(-1, 5).copy()
        for sp, ix in ((1, 2), (3, 4)):
            for r in np.nonzero(body[:, sp] == SP_LITERAL)[0]:
                body[r, ix] = blk.literal(float(fconst[body[r, ix]]))
        blk.expr_id = body[:, 0].tolist()
        blk.a_space, blk.a_idx = body[:, 1].tolist(), body[:, 2].tolist()
        blk.b_space, blk.b_idx = body[:, 3].tolist(), body[:, 4].tolist()
        return blk


class Buffer:
    __slots__ = ("kind", "dims", "alias_of", "const_data")

    def __init__(self, kind, dims, alias_of=-1, const_data=None):
        self.kind = int(kind)
        self.dims = tuple(int(d) for d in dims)
        self.alias_of = int(alias_of)
        self.const_data = const_data

    @property
    def size(self):
        n = 1
        for d in self.dims:
            n *= d
        return n


class TapeProgram:
    """In-memory form; ``to_bytes``/``from_bytes`` give the wire form."""

    def __init__(self, dtype=F64):
        self.dtype = dtype
        self.buffers: list[Buffer] = []
        self.inputs: list[int] = []
        self.instrs: list[Instr] = []
        self.output_buffer = -1

    # ------------------------------------------------------------------------------ build
    def add_buffer(self, kind, dims, alias_of=-1, const_data=None) -> int:
        if len(dims) > MAX_DIMS:
            raise ValueError(f"arrays of more than {MAX_DIMS} dimensions are not supported")
        self.buffers.append(Buffer(kind, dims, alias_of, const_data))
        return len(self.buffers) - 1

    # -------------------------------------------------------------------------- serialise
    def finalize(self):
        """Fix the sizes that are only known once recording is over: a scalar block's spill buffer has one element per
        slot read by a later block, and such reads keep being added while recording goes on."""
        for it in self.instrs:
            if it.scalar is not None and it.scalar.spill_buffer >= 0:
                self.buffers[it.scalar.spill_buffer].dims = (max(1, it.scalar.n_spill),)

    def to_bytes(self) -> bytes:
        self.finalize()
        npdt = NP_DTYPE[self.dtype]
        nb, ni, nin = len(self.buffers), len(self.instrs), len(self.inputs)
        bufs = np.zeros(nb, BUFFER_DTYPE)
        const_chunks, const_off = [], 0
        for i, b in enumerate(self.buffers):
            bufs[i]["kind"] = b.kind
            bufs[i]["ndims"] = len(b.dims)
            bufs[i]["dims"][: len(b.dims)] = b.dims
            bufs[i]["dims"][len(b.dims):] = 1
            bufs[i]["alias_of"] = b.alias_of
            bufs[i]["const_offset"] = -1
            if b.kind == BUF_CONST:
                data = np.asarray(b.const_data, dtype=npdt).reshape(-1, order="F")
                raw = data.tobytes()
                bufs[i]["const_offset"] = const_off
                const_chunks.append(raw)
                pad = _align8(len(raw)) - len(raw)
                if pad:
                    const_chunks.append(b"\0" * pad)
                const_off += len(raw) + pad
        n_rec = sum(1 + (max(len(it.inputs), 1) - 1) // WIRE_ARGS for it in self.instrs)
        ins = np.zeros(n_rec, INSTR_DTYPE)
        idx_chunks, idx_off = [], 0
        expr_chunks, expr_off = [], 0
        fconst_chunks, fconst_off = [], 0
        r = 0
        for it in self.instrs:
            rec = ins[r]
            r0 = r
            r += 1 + (max(len(it.inputs), 1) - 1) // WIRE_ARGS
            rec["opcode"] = it.opcode
            rec["n_in"] = len(it.inputs)
            rec["out"] = it.out
            rec["n_expr_args"] = it.n_expr_args
            if len(it.inputs) > MAX_ARGS:
                raise ValueError(f"instructions with more than {MAX_ARGS} inputs are not supported")
            if it.scalar is not None:
                blk = it.scalar
                table = np.zeros((len(blk.exprs), 3), np.int64)
                for k, (ops, fc, na) in enumerate(blk.exprs):
                    ops = ops.copy()
                    for row in ops:
                        if row[0] == E_CONST:
                            row[3] += fconst_off
                    table[k] = (expr_off, ops.shape[0], na)
                    expr_chunks.append(ops.reshape(-1))
                    expr_off += ops.shape[0] * 4
                    fconst_chunks.append(fc)
                    fconst_off += len(fc)
                lit_base = fconst_off
                fconst_chunks.append(np.asarray(blk.literals, np.float64))
                fconst_off += len(blk.literals)
                n = len(blk)
                body = np.empty((n, 5), np.int64)
                body[:, 0] = blk.expr_id
                body[:, 1], body[:, 2] = blk.a_space, blk.a_idx
                body[:, 3], body[:, 4] = blk.b_space, blk.b_idx
                for sp, ix in ((1, 2), (3, 4)):
                    lit = body[:, sp] == SP_LITERAL
                    body[lit, ix] += lit_base
                payload = np.concatenate([
                    np.asarray([n, len(blk.refbufs), len(blk.exports), len(blk.exprs)], np.int64),
                    np.asarray(blk.refbufs, np.int64), table.reshape(-1),
                    np.asarray(blk.exports, np.int64).reshape(-1), body.reshape(-1)])
                it.inputs = [Operand(it.out, CLS_TR, False, (), IDX_EXPLICIT, payload)]
                rec["n_in"] = 1
            if it.expr is not None:
                ops, fc = it.expr
                ops = np.asarray(ops, np.int32).reshape(-1, 4).copy()
                # literal-constant indices are made absolute in the program-wide pool
                for row in ops:
                    if row[0] == E_CONST:
                        row[3] += fconst_off
                rec["expr_offset"] = expr_off
                rec["expr_len"] = ops.shape[0]
                expr_chunks.append(ops.reshape(-1))
                expr_off += ops.shape[0] * 4
                fconst_chunks.append(np.asarray(fc, np.float64))
                fconst_off += len(fc)
            for j, op in enumerate(it.inputs):
                if j >= WIRE_ARGS and j % WIRE_ARGS == 0:            # continuation record
                    more = ins[r0 + j // WIRE_ARGS]
                    more["opcode"], more["out"] = OP_MORE_OPERANDS, -1
                    more["n_in"] = min(WIRE_ARGS, len(it.inputs) - j)
                o = ins[r0 + j // WIRE_ARGS]["in"][j % WIRE_ARGS]
                o["buffer"], o["cls"], o["tracked"] = op.buffer, op.cls, int(op.tracked)
                o["idx_kind"] = op.idx_kind
                o["ndims"] = len(op.dims)
                o["dims"][: len(op.dims)] = op.dims
                o["dims"][len(op.dims):] = 1
                if op.bound is not None:
                    o["bound"][: len(op.bound)] = op.bound
                    o["bound"][len(op.bound):] = 1
                else:
                    o["bound"][:] = 0
                if op.idx_kind == IDX_EXPLICIT:
                    m = np.ascontiguousarray(op.idx_map, dtype=np.int64).reshape(-1)
                    o["idx_offset"], o["idx_len"] = idx_off, m.size
                    idx_chunks.append(m)
                    idx_off += m.size
                else:
                    o["idx_offset"], o["idx_len"] = -1, 0
        idx = np.concatenate(idx_chunks) if idx_chunks else np.zeros(0, np.int64)
        expr = np.concatenate(expr_chunks) if expr_chunks else np.zeros(0, np.int32)
        fconst = np.concatenate(fconst_chunks) if fconst_chunks else np.zeros(0, np.float64)
        consts = b"".join(const_chunks)
        inputs = np.asarray(self.inputs, np.int32)

        hdr = np.zeros(1, HEADER_DTYPE)
        off = HEADER_DTYPE.itemsize
        sections = []

        def place(raw: bytes):
            nonlocal off
            start = off
            sections.append(raw)
            pad = _align8(len(raw)) - len(raw)
            if pad:
                sections.append(b"\0" * pad)
            off += len(raw) + pad
            return start

        h = hdr[0]
        h["magic"], h["version"], h["dtype"] = MAGIC, VERSION, self.dtype
        h["n_buffers"], h["n_instr"], h["n_inputs"] = nb, n_rec, nin
        h["output_buffer"] = self.output_buffer
        h["off_buffers"] = place(bufs.tobytes())
        h["off_inputs"] = place(inputs.tobytes())
        h["off_instr"] = place(ins.tobytes())
        h["off_idx"], h["n_idx"] = place(idx.tobytes()), idx.size
        h["off_expr"], h["n_expr"] = place(expr.tobytes()), expr.size
        h["off_fconst"], h["n_fconst"] = place(fconst.tobytes()), fconst.size
        h["off_const"], h["n_const"] = place(consts), len(consts)
        h["total_bytes"] = off
        return hdr.tobytes() + b"".join(sections)

    # ------------------------------------------------------------------------ deserialise
    @staticmethod
    def from_bytes(raw: bytes) -> "TapeProgram":
        mv = memoryview(raw)
        h = np.frombuffer(mv, HEADER_DTYPE, 1)[0]
        if int(h["magic"]) != MAGIC:
            raise ValueError("not a tape program (bad magic)")
        if int(h["version"]) != VERSION:
            raise ValueError(f"tape program version {int(h['version'])} != {VERSION}")
        if int(h["total_bytes"]) != len(raw):
            raise ValueError("tape program is truncated")
        p = TapeProgram(int(h["dtype"]))
        npdt = NP_DTYPE[p.dtype]
        es = np.dtype(npdt).itemsize
        nb, ni, nin = int(h["n_buffers"]), int(h["n_instr"]), int(h["n_inputs"])
        bufs = np.frombuffer(mv, BUFFER_DTYPE, nb, int(h["off_buffers"]))
        inputs = np.frombuffer(mv, np.int32, nin, int(h["off_inputs"]))
        ins = np.frombuffer(mv, INSTR_DTYPE, ni, int(h["off_instr"]))
        idx = np.frombuffer(mv, np.int64, int(h["n_idx"]), int(h["off_idx"]))
        expr = np.frombuffer(mv, np.int32, int(h["n_expr"]), int(h["off_expr"]))
        fconst = np.frombuffer(mv, np.float64, int(h["n_fconst"]), int(h["off_fconst"]))
        coff = int(h["off_const"])
        for b in bufs:
            nd = int(b["ndims"])
            dims = tuple(int(d) for d in b["dims"][:nd])
            data = None
            if int(b["kind"]) == BUF_CONST:
                n = int(np.prod(dims)) if dims else 1
                data = np.frombuffer(mv, npdt, n, coff + int(b["const_offset"])).reshape(dims, order="F")
            p.buffers.append(Buffer(int(b["kind"]), dims, int(b["alias_of"]), data))
        p.inputs = [int(i) for i in inputs]
        p.output_buffer = int(h["output_buffer"])
        for r, rec in enumerate(ins):
            if int(rec["opcode"]) == OP_MORE_OPERANDS:
                continue
            ops_in = []
            for j in range(int(rec["n_in"])):
                o = ins[r + j // WIRE_ARGS]["in"][j % WIRE_ARGS]
                nd = int(o["ndims"])
                m = None
                if int(o["idx_kind"]) == IDX_EXPLICIT:
                    m = idx[int(o["idx_offset"]): int(o["idx_offset"]) + int(o["idx_len"])]
                bound = tuple(int(x) for x in o["bound"][:MAX_DIMS])
                ops_in.append(Operand(int(o["buffer"]), int(o["cls"]), bool(o["tracked"]),
                                      tuple(int(d) for d in o["dims"][:nd]), int(o["idx_kind"]), m,
                                      bound if any(bound) else None))
            ex = None
            if int(rec["expr_len"]) > 0:
                eo, el = int(rec["expr_offset"]), int(rec["expr_len"])
                ops = expr[eo: eo + 4 * el].reshape(-1, 4).copy()
                local = []
                for row in ops:          # re-base literal indices to an instruction-local pool
                    if row[0] == E_CONST:
                        local.append(float(fconst[row[3]]))
                        row[3] = len(local) - 1
                ex = (ops, np.asarray(local, np.float64))
            blk = None
            if int(rec["opcode"]) == OP_SCALAR_BLOCK:
                blk = ScalarBlock.from_payload(ops_in[0].idx_map, expr, fconst, int(rec["out"]))
            p.instrs.append(Instr(int(rec["opcode"]), ops_in, int(rec["out"]), ex, int(rec["n_expr_args"]), blk))
        _ = es
        return p


def transpose_map(rows: int, cols: int) -> np.ndarray:
    """Index map of ``adjoint(A)`` for a ``rows x cols`` column-major parent.

    Element ``(i, j)`` of the ``cols x rows`` adjoint is ``A[j, i]``, whose linear index in
    the parent is ``j + i*rows`` (0-based) — i.e. ``TrackedReal.index`` as set by scalar
    ``getindex`` (reference ``src/tracked.jl:348-351``) reached through ``Adjoint`` indexing.
    The adjoint's own column-major position is ``k = i + j*cols``.
    """
    i = np.arange(cols, dtype=np.int64)[:, None]   # row of the adjoint
    j = np.arange(rows, dtype=np.int64)[None, :]   # column of the adjoint
    return (j + i * rows).reshape(-1, order="F")
