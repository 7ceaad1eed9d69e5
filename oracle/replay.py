"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

CPU restatement (numpy) of ReverseDiff's compiled-tape replay semantics, instruction by
instruction, for the opcode subset of the tape program.  All citations are relative to
/root/reference (JuliaDiff/ReverseDiff.jl v1.16.2).

PARITY PINNING.  The reference is pure Julia and no `julia` binary exists in this image, so
the reference cannot be executed here; it also stores no golden vectors for the five
benchmark configs ("parity unpinned" for those sizes).  The restatement is pinned instead on
(i) every in-scope known-answer literal of the reference's tests (LinAlgTests.jl:281-282,
compat/CompatTests.jl:5-11, the index iterables of TrackedTests.jl:641-716), committed under
tests/golden/reference_literals.json, (ii) independent closed forms and
(iii) central finite differences — see tests/test_oracle.py.

Third-party arithmetic restated here (not under /root/reference; compat ranges only,
Project.toml:20-36): ForwardDiff "0.10, 1" Dual arithmetic with DiffRules "1.4" derivative
rules (tanh' = 1 - tanh², exp' = exp, log' = 1/x, sqrt' = 1/(2√x), abs2' = 2x, (a/b)' =
(1/b, -a/b²), x^n → n·x^(n-1), …), DiffResults value+partials records, and
LinearAlgebra.mul! → OpenBLAS gemm (here numpy → OpenBLAS).

The file parses the wire format on its own (it does not import the product package), so a
serialisation bug cannot cancel out between product and checker.
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x3030324244525052
VERSION = 2
MAX_ARGS, MAX_DIMS = 4, 4          # (MAX_ARGS: operand slots per record; longer operand lists continue in OP_MORE_OPERANDS records)
OP_MORE_OPERANDS = 22

BUF_TA, BUF_TR, BUF_CONST = 0, 1, 2
CLS_TA, CLS_IDX, CLS_CONST, CLS_TR = 0, 1, 2, 3
IDX_NONE, IDX_EXPLICIT, IDX_TRANSPOSE = 0, 1, 2
(OP_MUL, OP_ADD, OP_SUB, OP_NEG, OP_SUM, OP_MEAN, OP_NBCAST, OP_MAP, OP_BCAST,
 OP_BCAST_ADD, OP_BCAST_SUB, OP_BCAST_MUL) = range(1, 13)
OP_BCAST_DIV, OP_BCAST_LDIV, OP_BCAST_POW, OP_GETINDEX, OP_DOT, OP_SUMDIMS, OP_TRACKER_BCAST, OP_GETINDEX_GENERIC = range(13, 21)
OP_SCALAR_BLOCK = 21
SP_POOL, SP_LITERAL, SP_NONE = -1, -2, -3
ELEMENTWISE = (OP_NBCAST, OP_MAP, OP_BCAST, OP_BCAST_ADD, OP_BCAST_SUB, OP_BCAST_MUL, OP_BCAST_DIV, OP_BCAST_LDIV, OP_BCAST_POW,
               OP_TRACKER_BCAST)
_LN2, _LN10 = 0.6931471805599453, 2.302585092994046      # Python floats: weak in NumPy's promotion, the tape's dtype is kept
(E_CONST, E_ADD, E_SUB, E_MUL, E_DIV, E_NEG, E_POWI, E_POW, E_TANH, E_EXP, E_LOG, E_SQRT,
 E_SIN, E_COS, E_ABS2, E_ABS, E_MAX, E_MIN, E_INV, E_SIGMOID, E_ARG,
 E_LOG1P, E_EXPM1, E_ASIN, E_ACOS, E_ATAN, E_SINH, E_COSH, E_TAN, E_ATAN2, E_HYPOT,
 E_EXP2, E_EXP10, E_LOG2, E_LOG10, E_CBRT, E_ASINH, E_ACOSH, E_ATANH, E_SEC, E_CSC, E_COT, E_SECH, E_CSCH, E_COTH) = range(1, 46)

_HEADER = np.dtype([
    ("magic", "<u8"), ("version", "<u4"), ("dtype", "<u4"),
    ("n_buffers", "<u4"), ("n_instr", "<u4"), ("n_inputs", "<u4"), ("output_buffer", "<i4"),
    ("off_buffers", "<u8"), ("off_inputs", "<u8"), ("off_instr", "<u8"),
    ("off_idx", "<u8"), ("n_idx", "<u8"), ("off_expr", "<u8"), ("n_expr", "<u8"),
    ("off_fconst", "<u8"), ("n_fconst", "<u8"), ("off_const", "<u8"), ("n_const", "<u8"),
    ("total_bytes", "<u8")])
_BUFFER = np.dtype([("kind", "<i4"), ("ndims", "<i4"), ("dims", "<i8", (MAX_DIMS,)),
                    ("alias_of", "<i4"), ("pad", "<i4"), ("const_offset", "<i8")])
_OPERAND = np.dtype([("buffer", "<i4"), ("cls", "<i4"), ("tracked", "<i4"), ("idx_kind", "<i4"),
                     ("idx_offset", "<i8"), ("idx_len", "<i8"), ("ndims", "<i4"), ("pad", "<i4"),
                     ("dims", "<i8", (MAX_DIMS,)), ("bound", "<i8", (MAX_DIMS,))])
_INSTR = np.dtype([("opcode", "<i4"), ("n_in", "<i4"), ("out", "<i4"), ("n_expr_args", "<i4"),
                   ("expr_offset", "<i8"), ("expr_len", "<i8"), ("in", _OPERAND, (MAX_ARGS,))])


class _Operand:
    pass


class _Instr:
    pass


# ======================================================================================
# ForwardDiff-style Dual evaluation of the scalar bytecode (vectorised over elements)
# ======================================================================================
def _powi(x, n):
    if n == 0:
        return np.ones_like(x)
    if n < 0:
        return 1.0 / _powi(x, -n)
    r = x
    for _ in range(n - 1):
        r = r * x
    return r


def dual_eval(ops, fconst, args, dtype, order=1):
    """One ``Dual{N}`` evaluation per element: value + N partials (``broadcast.jl:147-153``,
    ``elementwise.jl:234,301``); with ``order=2`` also the N×N second partials (hyper-dual),
    used only for the Hessian check.

    Returns ``(value, [partial_p], [[second_pq]])`` with every array broadcast to the common shape.
    """
    n = len(args)
    zero = np.zeros((), dtype)
    one = np.ones((), dtype)
    V = [np.asarray(a, dtype) for a in args]
    D = [[one if p == q else zero for p in range(n)] for q in range(n)]   # D[reg][p]
    H = [[[zero] * n for _ in range(n)] for _ in range(n)] if order == 2 else None

    def unary(a, f0, f1, f2=None):
        """chain rule: y = f(u): y_p = f'(u) u_p ; y_pq = f''(u) u_p u_q + f'(u) u_pq"""
        V.append(f0)
        D.append([f1 * D[a][p] for p in range(n)])
        if order == 2:
            H.append([[f2 * D[a][p] * D[a][q] + f1 * H[a][p][q] for q in range(n)] for p in range(n)])

    for code, a, b, imm in ops:
        code, a, b, imm = int(code), int(a), int(b), int(imm)
        if code == E_CONST:
            V.append(np.asarray(fconst[imm], dtype)); D.append([zero] * n)
            if order == 2: H.append([[zero] * n for _ in range(n)])
        elif code == E_ARG:
            V.append(V[imm]); D.append(list(D[imm]))
            if order == 2: H.append([list(r) for r in H[imm]])
        elif code in (E_ADD, E_SUB):
            s = one if code == E_ADD else -one
            V.append(V[a] + V[b] if code == E_ADD else V[a] - V[b])
            D.append([D[a][p] + s * D[b][p] for p in range(n)])
            if order == 2: H.append([[H[a][p][q] + s * H[b][p][q] for q in range(n)] for p in range(n)])
        elif code == E_MUL:
            V.append(V[a] * V[b])
            D.append([D[a][p] * V[b] + V[a] * D[b][p] for p in range(n)])
            if order == 2:
                H.append([[H[a][p][q] * V[b] + D[a][p] * D[b][q] + D[a][q] * D[b][p] + V[a] * H[b][p][q]
                           for q in range(n)] for p in range(n)])
        elif code == E_DIV:
            # (a/b)' = (1/b, -a/b^2)
            q0 = V[a] / V[b]
            ib = one / V[b]
            V.append(q0)
            da = [(D[a][p] - q0 * D[b][p]) * ib for p in range(n)]
            D.append(da)
            if order == 2:
                # y = a/b  =>  y_pq = (a_pq - y_p b_q - y_q b_p - y b_pq)/b
                H.append([[(H[a][p][q] - da[p] * D[b][q] - da[q] * D[b][p] - q0 * H[b][p][q]) * ib
                           for q in range(n)] for p in range(n)])
        elif code == E_NEG:
            unary(a, -V[a], -one, zero)
        elif code == E_POWI:
            x = V[a]
            f1 = imm * _powi(x, imm - 1) if imm != 0 else zero * x
            f2 = imm * (imm - 1) * _powi(x, imm - 2) if imm not in (0, 1) else zero * x
            unary(a, _powi(x, imm), f1, f2)
        elif code == E_POW:
            # ForwardDiff dual.jl, "exponentiation" (compat "0.10, 1"): three methods, selected by which operands are Duals.
            # imm (recorded by the host: expr.py pow_imm): bits 0-1 flavor, bit 2 base is a plain Real, bit 3 exponent is.
            #   flavor 0: one Dual{N} evaluation (∇broadcast broadcast.jl:147-153; @forward closures elementwise.jl:234-301;
            #             scalar instructions macros.jl:87-124): isconstant(y) / iszero(partials(x)) look at the whole vector
            #   flavor 1: tracker_∇broadcast, one Dual{1} evaluation per argument (broadcast.jl:218-224): per component
            #   flavor 2: (broadcast,^) caches e*b^(e-1) and log(b)*b^e directly (elementwise.jl:633-643)
            vx, vy = V[a], V[b]
            flavor, x_real, y_real = imm & 3, bool(imm & 4), bool(imm & 8)
            with np.errstate(all="ignore"):
                expv = vx ** vy
                if x_real and y_real:
                    d = [zero * expv for _ in range(n)]
                elif flavor == 2:
                    bs = vy * vx ** (vy - one)                 # base_partials_kernel
                    ex = np.log(vx) * expv                     # exp_partials_kernel
                    d = [np.where(D[a][p] != 0, bs * D[a][p], zero) + np.where(D[b][p] != 0, ex * D[b][p], zero) for p in range(n)]
                elif y_real:
                    # x::Dual ^ y::Real:  (y == 0 || iszero(partials(x))) ? zero : partials(x) * y * x^(y-1)
                    pw = vx ** (vy - one)
                    allz = np.logical_and.reduce([np.asarray(D[a][p] == 0) for p in range(n)])
                    d = [np.where((vy == 0) | (np.asarray(D[a][p] == 0) if flavor == 1 else allz), zero, D[a][p] * vy * pw)
                         for p in range(n)]
                elif x_real:
                    # x::Real ^ y::Dual:  deriv = (iszero(x) && v > 0) ? zero(expv) : expv*log(x)
                    deriv = np.where((vx == 0) & (vy > 0), zero, expv * np.log(vx))
                    d = [deriv * D[b][p] for p in range(n)]
                else:
                    powval = vy * vx ** (vy - one)
                    lg = np.where((vx == 0) & (vy > 0), zero, expv * np.log(vx))
                    cst = np.logical_and.reduce([np.asarray(D[b][p] == 0) for p in range(n)])
                    d = [D[a][p] * powval + D[b][p] * np.where(np.asarray(D[b][p] == 0) if flavor == 1 else cst, one, lg)
                         for p in range(n)]
                V.append(expv); D.append(d)
                if order == 2:
                    # second partials of x^y (not ForwardDiff's code path - the reference nests tapes for Hessians; this is
                    # calculus): f_xx = y(y-1)x^(y-2), f_xy = x^(y-1)(1 + y log x), f_yy = x^y log(x)^2
                    a_off = x_real | (y_real & np.asarray(vy == 0))
                    lx = np.log(vx)
                    fa, fb = vy * vx ** (vy - one), np.where((vx == 0) & (vy > 0), zero, expv * lx)
                    faa, fab, fbb = vy * (vy - one) * vx ** (vy - 2 * one), vx ** (vy - one) * (one + vy * lx), fb * lx
                    Hn = []
                    for p in range(n):
                        row = []
                        for q in range(n):
                            h = zero * expv
                            if not x_real:
                                h = np.where(a_off, h, fa * H[a][p][q] + faa * D[a][p] * D[a][q])
                            if not y_real:
                                h = h + fb * H[b][p][q] + fbb * D[b][p] * D[b][q]
                            if not x_real and not y_real:
                                h = h + fab * (D[a][p] * D[b][q] + D[a][q] * D[b][p])
                            row.append(h)
                        Hn.append(row)
                    H.append(Hn)
        elif code == E_TANH:
            t = np.tanh(V[a]); f1 = one - t * t
            unary(a, t, f1, -2 * t * f1)
        elif code == E_EXP:
            e = np.exp(V[a]); unary(a, e, e, e)
        elif code == E_LOG:
            unary(a, np.log(V[a]), one / V[a], -one / (V[a] * V[a]))
        elif code == E_SQRT:
            s = np.sqrt(V[a]); unary(a, s, one / (2 * s), -one / (4 * s * V[a]))
        elif code == E_SIN:
            unary(a, np.sin(V[a]), np.cos(V[a]), -np.sin(V[a]))
        elif code == E_COS:
            unary(a, np.cos(V[a]), -np.sin(V[a]), -np.cos(V[a]))
        elif code == E_ABS2:
            unary(a, V[a] * V[a], 2 * V[a], 2 * one)
        elif code == E_ABS:
            # ForwardDiff: abs(d::Dual) = signbit(d) ? -d : d   (derivative +1 at +0.0, -1 at -0.0)
            unary(a, np.abs(V[a]), np.where(np.signbit(V[a]), -one, one), zero)
        elif code in (E_MAX, E_MIN):
            pick_a = (V[a] > V[b]) if code == E_MAX else (V[a] < V[b])
            V.append(np.where(pick_a, V[a], V[b]))
            D.append([np.where(pick_a, D[a][p], D[b][p]) for p in range(n)])
            if order == 2:
                H.append([[np.where(pick_a, H[a][p][q], H[b][p][q]) for q in range(n)] for p in range(n)])
        elif code == E_INV:
            r = one / V[a]; unary(a, r, -r * r, 2 * r * r * r)
        elif code == E_SIGMOID:
            s = one / (one + np.exp(-V[a])); f1 = s * (one - s)
            unary(a, s, f1, f1 * (one - 2 * s))
        # DiffRules (v1.x, `@define_diffrule`): log1p' = inv(x + 1); expm1' = exp(x); asin' = inv(sqrt(1 - x^2)); acos' = -inv(sqrt(1 - x^2));
        # atan' = inv(1 + x^2); sinh' = cosh(x); cosh' = sinh(x); tan' = 1 + tan(x)^2  (second derivatives: calculus)
        elif code == E_LOG1P:
            u = one / (V[a] + one); unary(a, np.log1p(V[a]), u, -u * u)
        elif code == E_EXPM1:
            e = np.exp(V[a]); unary(a, np.expm1(V[a]), e, e)
        elif code == E_ASIN:
            w = one - V[a] * V[a]; u = one / np.sqrt(w); unary(a, np.arcsin(V[a]), u, V[a] * u / w)
        elif code == E_ACOS:
            w = one - V[a] * V[a]; u = one / np.sqrt(w); unary(a, np.arccos(V[a]), -u, -(V[a] * u / w))
        elif code == E_ATAN:
            u = one / (one + V[a] * V[a]); unary(a, np.arctan(V[a]), u, -2 * V[a] * u * u)
        elif code == E_SINH:
            unary(a, np.sinh(V[a]), np.cosh(V[a]), np.sinh(V[a]))
        elif code == E_COSH:
            unary(a, np.cosh(V[a]), np.sinh(V[a]), np.cosh(V[a]))
        elif code == E_TAN:
            tt = np.tan(V[a]); f1 = one + tt * tt; unary(a, tt, f1, 2 * tt * f1)
        # third tranche (DiffRules `@define_diffrule Base.*`): exp2' = exp2(x) * logtwo; exp10' = exp10(x) * logten; log2' = inv(x) / logtwo;
        # log10' = inv(x) / logten; cbrt' = inv(3 * cbrt(x)^2); asinh' = inv(sqrt(x^2 + 1)); acosh' = inv(sqrt(x^2 - 1));
        # atanh' = inv(1 - x^2); sec' = sec(x) * tan(x); csc' = -csc(x) * cot(x); cot' = -(1 + cot(x)^2); sech' = -tanh(x) * sech(x);
        # csch' = -coth(x) * csch(x); coth' = -(csch(x)^2)   (second derivatives: calculus; sec = inv(cos) etc. as Base defines them)
        elif code == E_EXP2:
            f0 = np.exp2(V[a]); f1 = f0 * _LN2; unary(a, f0, f1, f1 * _LN2)
        elif code == E_EXP10:
            f0 = np.power(10.0, V[a]); f1 = f0 * _LN10; unary(a, f0, f1, f1 * _LN10)
        elif code == E_LOG2:
            u = one / V[a]; f1 = u / _LN2; unary(a, np.log2(V[a]), f1, -(f1 * u))
        elif code == E_LOG10:
            u = one / V[a]; f1 = u / _LN10; unary(a, np.log10(V[a]), f1, -(f1 * u))
        elif code == E_CBRT:
            c = np.cbrt(V[a]); f1 = one / (3 * (c * c)); unary(a, c, f1, -2 * f1 / (3 * V[a]))
        elif code == E_ASINH:
            u = one / np.sqrt(V[a] * V[a] + one); unary(a, np.arcsinh(V[a]), u, -(V[a] * u * u * u))
        elif code == E_ACOSH:
            u = one / np.sqrt(V[a] * V[a] - one); unary(a, np.arccosh(V[a]), u, -(V[a] * u * u * u))
        elif code == E_ATANH:
            u = one / (one - V[a] * V[a]); unary(a, np.arctanh(V[a]), u, 2 * V[a] * u * u)
        elif code == E_SEC:
            sc = one / np.cos(V[a]); tt = np.tan(V[a]); unary(a, sc, sc * tt, sc * (tt * tt + sc * sc))
        elif code == E_CSC:
            cs = one / np.sin(V[a]); ct = one / np.tan(V[a]); unary(a, cs, -(cs * ct), cs * (ct * ct + cs * cs))
        elif code == E_COT:
            ct = one / np.tan(V[a]); f1 = -(one + ct * ct); unary(a, ct, f1, -2 * ct * f1)
        elif code == E_SECH:
            sh = one / np.cosh(V[a]); th = np.tanh(V[a]); unary(a, sh, -(th * sh), sh * (th * th - sh * sh))
        elif code == E_CSCH:
            ch = one / np.sinh(V[a]); ct = one / np.tanh(V[a]); unary(a, ch, -(ct * ch), ch * (ct * ct + ch * ch))
        elif code == E_COTH:
            ct = one / np.tanh(V[a]); ch = one / np.sinh(V[a]); f1 = -(ch * ch); unary(a, ct, f1, -2 * ct * f1)
        elif code in (E_ATAN2, E_HYPOT):
            # atan(y, x): (x / (y^2 + x^2), -y / (y^2 + x^2)); hypot(x, y): (x / hypot(x, y), y / hypot(x, y))
            va, vb = V[a], V[b]
            if code == E_ATAN2:
                r2 = va * va + vb * vb
                y = np.arctan2(va, vb); ga, gb = vb / r2, -va / r2
                i4 = one / (r2 * r2); gaa, gbb, gab = -2 * va * vb * i4, 2 * va * vb * i4, (va * va - vb * vb) * i4
            else:
                y = np.hypot(va, vb); ga, gb = va / y, vb / y
                i3 = one / (y * y * y); gaa, gbb, gab = vb * vb * i3, va * va * i3, -(va * vb) * i3
            V.append(y)
            D.append([ga * D[a][p] + gb * D[b][p] for p in range(n)])
            if order == 2:
                H.append([[ga * H[a][p][q] + gb * H[b][p][q] + gaa * D[a][p] * D[a][q] + gab * (D[a][p] * D[b][q] + D[a][q] * D[b][p])
                           + gbb * D[b][p] * D[b][q] for q in range(n)] for p in range(n)])
        else:
            raise ValueError(f"unknown scalar opcode {code}")
    shape = np.broadcast_shapes(*[np.shape(v) for v in V[:n]], np.shape(V[-1]))
    val = np.broadcast_to(np.asarray(V[-1], dtype), shape)
    part = [np.broadcast_to(np.asarray(D[-1][p], dtype), shape) for p in range(n)]
    sec = None
    if order == 2:
        sec = [[np.broadcast_to(np.asarray(H[-1][p][q], dtype), shape) for q in range(n)] for p in range(n)]
    return val, part, sec


# ======================================================================================
# Tape
# ======================================================================================
class OracleTape:
    """Buffers + instruction list, replayed like ``forward_pass!`` / ``reverse_pass!``
    (``src/tape.jl:75-90``; compiled variant ``src/api/tape.jl:122-134``)."""

    def __init__(self, program_bytes: bytes, keep_reference_copies: bool = True):
        mv = memoryview(program_bytes)
        h = np.frombuffer(mv, _HEADER, 1)[0]
        assert int(h["magic"]) == MAGIC and int(h["version"]) == VERSION, "bad tape program"
        assert int(h["total_bytes"]) == len(program_bytes), "truncated tape program"
        self.dtype = np.dtype(np.float64 if int(h["dtype"]) == 0 else np.float32)
        self.keep_reference_copies = keep_reference_copies
        bufs = np.frombuffer(mv, _BUFFER, int(h["n_buffers"]), int(h["off_buffers"]))
        self.inputs = [int(i) for i in np.frombuffer(mv, np.int32, int(h["n_inputs"]), int(h["off_inputs"]))]
        ins = np.frombuffer(mv, _INSTR, int(h["n_instr"]), int(h["off_instr"]))
        idx = np.frombuffer(mv, np.int64, int(h["n_idx"]), int(h["off_idx"]))
        expr = np.frombuffer(mv, np.int32, int(h["n_expr"]), int(h["off_expr"]))
        self.fconst = np.frombuffer(mv, np.float64, int(h["n_fconst"]), int(h["off_fconst"]))
        coff = int(h["off_const"])
        self.output = int(h["output_buffer"])
        self.kind, self.dims, self.value, self.deriv = [], [], [], []
        for b in bufs:
            nd = int(b["ndims"]); dims = tuple(int(d) for d in b["dims"][:nd])
            n = int(np.prod(dims)) if dims else 1
            kind = int(b["kind"])
            self.kind.append(kind); self.dims.append(dims)
            al = int(b["alias_of"])
            if al >= 0:   # reshape aliases value AND deriv (tracked.jl:402-408)
                self.value.append(self.value[al].reshape(dims, order="F"))
                self.deriv.append(self.deriv[al].reshape(dims, order="F"))
            elif kind == BUF_CONST:
                data = np.frombuffer(mv, self.dtype, n, coff + int(b["const_offset"]))
                self.value.append(np.array(data.reshape(dims, order="F"), order="F"))
                self.deriv.append(None)
            else:
                self.value.append(np.zeros(dims, self.dtype, order="F"))
                self.deriv.append(np.zeros(dims, self.dtype, order="F"))
        self.instrs = []
        for r_, rec in enumerate(ins):
            if int(rec["opcode"]) == OP_MORE_OPERANDS:       # operands 5.. of the record before it (broadcast.jl:142-146 is N-ary)
                continue
            it = _Instr()
            it.opcode, it.out = int(rec["opcode"]), int(rec["out"])
            it.n_expr_args = int(rec["n_expr_args"])
            it.inputs = []
            for j in range(int(rec["n_in"])):
                o = ins[r_ + j // MAX_ARGS]["in"][j % MAX_ARGS]
                op = _Operand()
                op.buffer, op.cls, op.tracked = int(o["buffer"]), int(o["cls"]), bool(o["tracked"])
                nd = int(o["ndims"])
                op.dims = tuple(int(d) for d in o["dims"][:nd])
                op.bound = tuple(int(x) for x in o["bound"])
                k = int(o["idx_kind"])
                op.unique = True
                if k == IDX_EXPLICIT:
                    op.idx = idx[int(o["idx_offset"]): int(o["idx_offset"]) + int(o["idx_len"])]
                    # duplicates must each add once (sequential semantics); without duplicates a vectorised += is the same thing
                    op.unique = np.unique(op.idx).size == op.idx.size
                elif k == IDX_TRANSPOSE:
                    # adjoint of a (rows x cols) parent: element (i,j) -> j + i*rows   (tracked.jl:348-351)
                    rows, cols = self.dims[op.buffer]
                    ii = np.arange(cols, dtype=np.int64)[:, None]; jj = np.arange(rows, dtype=np.int64)[None, :]
                    op.idx = (jj + ii * rows).reshape(-1, order="F")
                else:
                    op.idx = None
                it.inputs.append(op)
            it.expr = None
            if int(rec["expr_len"]) > 0:
                eo, el = int(rec["expr_offset"]), int(rec["expr_len"])
                it.expr = expr[eo: eo + 4 * el].reshape(-1, 4)
            it.cache = None
            if it.opcode == OP_SCALAR_BLOCK:
                it.block = _ScalarBlock(it.inputs[0].idx, expr)
            self.instrs.append(it)

    # ------------------------------------------------------------------ operand accessors
    def _val(self, op):
        """``value(x)`` after ``pull_value!`` (``tracked.jl:178-181``).  For IDX operands the reference
        re-gathers element by element and then densifies with ``map(value, a)`` (``tracked.jl:103``)."""
        if op.cls == CLS_IDX:
            flat = self.value[op.buffer].reshape(-1, order="F")
            g = flat[op.idx]                       # pull_value!
            if self.keep_reference_copies:
                g = g.copy()                       # value(::Array{TrackedReal}) dense copy
            return g.reshape(op.dims, order="F")
        v = self.value[op.buffer]
        return v if op.cls != CLS_TR else np.asarray(v).reshape(())

    def _inc(self, op, x, sign=1.0):
        """``increment_deriv!`` / ``decrement_deriv!`` (``propagation.jl:33-73``)."""
        if not op.tracked:
            return
        d = self.deriv[op.buffer]
        if op.cls == CLS_IDX:
            # per element: pull_deriv!, add, push_deriv! (propagation.jl:45-50) — sequential, duplicates add
            flat = d.reshape(-1, order="F")
            upd = sign * np.asarray(x, self.dtype).reshape(-1, order="F")
            if op.unique:
                flat[op.idx] += upd
            else:
                np.add.at(flat, op.idx, upd)
            self.deriv[op.buffer][...] = flat.reshape(d.shape, order="F")
        elif op.cls == CLS_TR:
            d[...] = d + sign * x
        else:
            if sign > 0:
                d += np.asarray(x, self.dtype).reshape(d.shape, order="F")
            else:
                d -= np.asarray(x, self.dtype).reshape(d.shape, order="F")

    def _unseed(self, b):
        """``unseed!`` (``tracked.jl:207-213``)."""
        self.deriv[b][...] = 0

    # ----------------------------------------------------------------------- public API
    def set_input(self, slot, x):
        """``value!(input_hook(t)[slot], x)`` (``tracked.jl:156-163``)."""
        b = self.inputs[slot]
        self.value[b][...] = np.asarray(x, self.dtype).reshape(self.dims[b], order="F")

    def forward(self):
        """``forward_pass!`` in recorded order (``src/tape.jl:75-80``)."""
        for it in self.instrs:
            self._forward_exec(it)

    def reverse(self):
        """``reverse_pass!`` in reverse order (``src/tape.jl:85-90``)."""
        for it in reversed(self.instrs):
            self._reverse_exec(it)

    def gradient(self, inputs):
        """``gradient!(result, tape, input)`` (``src/api/gradients.jl:78-82``):
        seeded_forward_pass! (``api/tape.jl:40-44``) + seeded_reverse_pass! (``api/utils.jl:27-34``)."""
        for i, x in enumerate(inputs):
            self.set_input(i, x)
        self.forward()
        for b in self.inputs:
            self._unseed(b)                       # unseed!(input)
        self.deriv[self.output][...] = 1          # seed!(output)
        self.reverse()
        val = np.asarray(self.value[self.output]).reshape(()).copy()
        return val, [self.deriv[b].copy(order="F") for b in self.inputs]   # extract_result!

    def jacobian(self, inputs, slot=0, rows=None):
        """``jacobian!`` (``src/api/jacobians.jl:120-124``): one forward sweep then one one-hot seeded
        reverse sweep per output element (``src/api/utils.jl:44-58``).  Result is column-major
        ``length(output) x length(input)``."""
        for i, x in enumerate(inputs):
            self.set_input(i, x)
        self.forward()
        out_d = self.deriv[self.output]
        m = out_d.size
        xin = self.inputs[slot]
        n = self.deriv[xin].size
        rows = range(m) if rows is None else rows
        J = np.zeros((len(rows), n), self.dtype, order="F")
        of = out_d.reshape(-1, order="F")
        for r, i in enumerate(rows):
            for b in self.inputs:
                self._unseed(b)
            of[i] = 1                              # seed!(output, i)
            self.deriv[self.output][...] = of.reshape(out_d.shape, order="F")
            self.reverse()
            J[r, :] = self.deriv[xin].reshape(-1, order="F")
            of = self.deriv[self.output].reshape(-1, order="F")
            of[i] = 0                              # unseed!(output, i)
            self.deriv[self.output][...] = of.reshape(out_d.shape, order="F")
        return self.value[self.output].copy(order="F"), J

    # ------------------------------------------------------------------ instruction execs
    def _forward_exec(self, it):
        op = it.opcode
        out = self.value[it.out]
        a = it.inputs
        if op == OP_MUL:                      # arithmetic.jl:245-255
            va, vb = self._val(a[0]), self._val(a[1])
            res = va @ vb                     # mul! -> BLAS gemm / gemv
            out[...] = res.reshape(out.shape, order="F")
        elif op == OP_ADD:                    # arithmetic.jl:55-61, plus! :7-12
            out[...] = self._val(a[0]) + self._val(a[1])
        elif op == OP_SUB:                    # arithmetic.jl:142-154, minus! :69-79
            out[...] = self._val(a[0]) - self._val(a[1])
        elif op == OP_NEG:
            out[...] = -self._val(a[0])
        elif op == OP_SUM:                    # reductions.jl:23-27
            self.value[it.out][...] = np.sum(self._val(a[0]))
        elif op == OP_MEAN:                   # reductions.jl:81-85
            self.value[it.out][...] = np.mean(self._val(a[0]))
        elif op in ELEMENTWISE:               # broadcast.jl:196-204 ; elementwise.jl:323-343
            nd = len(self.dims[it.out])
            vals = [self._val(o).reshape(tuple(o.dims) + (1,) * (nd - len(o.dims)), order="F") for o in a]
            val, part, _ = dual_eval(it.expr, self.fconst, vals, self.dtype)
            it.cache = [np.array(p, order="F") for p in part]          # results[I].partial[p]
            out[...] = val
        elif op in (OP_GETINDEX, OP_GETINDEX_GENERIC):   # tracked.jl:331-340,377-388
            flat = self.value[a[0].buffer].reshape(-1, order="F")
            out[...] = flat[a[0].idx].reshape(out.shape, order="F")
        elif op == OP_DOT:                    # reductions.jl:106-112
            self.value[it.out][...] = np.dot(self._val(a[0]).reshape(-1, order="F"), self._val(a[1]).reshape(-1, order="F"))
        elif op == OP_SUMDIMS:                # reductions.jl:57-60
            x = self._val(a[0])
            axes = tuple(d for d in range(x.ndim) if out.shape[d] == 1 and x.shape[d] != 1)
            out[...] = np.sum(x, axis=axes, keepdims=True)
        elif op == OP_SCALAR_BLOCK:
            self._scalar_forward(it.block)
        else:
            raise NotImplementedError(f"opcode {op}")

    def _reverse_exec(self, it):
        op = it.opcode
        a = it.inputs
        delta = self.deriv[it.out]
        if op == OP_SCALAR_BLOCK:
            self._scalar_reverse(it.block)
            return
        if op == OP_MUL:                      # arithmetic.jl:260-285 (generic method on materialised operands)
            va, vb = self._val(a[0]), self._val(a[1])
            d2 = np.asarray(delta).reshape((va.shape[0], -1) if va.ndim == 2 else delta.shape, order="F")
            if a[0].tracked:
                vb2 = vb.reshape(vb.shape[0], -1, order="F")
                a_tmp = d2 @ vb2.T            # mul!(a_tmp, output_deriv, transpose(value(b)))
                self._inc(a[0], a_tmp.reshape(a[0].dims, order="F"))
            if a[1].tracked:
                b_tmp = va.T @ d2             # mul!(b_tmp, transpose(value(a)), output_deriv)
                self._inc(a[1], b_tmp.reshape(a[1].dims, order="F"))
        elif op == OP_ADD:                    # arithmetic.jl:45-53
            self._inc(a[0], delta); self._inc(a[1], delta)
        elif op == OP_SUB:                    # arithmetic.jl:127-140
            self._inc(a[0], delta); self._inc(a[1], delta, -1.0)
        elif op == OP_NEG:
            self._inc(a[0], delta, -1.0)
        elif op == OP_SUM:                    # reductions.jl:15-21 ; propagation.jl:34,38-43
            if a[0].tracked:
                self._inc(a[0], np.full(a[0].dims, np.asarray(delta).reshape(()), self.dtype))
        elif op == OP_MEAN:                   # reductions.jl:73-79
            if a[0].tracked:
                n = int(np.prod(a[0].dims))
                self._inc(a[0], np.full(a[0].dims, (self.dtype.type(1) / self.dtype.type(n)) *
                                        np.asarray(delta).reshape(()), self.dtype))
        elif op in ELEMENTWISE:               # broadcast.jl:163-194 ; elementwise.jl:349-408,466-537
            nd = len(self.dims[it.out])
            for p, o in enumerate(a):
                if not o.tracked:
                    continue
                contrib = delta * it.cache[p]                         # x[i] * getpartial(results[i], p)
                if o.cls == CLS_TR:                                   # sequential scalar accumulation (propagation.jl:98-108)
                    self._inc(o, np.add.accumulate(contrib.reshape(-1, order="F"))[-1])
                    continue
                odims = tuple(o.dims) + (1,) * (nd - len(o.dims))
                red = contrib
                for ax in range(nd):                                  # min(bound, I): size-1 dims collapse
                    if odims[ax] == 1 and red.shape[ax] != 1:
                        # sequential accumulation in column-major order (propagation.jl:90-95)
                        red = np.add.accumulate(red, axis=ax).take([-1], axis=ax)
                self._inc(o, red.reshape(o.dims, order="F"))
        elif op == OP_DOT:                    # reductions.jl:114-131
            dl = np.asarray(delta).reshape(())
            if a[0].tracked:
                self._inc(a[0], dl * self._val(a[1]).reshape(a[0].dims, order="F"))
            if a[1].tracked:
                self._inc(a[1], dl * self._val(a[0]).reshape(a[1].dims, order="F"))
        elif op == OP_SUMDIMS:                # reductions.jl:48-55 ; propagation.jl:217-223
            if a[0].tracked:
                self._inc(a[0], np.broadcast_to(delta, a[0].dims))
        elif op in (OP_GETINDEX, OP_GETINDEX_GENERIC):   # tracked.jl:318-329,363-376
            d = self.deriv[a[0].buffer]
            flat = d.reshape(-1, order="F")
            if a[0].unique:
                flat[a[0].idx] += np.asarray(delta).reshape(-1, order="F")
            else:
                np.add.at(flat, a[0].idx, np.asarray(delta).reshape(-1, order="F"))
            d[...] = flat.reshape(d.shape, order="F")
        else:
            raise NotImplementedError(f"opcode {op}")
        self._unseed(it.out)                  # every reverse exec ends with unseed!(output)


    # ------------------------------------------------------------------ scalar instructions
    def _flat(self, arrs, b):
        """Column-major flat VIEW of a buffer (buffers are allocated order='F', so this never copies)."""
        f = arrs[b].reshape(-1, order="F")
        assert f.size == 0 or np.shares_memory(f, arrs[b])
        return f

    def _scalar_forward(self, blk):
        """``scalar_forward_exec!`` for every instruction of the run, in tape order (``scalars.jl:80-123``): pull the input
        values (``pull_value!``: an origin-linked TrackedReal re-reads ``origin.value[index]``, ``tracked.jl:178-181``),
        one ``ForwardDiff.derivative!`` / ``gradient!`` evaluation, store the value and cache the partials."""
        if blk.run_c(self, forward=True):
            return
        dt = self.dtype.type
        val = blk.val = np.zeros(blk.n, self.dtype)
        pa = blk.pa = np.zeros(blk.n, self.dtype)
        pb = blk.pb = np.zeros(blk.n, self.dtype)
        flats = [self._flat(self.value, b) for b in blk.refbufs]

        def get(space, idx):
            if space == SP_POOL: return val[idx]
            if space == SP_LITERAL: return dt(self.fconst[idx])
            if space == SP_NONE: return None
            return flats[space][idx]
        for i in range(blk.n):
            args = [v for v in (get(blk.a_space[i], blk.a_idx[i]), get(blk.b_space[i], blk.b_idx[i])) if v is not None]
            eo, el, na = blk.exprs[blk.expr_id[i]]
            v, part, _ = dual_eval(blk.expr_words[eo: eo + 4 * el].reshape(-1, 4), self.fconst, args, self.dtype)
            val[i] = v
            if na >= 1: pa[i] = part[0]
            if na >= 2: pb[i] = part[1]
        for slot, r, e in blk.exports:            # the slot IS the buffer element for everything outside the run
            flats[r][e] = val[slot]

    def _scalar_reverse(self, blk):
        """``scalar_reverse_exec!`` in reverse tape order (``scalars.jl:52-74``): ``increment_deriv!(input, deriv(output) *
        partial)`` — for an origin-linked TrackedReal that is pull_deriv!, add, push_deriv! on ``origin.deriv[index]``
        (``propagation.jl:33-36``, ``tracked.jl:186-197``) — then ``unseed!(output)``."""
        dflats = [self._flat(self.deriv, b) for b in blk.refbufs]
        der = blk.der = np.zeros(blk.n, self.dtype)
        for slot, r, e in blk.exports:            # adjoints that later instructions pushed into the exported slots
            der[slot] = der[slot] + dflats[r][e]
            dflats[r][e] = 0                      # (one TrackedReal, one deriv field: it now lives in der[slot])
        if blk.run_c(self, forward=False):
            return
        for i in range(blk.n - 1, -1, -1):
            d = der[i]
            for space, idx, p in ((blk.a_space[i], blk.a_idx[i], blk.pa[i]), (blk.b_space[i], blk.b_idx[i], blk.pb[i])):
                if space == SP_POOL:
                    der[idx] = der[idx] + d * p
                elif space >= 0:
                    dflats[space][idx] = dflats[space][idx] + d * p
            der[i] = 0                            # unseed!(output)


class _ScalarBlock:
    """Parsed payload of an OP_SCALAR_BLOCK entry (layout: reversediff.jl_b200/program.py ``ScalarBlock``)."""

    use_c = True

    def __init__(self, w, expr_words):
        w = np.asarray(w, np.int64)
        n, nref, nexp, nex = (int(v) for v in w[:4])
        o = 4
        self.n = n
        self.refbufs = [int(v) for v in w[o:o + nref]]; o += nref
        self.exprs = [tuple(int(v) for v in row) for row in w[o:o + 3 * nex].reshape(-1, 3)]; o += 3 * nex
        self.exports = [tuple(int(v) for v in row) for row in w[o:o + 3 * nexp].reshape(-1, 3)]; o += 3 * nexp
        body = np.ascontiguousarray(w[o:o + 5 * n].reshape(-1, 5))
        self.body = body
        self.expr_id, self.a_space, self.a_idx, self.b_space, self.b_idx = (body[:, k].tolist() for k in range(5))
        self.expr_words = np.ascontiguousarray(expr_words, np.int32)
        self.expr_table = np.ascontiguousarray(np.asarray(self.exprs, np.int64).reshape(-1, 3))

    def run_c(self, tape, forward):
        """Same loops in C (oracle/scalar_block.c) when the helper library has been built — identical statements, used so
        that the CPU baseline is not a Python-interpreter benchmark.  Returns False when unavailable (pure-Python loops run)."""
        lib = _scalar_lib() if self.use_c else None
        if lib is None or tape.dtype != np.float64:
            return False
        import ctypes as C
        nref = len(self.refbufs)
        arrs = tape.value if forward else tape.deriv
        ptrs = (C.c_void_p * max(nref, 1))(*[tape._flat(arrs, b).ctypes.data for b in self.refbufs])
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        if forward:
            self.val = np.zeros(self.n); self.pa = np.zeros(self.n); self.pb = np.zeros(self.n)
            rc = lib.scalar_block_forward(C.c_int64(self.n), P(self.body), P(self.expr_table), P(self.expr_words), P(tape.fconst),
                                          ptrs, P(self.val), P(self.pa), P(self.pb))
            if rc != 0:
                raise ValueError(f"scalar_block_forward: unknown scalar opcode {rc}")
            flats = [tape._flat(tape.value, b) for b in self.refbufs]
            for slot, r, e in self.exports:
                flats[r][e] = self.val[slot]
        else:
            lib.scalar_block_reverse(C.c_int64(self.n), P(self.body), ptrs, P(self.der), P(self.pa), P(self.pb))
        return True


_SCALAR_LIB = [False]


def _scalar_lib():
    if _SCALAR_LIB[0] is False:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libscalar_block.so")
        try:
            lib = ctypes.CDLL(path)
            lib.scalar_block_forward.restype = ctypes.c_int
            lib.scalar_block_reverse.restype = ctypes.c_int
            _SCALAR_LIB[0] = lib
        except OSError:
            _SCALAR_LIB[0] = None
    return _SCALAR_LIB[0]


# ======================================================================================
# Hessian check (forward-over-reverse restated densely; small n only)
# ======================================================================================
def hessian_dense(program_bytes, x, eps=None):
    """Hessian of a scalar tape w.r.t. input slot 0 as the Jacobian of the replayed gradient, by
    *complex-free* central differences of ``OracleTape.gradient`` — an independent (non-AD) check used
    next to the closed forms; O(n) replays, small n only."""
    t = OracleTape(program_bytes)
    x = np.asarray(x, np.float64)
    n = x.size
    H = np.zeros((n, n), order="F")
    for j in range(n):
        h = eps or 1e-6 * max(1.0, abs(x.reshape(-1, order="F")[j]))
        xp = x.copy(order="F"); xm = x.copy(order="F")
        xp.reshape(-1, order="F")[j] += h; xm.reshape(-1, order="F")[j] -= h
        _, gp = t.gradient([xp]); _, gm = t.gradient([xm])
        H[:, j] = (gp[0].reshape(-1, order="F") - gm[0].reshape(-1, order="F")) / (2 * h)
    return H
