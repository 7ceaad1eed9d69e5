"""The DEFAULT FP64 `*` kernel (exact int8 slices on tcgen05) on inputs built to break a norm-wise scheme:
magnitude spread INSIDE rows, products that cancel, K = 65536, denormals.  What is asserted:

* every launch's error stays inside the rigorous bound the kernel itself certifies (K c 2^ea 2^eb, umma.cu guard_coeff), measured
  against references that are exact by construction;
* the accuracy certificate (bound <= 2^-40 max|product|) passes where the norm-wise result is good to 1e-12 and FAILS where it
  is not, and a tape then replays the call with the DMMA kernels before returning (bit-identical to a DMMA-only tape);
* the component-wise error (|err| / (|A||B|)) is reported next to DMMA's for every case (printed with -s; recorded in
  profiles/README.md), because that is where a norm-wise scheme and DGEMM differ.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TAU = 2.0 ** -40


def _gemm(rd, ctx, A, B, mode):
    """C = A * B through rdb_gemm (column-major device buffers); returns (C, guard) with guard = (checked, flagged, worst)."""
    import torch
    m, k = A.shape; n = B.shape[1]
    dA, dB = rd.engine.device_array(A), rd.engine.device_array(B)
    dC = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    ctx.set_gemm_f64_mode(mode)
    ctx.gemm_guard()                                              # clear the counters
    ctx.gemm(0, 0, 0, m, n, k, 1.0, dA.data_ptr(), m, dB.data_ptr(), k, 0.0, dC.data_ptr(), m)
    ctx.synchronize()
    g = ctx.gemm_guard()
    ctx.set_gemm_f64_mode(0)
    return dC.cpu().numpy().reshape(m, n, order="F"), g


def _exact_product_51bit(rng, m, n, k):
    """Operands with ~51 significant bits whose product is known to ~2^-64: a = sum_t a_t 2^(-17 (t+1)) with 17-bit signed
    integer pieces, so every piece product A_t B_u is an integer matrix below 2^(32 + log2 k) <= 2^48 - exact in ANY FP64 GEMM,
    whatever its summation order - and the nine of them are combined in extended precision."""
    Ap = [rng.integers(-(1 << 16), 1 << 16, (m, k)).astype(np.float64) for _ in range(3)]
    Bp = [rng.integers(-(1 << 16), 1 << 16, (k, n)).astype(np.float64) for _ in range(3)]
    A = sum(Ap[t] * 2.0 ** (-17 * (t + 1)) for t in range(3))          # exact: disjoint bit ranges
    B = sum(Bp[t] * 2.0 ** (-17 * (t + 1)) for t in range(3))
    P = np.zeros((m, n), np.longdouble)
    for t in range(3):
        for u in range(3):
            P += (Ap[t] @ Bp[u]).astype(np.longdouble) * np.longdouble(2.0) ** (-17 * (t + u + 2))
    return np.asfortranarray(A), np.asfortranarray(B), P


def _longdouble_product(A, B, block=128):
    """A * B accumulated in x87 extended precision (64-bit mantissa), K-blocked to bound the temporary."""
    P = np.zeros((A.shape[0], B.shape[1]), np.longdouble)
    for k0 in range(0, A.shape[1], block):
        P += A[:, k0:k0 + block].astype(np.longdouble) @ B[k0:k0 + block].astype(np.longdouble)
    return P


def _report(name, C, Cd, P, A, B):
    absAB = (np.abs(A) @ np.abs(B)).astype(np.longdouble)
    e_i8 = np.abs(C.astype(np.longdouble) - P); e_dm = np.abs(Cd.astype(np.longdouble) - P)
    pmax = float(np.max(np.abs(P)))
    line = (f"[gemm-adversarial] {name}: norm-wise max|err|/max|P| int8 {float(e_i8.max()) / pmax:.3e} dmma {float(e_dm.max()) / pmax:.3e}; "
            f"component-wise max|err|/(|A||B|) int8 {float((e_i8 / absAB).max()):.3e} dmma {float((e_dm / absAB).max()):.3e}")
    print(line)
    return float(e_i8.max()), float(e_dm.max()), pmax


def test_k65536_gaussian_like_51bit_operands(rd, ctx):
    """K = 65536: four exact-int32 K-chunks, operands with 51 significant bits in every element (all 7 slices carry digits)."""
    rng = np.random.default_rng(11)
    m = n = 256; k = 65536
    A, B, P = _exact_product_51bit(rng, m, n, k)
    C, (chk, flg, worst) = _gemm(rd, ctx, A, B, 2)
    Cd, _ = _gemm(rd, ctx, A, B, 1)
    e8, ed, pmax = _report("K=65536 signed 51-bit", C, Cd, P, A, B)
    assert chk == 4 and flg == 0 and worst <= TAU                  # four chunks, all certified
    assert e8 <= worst * pmax                                      # the measured error is inside the certified bound
    assert e8 < 1e-12 * pmax and ed < 1e-12 * pmax


def test_in_row_magnitude_spread(rd, ctx):
    """Elements spread over 2^-46..1 (1e-14..1) inside every row / column.  Norm-wise the sliced product is as good as ever (small
    elements lose bits but weigh little); component-wise it is not DGEMM - both are reported.  The rigorous bound cannot know that
    the dropped slice pairs of 1024 tiny-but-non-zero elements cancel statistically, so the certificate is allowed to fail here
    (conservative: a tape would replay with DMMA); what must hold is the bound itself and the 1e-12 norm-wise accuracy."""
    rng = np.random.default_rng(12)
    m = n = 384; k = 1024
    A = np.asfortranarray(rng.standard_normal((m, k)) * np.ldexp(1.0, -rng.integers(0, 47, (m, k))))
    B = np.asfortranarray(rng.standard_normal((k, n)) * np.ldexp(1.0, -rng.integers(0, 47, (k, n))))
    P = _longdouble_product(A, B)
    C, (chk, flg, worst) = _gemm(rd, ctx, A, B, 2)
    Cd, _ = _gemm(rd, ctx, A, B, 1)
    e8, ed, pmax = _report("in-row spread 2^-46..1", C, Cd, P, A, B)
    print(f"[gemm-adversarial] in-row spread: certificate ratio {worst:.3e} (tau {TAU:.3e}), flagged {flg}")
    assert chk == 1
    assert e8 <= worst * pmax and e8 < 1e-12 * pmax


def test_one_hot_seed_operand_is_certified(rd, ctx):
    """jacobian!'s adjoint products have a one-hot seed matrix as one operand.  A bound of the form K * max|row| * max|col| would
    fail them (K = 2048 here, one non-zero per column); the certificate counts non-zeros and 1-norms per row, so they pass."""
    rng = np.random.default_rng(16)
    m = k = 2048; n = 512
    A = np.asfortranarray(rng.random((m, k)) / k)
    D = np.zeros((k, n), order="F")
    D[rng.permutation(k)[:n], np.arange(n)] = rng.random(n) + 0.5          # one scaled unit vector per column
    C, (chk, flg, worst) = _gemm(rd, ctx, A, D, 2)
    assert chk == 1 and flg == 0 and worst < 1e-14, (chk, flg, worst)
    assert np.max(np.abs(C - A @ D)) < 1e-15 * np.max(np.abs(C))


def test_certificate_fails_on_misaligned_magnitudes_and_tape_replays_with_dmma(rd, orc, ctx):
    """Row i of A is (1, eps, eps, ...), column j of B is (eps, 1, 1, ...)': every product is K*eps while the row / column
    maxima are 1 - the norm-wise bound (~K 2^-55) is far above 2^-40 of the result, the certificate must fail, and a tape must
    return what its DMMA-only twin returns."""
    m = n = k = 512
    eps = 2.0 ** -46
    rng = np.random.default_rng(13)
    A = np.full((m, k), eps) * (1 + rng.integers(0, 1 << 20, (m, k)) * 2.0 ** -20); A[:, 0] = 1.0
    B = np.ones((k, n)) * (1 + rng.integers(0, 1 << 20, (k, n)) * 2.0 ** -20); B[0, :] = eps
    A, B = np.asfortranarray(A), np.asfortranarray(B)
    P = _longdouble_product(A, B)
    C, (chk, flg, worst) = _gemm(rd, ctx, A, B, 2)
    Cd, _ = _gemm(rd, ctx, A, B, 1)
    e8, ed, pmax = _report("misaligned magnitudes", C, Cd, P, A, B)
    assert chk == 1 and flg == 1 and worst > TAU                    # flagged
    assert e8 <= worst * pmax                                       # ... and the bound it was flagged with holds
    assert ed < 1e-13 * pmax
    # the same product on a tape: gradient! replays the call with DMMA and says so
    f = lambda a, b: rd.sum(a @ b)
    t0 = rd.compile(rd.GradientTape(f, (A, B)), gemm_f64_mode=0)
    t1 = rd.compile(rd.GradientTape(f, (A, B)), gemm_f64_mode=1)
    r0 = (rd.DiffResult(0.0, np.zeros_like(A)), rd.DiffResult(0.0, np.zeros_like(B)))
    r1 = (rd.DiffResult(0.0, np.zeros_like(A)), rd.DiffResult(0.0, np.zeros_like(B)))
    rd.gradient_(r0, t0, (A, B)); rd.gradient_(r1, t1, (A, B))
    gs = t0.handle.guard_stats()
    assert gs["flagged"] >= 1 and gs["fallbacks"] == 1, gs
    assert r0[0].value == r1[0].value
    assert np.array_equal(r0[0].gradient, r1[0].gradient) and np.array_equal(r0[1].gradient, r1[1].gradient)
    assert np.array_equal(t0.handle.get_value(2), t1.handle.get_value(2))          # the `*` output itself (buffer 2)
    assert abs(float(r0[0].value) - float(P.sum())) < 1e-12 * float(P.sum())
    # three failing calls in a row and the tape stops trying the int8 path
    for _ in range(3):
        rd.gradient_(r0, t0, (A, B))
    gs2 = t0.handle.guard_stats()
    # calls 2 and 3 fail again (fallbacks 2, 3) and make the tape DMMA-only; call 4 launches no int8-slice kernel at all
    assert gs2["fallbacks"] == 3 and gs2["checked"] == 3 * gs["checked"], gs2


def test_heavy_cancellation(rd, ctx):
    """A = [A1 A1], B = [B1; -B1 + E]: the product collapses to A1*E, 2^-30 of what the row / column maxima suggest.  DGEMM's
    component-wise bound K u |A||B| is all anyone can promise here; the certificate fails and the fallback is DMMA."""
    rng = np.random.default_rng(14)
    m = n = 256; k = 512
    A1 = rng.integers(-(1 << 20), 1 << 20, (m, k)).astype(np.float64) * 2.0 ** -20
    B1 = rng.integers(-(1 << 20), 1 << 20, (k, n)).astype(np.float64) * 2.0 ** -20
    E = rng.integers(-(1 << 10), 1 << 10, (k, n)).astype(np.float64) * 2.0 ** -40
    A = np.asfortranarray(np.hstack([A1, A1])); B = np.asfortranarray(np.vstack([B1, -B1 + E]))
    P = _longdouble_product(A1, E)                                    # exact: 21-bit x 11-bit integers, 512 terms
    C, (chk, flg, worst) = _gemm(rd, ctx, A, B, 2)
    Cd, _ = _gemm(rd, ctx, A, B, 1)
    e8, ed, pmax = _report("heavy cancellation", C, Cd, P, A, B)
    assert flg == 1 and e8 <= worst * pmax
    absAB = np.abs(A) @ np.abs(B)
    assert np.max(np.abs(Cd - P.astype(np.float64)) / absAB) < 2 * k * 1.2e-16


def test_denormals(rd, ctx):
    """Denormals next to normal numbers are below every slice (flushed, norm-wise harmless); an all-denormal operand is scaled up
    like any other row and comes back within the subnormal quantum."""
    rng = np.random.default_rng(15)
    m = n = k = 256
    A = np.asfortranarray(rng.random((m, k))); B = np.asfortranarray(rng.random((k, n)))
    A[::3, ::5] = 4.9e-324 * rng.integers(1, 1000, A[::3, ::5].shape)
    P = _longdouble_product(A, B)
    C, (chk, flg, worst) = _gemm(rd, ctx, A, B, 0)
    assert flg == 0 and np.max(np.abs(C - P.astype(np.float64))) < 1e-12 * float(np.max(np.abs(P)))
    A2 = np.asfortranarray(rng.integers(1, 1 << 20, (m, k)).astype(np.float64) * 4.9406564584124654e-324)
    B2 = np.asfortranarray(rng.integers(1, 1 << 10, (k, n)).astype(np.float64))
    P2 = (A2 / 4.9406564584124654e-324) @ B2 * 4.9406564584124654e-324          # integers < 2^38: exact, times the quantum
    C2, _ = _gemm(rd, ctx, A2, B2, 0)
    assert np.max(np.abs(C2 - P2)) <= 2 * 4.9406564584124654e-324 * 1.0 + 1e-12 * np.max(np.abs(P2))
