"""Plane trimming of the int8-slice FP64 `*` kernels (umma.cu: planes_publish_*, oz2_issue_trimmed, the trimmed cluster path).

Every slicing pass ORs the fixed-point images of the operand into one word; digit planes below its lowest set bit are identically
zero, and the tcgen05 kernels neither load nor multiply them.  Skipped products are exact zeros, so results must not change by a
single bit.  What is asserted:

* operands of few significant bits against an FP64 product that is exact by construction (every partial sum representable, so
  numpy, DGEMM and the sliced kernel must all agree with `==`), on every kernel the dispatcher can pick: the 2-CTA cluster kernel
  (with and without the role swap, both multicast geometries), the single-CTA kernel (odd number of column tiles) and the
  persistent cluster kernel (short contractions);
* the BASELINE gradient tape `sum(a'*b + a*b')`: the adjoint of `sum` is a constant fill (one live plane), so four of its six
  products issue 7 slice pairs instead of 28 - the counter says so - and the gradient is bit-identical to a process that runs
  with trimming switched off (RDB_OZAKI_TRIM=0).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _operand(rng, rows, cols, bits):
    """entries in [-1, 1] with `bits` fractional bits (bits = 0: all ones)"""
    if bits == 0:
        return np.ones((rows, cols))
    return np.round(rng.uniform(-1, 1, (rows, cols)) * 2.0 ** bits) / 2.0 ** bits


# (m, n, k): which kernel the dispatcher picks is noted per shape
SHAPES = [(512, 512, 4096),      # cluster kernel, role swap allowed, row tiles even (either multicast geometry); one-plane operands: CTA-pair kernel
          (1024, 768, 2304),     # the same, rectangular (the pair kernel's tile walk differs between C and C^T)
          (384, 512, 2304),      # cluster kernel, 3 row tiles: the B-role multicast needs an even count and is not taken unswapped
          (448, 512, 2304),      # cluster kernel, odd number of 64-wide row tiles: no role swap
          (512, 384, 640),       # swapped roles with an odd number of row tiles (the A-role multicast geometry)
          (256, 512, 40960),     # three exact-int32 K-chunks, each sliced (and trimmed) on its own
          (640, 448, 512),       # odd number of column tiles: single-CTA kernel
          (2560, 2048, 1024)]    # short contraction, more tiles than resident clusters: persistent cluster kernel


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("bits_a,bits_b", [(0, 36), (36, 0), (0, 0), (16, 16), (24, 8), (8, 24), (1, 33)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1)])
def test_trimmed_gemm_is_exact(rd, ctx, m, n, k, bits_a, bits_b, ta, tb):
    import torch
    assert bits_a + bits_b + int(np.ceil(np.log2(k))) + 1 <= 53          # every partial sum is exact in FP64
    rng = np.random.default_rng(m + 3 * n + 7 * k + 11 * bits_a + 13 * bits_b)
    opA = _operand(rng, m, k, bits_a); opB = _operand(rng, k, n, bits_b)
    A = opA.T if ta else opA; B = opB.T if tb else opB                    # as stored (column-major)
    C0 = np.round(rng.uniform(-4, 4, (m, n)) * 256.0) / 256.0
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).cuda(); dB = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
    ctx.set_gemm_f64_mode(2)
    try:
        for alpha, beta in [(1.0, 0.0), (2.0, 1.0)]:
            dC = torch.from_numpy(np.ascontiguousarray(C0.T)).cuda()
            ctx.gemm(0, ta, tb, m, n, k, alpha, dA.data_ptr(), A.shape[0], dB.data_ptr(), B.shape[0], beta, dC.data_ptr(), m)
            ctx.synchronize()
            got = dC.cpu().numpy().T
            ref = alpha * (opA @ opB) + beta * C0
            assert np.array_equal(got, ref), (alpha, beta, float(np.max(np.abs(got - ref))))
    finally:
        ctx.set_gemm_f64_mode(0)


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as graft
rd = graft.load_package()
n = {n}
rng = np.random.default_rng(5)
a0, b0 = rng.random((n, n)), rng.random((n, n))
f = rd.workloads.f_gradient_example if {objective!r} == "sum" else (lambda a, b: rd.mean(a.T @ b + a @ b.T))
tape = rd.GradientTape(f, (a0, b0))
ct = rd.compile(tape, use_graph={graph})
a, b = rng.standard_normal((n, n)), rng.standard_normal((n, n))
ga, gb = np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
res = rd.DiffResult(0.0, ga), rd.DiffResult(0.0, gb)
for _ in range({reps}):
    rd.gradient_(res, ct, (a, b))
np.savez({out!r}, ga=ga, gb=gb, value=res[0].value, pairs=ct.handle.guard_stats()["slice_pairs"], checked=ct.handle.guard_stats()["checked"])
"""


@pytest.mark.parametrize("n,graph", [(512, 0), (256, 1)])
def test_cfg2_adjoint_of_sum_is_one_plane_and_results_do_not_change(tmp_path, n, graph):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for trim in ("1", "0"):
        out = str(tmp_path / f"trim{trim}.npz")
        e = dict(os.environ); e["RDB_OZAKI_TRIM"] = trim
        r = subprocess.run([sys.executable, "-c", _CHILD.format(root=root, n=n, out=out, graph=graph, reps=3 if graph else 1, objective="sum")], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        outs[trim] = np.load(out)
    on, off = outs["1"], outs["0"]
    if not graph:                                                                    # (a graph replay counts its launches per replay)
        assert int(off["checked"]) == 6 and int(off["pairs"]) == 6 * 28              # six `*` launches, every slice pair
        assert int(on["checked"]) == 6 and int(on["pairs"]) == 2 * 28 + 4 * 7        # the four adjoint products meet a one-plane operand
    else:
        assert int(on["pairs"]) * 168 == int(off["pairs"]) * 84 and int(on["pairs"]) > 0
    assert np.array_equal(on["ga"], off["ga"]) and np.array_equal(on["gb"], off["gb"]) and float(on["value"]) == float(off["value"])


def test_trimmed_cta_pair_kernel_in_child_process():
    """umma_ozaki2tp_kernel (cta_group::2, one-plane products) is off by default (slower, see umma.cu): rerun the exactness tests of the
    shapes it takes with RDB_OZAKI_TPAIR=1."""
    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ); e["RDB_OZAKI_TPAIR"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_plane_trim.py"), "-m", "gpu", "-q", "-x",
                        "-k", "test_trimmed_gemm_is_exact and (512-512-4096 or 1024-768-2304)", "-p", "no:cacheprovider"],
                       cwd=os.path.dirname(here), env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.parametrize("objective,n,graph", [("sum", 512, 0), ("mean", 512, 0), ("mean", 256, 1)])
def test_uniform_adjoint_is_sliced_from_one_element_with_the_same_result(tmp_path, objective, n, graph):
    """deriv(x) right after the adjoint of sum / mean is a constant fill and the engine knows it (GemmState::uniform): its digit planes
    are written from element 0 (k_slice_uniform) instead of being cut from the matrix.  Same digits, same scales, same plane word -
    the gradient must equal, bit for bit, the one of a process that always runs the generic slicing passes (RDB_SLICE_UNIFORM=0).
    `mean` makes the fill value 1/n^2: a full-width mantissa, every plane live."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for uni in ("1", "0"):
        out = str(tmp_path / f"uni{uni}.npz")
        e = dict(os.environ); e["RDB_SLICE_UNIFORM"] = uni
        r = subprocess.run([sys.executable, "-c", _CHILD.format(root=root, n=n, out=out, graph=graph, reps=3 if graph else 1, objective=objective)],
                           env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        outs[uni] = np.load(out)
    on, off = outs["1"], outs["0"]
    assert int(on["pairs"]) == int(off["pairs"]) and int(on["pairs"]) > 0
    assert np.array_equal(on["ga"], off["ga"]) and np.array_equal(on["gb"], off["gb"]) and float(on["value"]) == float(off["value"])
