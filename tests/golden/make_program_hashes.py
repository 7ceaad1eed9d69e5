"""SHA-256 of the tape programs the Python recorder emits for the five BASELINE configs - the byte-level contract the Julia
lowering (julia/ReverseDiffB200.jl `lower`) must meet; julia/golden_roundtrip.jl records the same functions with ReverseDiff.jl
and compares.  Constants frozen into a tape are exactly representable (k/16) so both languages produce the same bits; record-time
input values do not enter a program (only shapes do).

    python tests/golden/make_program_hashes.py          # rewrites tests/golden/cfg*.program.sha256
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def pattern(rows, cols, a=7, b=13):
    """M[i, j] = ((a i + b j) mod 17 - 8) / 16, 0-based: exact in binary floating point."""
    i = np.arange(rows)[:, None]; j = np.arange(cols)[None, :]
    return np.asfortranarray(((a * i + b * j) % 17 - 8) / 16.0)


def programs():
    rd = graft.load_package()
    W = rd.workloads
    out = {}
    ones = lambda *s: np.ones(s, order="F")
    # cfg 1: examples/gradient.jl f(a, b) = sum(a' * b + a * b') at 100 x 100 (explicit transpose maps)
    out["cfg1"] = rd.GradientTape(W.f_gradient_example, (ones(100, 100), ones(100, 100))).program_bytes()
    # cfg 2: the same at 4096 x 4096, Float64 and Float32 (transposes travel as a tag above 2^20 elements)
    out["cfg2_f64"] = rd.GradientTape(W.f_gradient_example, (ones(4096, 4096), ones(4096, 4096))).program_bytes()
    out["cfg2_f32"] = rd.GradientTape(W.f_gradient_example, (ones(4096, 4096).astype(np.float32), ones(4096, 4096).astype(np.float32)),
                                      dtype=np.float32).program_bytes()
    # cfg 3: MLP with X (d x N) frozen in the tape, d = h = o = 8, N = 32
    d = h = o = 8; N = 32
    out["cfg3"] = rd.GradientTape(W.make_f_mlp(pattern(d, N)), (ones(h, d), np.ones(h), ones(o, h), np.ones(o))).program_bytes()
    # cfg 4: f(x) = (A*x + x) .* tanh.(B*x), n = 16
    n = 16
    out["cfg4"] = rd.JacobianTape(W.make_f_jacobian(pattern(n, n), pattern(n, n, 5, 3)), np.ones(n)).program_bytes()
    # cfg 5: extended Rosenbrock in map form, n = 16384 (the array tape the forward-over-reverse hessian! replays)
    out["cfg5"] = rd.HessianTape(W.f_rosenbrock, np.ones(16384)).program_bytes()
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, raw in programs().items():
        with open(os.path.join(here, f"{name}.program.sha256"), "w") as f:
            f.write(f"{hashlib.sha256(raw).hexdigest()}  {len(raw)}\n")
        print(name, len(raw), hashlib.sha256(raw).hexdigest())
