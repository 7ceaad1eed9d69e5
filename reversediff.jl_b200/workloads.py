"""The five BASELINE.json configs as recordable functions + their synthetic inputs.

Seeds: record-time inputs seed 1, replay inputs seed 2, constants frozen into the tape seed 3 — the
reference's tests also record at one random point and replay at another (test/api/GradientTests.jl:40-43).
"""
from __future__ import annotations

import numpy as np

from . import tracked as T
from . import expr as E


# cfg 1/2 — examples/gradient.jl:8  f(a, b) = sum(a' * b + a * b')
def f_gradient_example(a, b):
    return T.sum(a.T @ b + a @ b.T)


def inputs_gradient_example(n, dtype=np.float64, seed=2):
    rng = np.random.default_rng(seed)
    return (np.asfortranarray(rng.random((n, n)).astype(dtype)),
            np.asfortranarray(rng.random((n, n)).astype(dtype)))


# cfg 3 — sum(tanh.(W2*tanh.(W1*X .+ b1) .+ b2)), X frozen in the tape
def make_f_mlp(X):
    def f(W1, b1, W2, b2):
        h = T.bcast(lambda x, y: E.tanh(x + y), W1 @ X, b1)
        y = T.bcast(lambda x, y: E.tanh(x + y), W2 @ h, b2)
        return T.sum(y)
    return f


def inputs_mlp(d, h, o, N, dtype=np.float64, seed=2, const_seed=3):
    rng = np.random.default_rng(seed)
    crng = np.random.default_rng(const_seed)
    X = np.asfortranarray(crng.uniform(-1, 1, (d, N)).astype(dtype))
    W1 = np.asfortranarray((rng.uniform(-1, 1, (h, d)) / np.sqrt(d)).astype(dtype))
    b1 = rng.uniform(-0.1, 0.1, h).astype(dtype)
    W2 = np.asfortranarray((rng.uniform(-1, 1, (o, h)) / np.sqrt(h)).astype(dtype))
    b2 = rng.uniform(-0.1, 0.1, o).astype(dtype)
    return X, (W1, b1, W2, b2)


# cfg 4 — examples/jacobian.jl-style  f(x) = (A*x + x) .* tanh.(B*x)
def make_f_jacobian(A, B):
    def f(x):
        s = A @ x + x
        v = B @ x
        return T.bcast(lambda s_, v_: s_ * E.tanh(v_), s, v)
    return f


def inputs_jacobian(n, dtype=np.float64, seed=2, const_seed=3):
    rng = np.random.default_rng(seed)
    crng = np.random.default_rng(const_seed)
    A = np.asfortranarray((crng.random((n, n)) / n).astype(dtype))
    B = np.asfortranarray((crng.random((n, n)) / n).astype(dtype))
    return A, B, rng.random(n).astype(dtype)


# cfg 5 — extended Rosenbrock, benchmark/benchmark.jl:60-76 in map form
def f_rosenbrock(x):
    g = T.forward(lambda o, e: 100.0 * (e - o ** 2) ** 2 + (1.0 - o) ** 2)
    return T.sum(T.map_(g, x[0::2], x[1::2]))


def inputs_rosenbrock(n, dtype=np.float64, seed=2):
    return np.random.default_rng(seed).random(n).astype(dtype)


# cfg 5 as the reference itself would record it through scalar indexing (benchmark/benchmark.jl:60-76 is a loop over x[i]):
# every operation is one ScalarInstruction (macros.jl:84-131); scalar getindex records nothing (tracked.jl:348-351)
def f_rosenbrock_loop(x):
    acc = 0.0
    for i in range(x.size // 2):
        o, e = x[2 * i], x[2 * i + 1]
        acc = acc + (100.0 * (e - o ** 2) ** 2 + (1.0 - o) ** 2)
    return acc


# benchmark/benchmark.jl:84-95
def f_ackley_loop(x):
    a, b, c = 20.0, -0.2, 2.0 * np.pi
    n = x.size
    s1 = 0.0
    s2 = 0.0
    for i in range(n):
        s1 = s1 + x[i] ** 2
        s2 = s2 + E.cos(c * x[i])
    return -a * E.exp(b * E.sqrt(s1 / n)) - E.exp(s2 / n) + a + np.e
