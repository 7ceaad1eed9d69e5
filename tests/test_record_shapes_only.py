"""Recording from DEVICE arrays is shapes-only (`TrackedArray{V,D,N,VA,DA}` with device storage,
src/tracked.jl:70-87; recorder src/api/tape.jl:197-209): no element of an input crosses to the host, the operator-overloading pass
infers shapes on zero-stride placeholders, and the program must be byte for byte the one a host recording produces.  Runs without
a GPU: `Dev` only quacks like a torch CUDA tensor (`is_cuda`, `data_ptr`, `shape`, `dtype`) and raises if anything tries to read it.
"""
import numpy as np
import pytest


class Dev:
    is_cuda = True

    def __init__(self, a):
        self.shape = tuple(a.shape)
        self.dtype = "torch.float64" if a.dtype == np.float64 else "torch.float32"

    def data_ptr(self):
        return 0

    def detach(self):                                   # the old route (copy the values to the host) must not be taken
        raise AssertionError("a shapes-only recording read a device array")

    def __array__(self, *a, **k):
        raise AssertionError("a shapes-only recording read a device array")


def _cases(rd):
    W = rd.workloads
    rng = np.random.default_rng(0)
    a0, b0 = rng.random((64, 64)), rng.random((64, 64))
    X, mlp_in = W.inputs_mlp(16, 24, 8, 40)
    A, B, xj = W.inputs_jacobian(48)
    xr = W.inputs_rosenbrock(32)
    many = [np.asfortranarray(rng.random(s)) for s in [(6, 10), (6, 10), (6,), (6, 10), (1, 10)]]

    def f_many(*xs):
        return rd.sum(rd.bcast(lambda a, b, c, d, e: a * b + rd.expr.tanh(c) * d - e, *xs))

    def f_mixed(x):
        y = x.reshape(8, 4)
        return rd.dot(y[:, 1], y[:, 2]) + rd.sum(rd.sum(y, dims=1)) + rd.mean(-y) + rd.sum(y[[0, 3, 5], :] .T @ y[[0, 3, 5], :])

    return [
        ("cfg2", rd.GradientTape, W.f_gradient_example, (a0, b0)),
        ("cfg2_f32", rd.GradientTape, W.f_gradient_example, (a0.astype(np.float32), b0.astype(np.float32))),
        ("cfg3", rd.GradientTape, W.make_f_mlp(X), tuple(mlp_in)),
        ("cfg4", rd.JacobianTape, W.make_f_jacobian(A, B), xj),
        ("cfg5", rd.HessianTape, W.f_rosenbrock, xr),
        ("loop", rd.GradientTape, W.f_rosenbrock_loop, xr),          # scalar-instruction tape (no value-dependent branches)
        ("nbcast5", rd.GradientTape, f_many, tuple(many)),
        ("mixed", rd.GradientTape, f_mixed, xr),
    ]


@pytest.mark.parametrize("idx", range(8))
def test_device_recording_is_shapes_only_and_byte_identical(rd, idx):
    name, Tape, f, inp = _cases(rd)[idx]
    host = Tape(f, inp)
    dev_inp = tuple(Dev(x) for x in inp) if isinstance(inp, tuple) else Dev(inp)
    dev = Tape(f, dev_inp)
    assert dev.tape.shapes_only and not host.tape.shapes_only, name
    assert dev.program_bytes() == host.program_bytes(), name
    assert dev.tape.dtype == host.tape.dtype
    for t in dev.input:                                   # placeholders own one element whatever the shape
        assert all(st == 0 for st in t.value.strides) and not t.value.flags.writeable


def test_value_dependent_control_flow_is_refused(rd):
    def f(x):
        s = x[0] * x[1]
        if s > 0.5:                                       # needs a value: cannot be recorded without one
            s = s * 2.0
        return s + rd.sum(x)

    x = np.arange(1.0, 5.0)
    rd.GradientTape(f, x)                                 # from host arrays: fine (the branch taken is frozen into the tape, as in the reference)
    with pytest.raises(TypeError, match="value-dependent control flow"):
        rd.GradientTape(f, Dev(x))


def test_mixed_host_and_device_inputs_record_shapes_only_too(rd):
    rng = np.random.default_rng(1)
    a0, b0 = rng.random((40, 40)), rng.random((40, 40))
    host = rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0))
    mixed = rd.GradientTape(rd.workloads.f_gradient_example, (Dev(a0), b0))
    assert mixed.tape.shapes_only and mixed.program_bytes() == host.program_bytes()


def test_nested_hessian_recording_from_a_device_array(rd):
    """`HessianTape(mode="reverse")` - the reference's own reverse-over-reverse scalar tape (nested.py) - from a device input."""
    x = rd.workloads.inputs_rosenbrock(16)
    host = rd.HessianTape(rd.workloads.f_rosenbrock_loop, x, mode="reverse")
    dev = rd.HessianTape(rd.workloads.f_rosenbrock_loop, Dev(x), mode="reverse")
    assert dev.tape.shapes_only and dev.program_bytes() == host.program_bytes()
