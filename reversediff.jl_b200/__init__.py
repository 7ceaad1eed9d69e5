"""reversediff.jl_b200 — B200-native replay engine for ReverseDiff compiled tapes (host side).

The directory name carries a dot (fixed by the repo layout contract), so it is imported through
``__graft_entry__.load_package()`` (or ``tests/conftest.py``) under the module name
``reversediff_jl_b200``.

Scope: ONLY the compiled-tape replay path — ``compile`` + ``gradient!/jacobian!/hessian!(result,
tape, input)`` — see DESIGN.md.  Recording is the host's operator-overloading pass (``tracked.py``; shapes-only when the inputs
are device arrays: no element of them is read), replay happens in ``csrc/librdb200.so`` (hand-written sm_100a CUDA) through the
C ABI of ``include/rdb200.h``.
"""
from .tracked import (TrackedArray, TrackedReal, InstructionTape, ForwardOptimize, SkipOptimize,  # noqa: F401
                      forward, skip, value, istracked, sum, mean, dot, bcast, map_, broadcast, collect, convert)
from .expr import (tanh, exp, log, sqrt, sin, cos, abs2, abs_, inv, logistic, max_, min_, log1p, expm1, asin, acos, atan, sinh, cosh,
                   tan, hypot, exp2, exp10, log2, log10, cbrt, asinh, acosh, atanh, sec, csc, cot, sech, csch, coth)          # noqa: F401
from .api import (GradientTape, JacobianTape, HessianTape, CompiledTape, DiffResult, compile,      # noqa: F401
                  gradient_, jacobian_, hessian_, gradient, jacobian, hessian, GradientPipeline, seeded_forward_pass, shard_range, shard_bounds,
                  make_comm, gather_jacobian_, gather_hessian_)
from . import program, expr, engine, workloads                                                    # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
