"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/replay.py for the rules).

Literal restatement, in pure Python, of how the REFERENCE computes Hessians: reverse-over-reverse on a scalar tape.
``HessianTape`` (``src/api/tape.jl:285-293``) tracks the input on an OUTER tape ``jtp``, wraps those tracked scalars in a
second layer of ``TrackedReal`` on an INNER tape (``HessianConfig``, ``src/api/Config.jl:152-156``), runs ``f`` — every scalar
operation records one ``ScalarInstruction`` holding its partials (``src/macros.jl:84-131``) on the inner tape while the arithmetic
on the wrapped values records on the outer tape — then runs the inner reverse sweep (``scalar_reverse_exec!``,
``src/derivatives/scalars.jl:52-74``: ``input.deriv += output.deriv * partial``; ``unseed!(output)``), whose adjoint arithmetic is
also recorded on the outer tape.  ``hessian!`` is then the Jacobian of the recorded gradient: one one-hot seeded reverse sweep of the
outer tape per gradient component (``src/api/utils.jl:44-58``).

Small n only (O(n * tape length) Python operations).  The engine does NOT work this way (it replays the array tape
forward-over-reverse); agreement of the two is the parity statement for row H of SURVEY.md §8.
"""
from __future__ import annotations

import math


class Tape:
    def __init__(self):
        self.instrs = []          # (inputs, output, partials)


def _val(x):
    return x.value if isinstance(x, TReal) else x


def _zero_like(v):
    return 0.0


class TReal:
    """``TrackedReal{V,D}`` (``src/tracked.jl:47-57``); ``value``/``deriv`` may themselves be TReals of an outer tape."""
    __slots__ = ("value", "deriv", "tape")

    def __init__(self, value, tape):
        self.value, self.deriv, self.tape = value, 0.0, tape

    # ---- recording helpers: one ScalarInstruction per operation, partials from the DiffRules table
    @staticmethod
    def _rec(tape, inputs, value, partials):
        out = TReal(value, tape)
        tape.instrs.append((inputs, out, partials))
        return out

    def _bin(self, other, fval, pa, pb, swap=False):
        a, b = (other, self) if swap else (self, other)
        ta = a.tape if isinstance(a, TReal) else None
        tb = b.tape if isinstance(b, TReal) else None
        # an operand that lives on an OUTER tape (or a plain number) is a constant for this tape
        tape = self.tape
        av = a.value if (isinstance(a, TReal) and ta is tape) else a
        bv = b.value if (isinstance(b, TReal) and tb is tape) else b
        ins, parts = [], []
        if isinstance(a, TReal) and ta is tape:
            ins.append(a); parts.append(pa(av, bv))
        if isinstance(b, TReal) and tb is tape:
            ins.append(b); parts.append(pb(av, bv))
        return TReal._rec(tape, ins, fval(av, bv), parts)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, lambda a, b: 1.0, lambda a, b: 1.0)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, lambda a, b: 1.0, lambda a, b: 1.0, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, lambda a, b: 1.0, lambda a, b: -1.0)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, lambda a, b: 1.0, lambda a, b: -1.0, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, lambda a, b: b, lambda a, b: a)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, lambda a, b: b, lambda a, b: a, True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b, lambda a, b: 1.0 / b, lambda a, b: -(a / (b * b)))
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, lambda a, b: 1.0 / b, lambda a, b: -(a / (b * b)), True)

    def _un(self, fval, fder):
        v = self.value
        return TReal._rec(self.tape, [self], fval(v), [fder(v)])

    def __neg__(self): return self._un(lambda v: -v, lambda v: -1.0)

    def __pow__(self, n):
        assert isinstance(n, int) and n >= 1
        return self._un(lambda v: _powi(v, n), lambda v: n * _powi(v, n - 1) if n > 1 else 1.0)

    def tanh(self):
        return self._un(lambda v: tanh(v), lambda v: 1.0 - tanh(v) * tanh(v))

    def exp(self):
        return self._un(lambda v: exp(v), lambda v: exp(v))


def _powi(v, n):
    if n == 0:
        return 1.0
    r = v
    for _ in range(n - 1):
        r = r * v
    return r


def tanh(x):
    return x.tanh() if isinstance(x, TReal) else math.tanh(x)


def exp(x):
    return x.exp() if isinstance(x, TReal) else math.exp(x)


def reverse_pass(tape: Tape):
    """``reverse_pass!`` over scalar instructions (``src/tape.jl:85-90``; ``scalars.jl:52-74``)."""
    for ins, out, parts in reversed(tape.instrs):
        for x, p in zip(ins, parts):
            x.deriv = x.deriv + out.deriv * p          # increment_deriv!(input, output_deriv * partial)
        out.deriv = 0.0                                 # unseed!(output)


def hessian_reverse_over_reverse(f, x):
    """H(f)(x) the reference's way.  ``f`` takes a list of scalars (TReals) and returns one."""
    n = len(x)
    jtp, gtp = Tape(), Tape()
    outer = [TReal(float(v), jtp) for v in x]                   # jcfg.input            (Config.jl:153)
    inner = [TReal(o, gtp) for o in outer]                      # gcfg.input :: TrackedReal{TrackedReal}  (Config.jl:154)
    y = f(inner)                                                # GradientTape(f, ht.input, gcfg)  (api/tape.jl:290)
    for t in inner:
        t.deriv = 0.0
    y.deriv = 1.0                                               # seed!(output)
    reverse_pass(gtp)                                           # seeded_reverse_pass!(ht.output, gt.output, gt.input, gt.tape)
    grad = [t.deriv for t in inner]                             # outer-tape scalars: x -> grad f(x)
    value = y.value.value if isinstance(y.value, TReal) else y.value
    H = [[0.0] * n for _ in range(n)]
    for i in range(n):                                          # seeded_reverse_pass!(result, output::Vector, input, tape)  (api/utils.jl:44-58)
        for o in outer:
            o.deriv = 0.0
        gi = grad[i]
        if isinstance(gi, TReal):
            gi.deriv = 1.0
            reverse_pass(jtp)
            for j in range(n):
                H[i][j] = outer[j].deriv
            gi.deriv = 0.0
    g = [gi.value if isinstance(gi, TReal) else gi for gi in grad]
    return value, g, H
