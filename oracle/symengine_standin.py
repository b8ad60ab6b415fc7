"""TEST INFRASTRUCTURE ONLY -- sympy-backed stand-in for the `symengine` module.

The reference (`/root/reference/traversome/ModelFitMaxLike.py:7`, `traversome.py:22`)
imports `symengine`, which is not installed in this image and cannot be fetched.
`install()` registers a process-local module named `symengine` exposing exactly the
three names the reference uses (`ModelFitMaxLike.py:113,179,523,534-540`):

    Symbol, log, lambdify(args=, exprs=, backend=)

so that every reference file imports and runs UNMODIFIED inside this container.
It is used only by `oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/`.  Every number produced through it is labelled
"reference + sympy stand-in for symengine".  Timings through it are meaningless
(sympy is orders of magnitude slower than symengine+LLVM) and are never reported.

The callable returned by `lambdify` takes one array argument (scipy style), returns
an array of len(exprs) float64 values and -- like the LLVM-compiled original --
yields inf/nan instead of raising when a proportion sits on the 0 boundary.
"""
import sys
import types

import numpy as np
import sympy


def _lambdify(args, exprs, backend=None, **_kw):
    args = tuple(args)
    fns = [sympy.lambdify(args, e, modules="numpy") for e in exprs]

    def call(x):
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            out = np.array([np.float64(f(*[np.float64(v) for v in x])) for f in fns], dtype=np.float64)
        return out

    return call


def install():
    if "symengine" in sys.modules and not getattr(sys.modules["symengine"], "_TVS_STANDIN", False):
        return sys.modules["symengine"]  # a real symengine is present: use it
    mod = types.ModuleType("symengine")
    mod._TVS_STANDIN = True
    mod.Symbol = sympy.Symbol
    mod.log = sympy.log
    mod.lambdify = _lambdify
    sys.modules["symengine"] = mod
    return mod
