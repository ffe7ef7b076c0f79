"""Builds tests/golden/scalar_kats.json: the reference's scalar known-answer tests.  TEST INFRASTRUCTURE ONLY.

The reference checks its radial / angular primitives, potentials and smoothing solvers against sympy restatements at
hand-picked arguments (mythos/energy/dna1/tests/test_base_functions.py:11-167, dna2/tests/test_base_functions.py:11-24,
dna1/tests/test_base_smoothing_functions.py, energy/tests/test_potentials.py).  This script runs IN THE BUILD CONTAINER
(it reads /root/reference, which does not exist on the GPU box) and records, for every parametrised case of those tests:

* ``symbolic``  -- the value of the reference's own sympy restatement (the "expected" side of the reference's test),
  from energy/tests/symbolic_potentials.py, dna1/tests/symbolic_base_functions.py,
  dna1/tests/symbolic_base_smoothing_functions.py, dna2/tests/symbolic_base_functions.py, loaded by file path;
* ``reference`` -- the value of the reference's actual function (mythos/energy/potentials.py,
  dna1/base_functions.py, dna2/base_functions.py, dna1/base_smoothing_functions.py), loaded by file path with numpy
  standing in for ``jax.numpy`` (the modules use only where / exp / log and arithmetic; jax itself is not installable
  here).

plus ``boundary`` cases that sit exactly ON the branch breakpoints (the reference's conditions are strict ``<``), answered
by the reference's functions under the same numpy shim.  The parametrize tables are read from the reference's test
files with ``ast`` (the test modules themselves import jax).

    python oracle/build_scalar_kats.py
"""

from __future__ import annotations

import ast
import importlib.util
import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/mythos")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "scalar_kats.json"


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _shim():
    """Package skeleton + numpy-as-jax.numpy so the reference's scalar modules import without jax / jax_md."""
    for pkg in ("mythos", "mythos.utils", "mythos.energy", "mythos.energy.tests", "mythos.energy.dna1", "mythos.energy.dna1.tests",
                "mythos.energy.dna2", "mythos.energy.dna2.tests"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    typ = types.ModuleType("mythos.utils.types")
    typ.Scalar = typ.ARR_OR_SCALAR = float
    sys.modules["mythos.utils.types"] = typ
    jax = types.ModuleType("jax")
    jax.numpy = np
    sys.modules["jax"] = jax
    sys.modules["jax.numpy"] = np


def _parametrized(path: Path) -> dict[str, tuple[list[str], list[tuple]]]:
    """test function name -> (argument names, list of argument tuples) from @pytest.mark.parametrize decorators."""
    out = {}
    for node in ast.parse(path.read_text()).body:
        if not isinstance(node, ast.FunctionDef):
            continue
        for dec in node.decorator_list:
            if isinstance(dec, ast.Call) and getattr(dec.func, "attr", "") == "parametrize":
                names = ast.literal_eval(dec.args[0])
                names = [n.strip() for n in names.split(",")] if isinstance(names, str) else list(names)
                out[node.name] = (names, [tuple(c) for c in ast.literal_eval(dec.args[1])])
    return out


def _val(x):
    return [float(v) for v in x] if isinstance(x, (tuple, list)) else float(x)


def main() -> None:
    _shim()
    sym_pot = _load("mythos.energy.tests.symbolic_potentials", REF / "energy/tests/symbolic_potentials.py")
    sym_bf = _load("mythos.energy.dna1.tests.symbolic_base_functions", REF / "energy/dna1/tests/symbolic_base_functions.py")
    sym_bsf = _load("mythos.energy.dna1.tests.symbolic_base_smoothing_functions",
                    REF / "energy/dna1/tests/symbolic_base_smoothing_functions.py")
    sym_bf2 = _load("mythos.energy.dna2.tests.symbolic_base_functions", REF / "energy/dna2/tests/symbolic_base_functions.py")
    pot = _load("mythos.energy.potentials", REF / "energy/potentials.py")
    bf = _load("mythos.energy.dna1.base_functions", REF / "energy/dna1/base_functions.py")
    bsf = _load("mythos.energy.dna1.base_smoothing_functions", REF / "energy/dna1/base_smoothing_functions.py")
    bf2 = _load("mythos.energy.dna2.base_functions", REF / "energy/dna2/base_functions.py")

    tables = [
        (REF / "energy/tests/test_potentials.py", sym_pot, pot, "energy/tests/test_potentials.py"),
        (REF / "energy/dna1/tests/test_base_functions.py", sym_bf, bf, "energy/dna1/tests/test_base_functions.py"),
        (REF / "energy/dna2/tests/test_base_functions.py", sym_bf2, bf2, "energy/dna2/tests/test_base_functions.py"),
        (REF / "energy/dna1/tests/test_base_smoothing_functions.py", sym_bsf, bsf, "energy/dna1/tests/test_base_smoothing_functions.py"),
    ]
    cases = []
    for path, sym, real, cite in tables:
        for test, (names, rows) in _parametrized(path).items():
            fn = test.removeprefix("test_")
            if fn == "harmonic":  # energy/tests/test_potentials.py:60 re-tests v_harmonic
                fn = "v_harmonic"
            if not hasattr(real, fn) and hasattr(real, "_" + fn):
                fn = "_" + fn
            for row in rows:
                args = [float(v) for v in row]
                with np.errstate(all="ignore"):
                    ref_v = getattr(real, fn)(*args)
                cases.append({"fn": fn.lstrip("_"), "args": dict(zip(names, args)), "arg_order": names,
                              "symbolic": _val(getattr(sym, fn)(*args)), "reference": _val(ref_v), "from": f"{cite}::{test}"})

    # on-the-breakpoint arguments: strict '<' sends them to the next branch / to 0 (answered by the reference's functions)
    boundary = []

    def b(fn, mod, **kw):
        with np.errstate(all="ignore"):
            boundary.append({"fn": fn, "args": {k: float(v) for k, v in kw.items()}, "arg_order": list(kw),
                             "reference": _val(getattr(mod, fn)(*[float(v) for v in kw.values()])), "from": "boundary"})

    f1 = dict(r_low=0.5, r_high=1.0, r_c_low=0.25, r_c_high=1.5, eps=1.0, a=2.0, r0=0.8, r_c=1.1, b_low=3.0, b_high=0.7)
    for r in (0.25, 0.5, 1.0, 1.5, 0.2499999, 0.5000001, 0.9999999, 1.4999999):
        b("f1", bf, r=r, **f1)
    f2 = dict(r_low=0.5, r_high=1.0, r_c_low=0.25, r_c_high=1.5, k=46.0, r0=0.575, r_c=0.675, b_low=3.0, b_high=0.7)
    for r in (0.25, 0.5, 1.0, 1.5, 0.75):
        b("f2", bf, r=r, **f2)
    f3 = dict(r_star=0.32, r_c=0.335, eps=2.0, sigma=0.33, b=892.0)
    for r in (0.32, 0.335, 0.31, 0.33, 0.4):
        b("f3", bf, r=r, **f3)
    f4 = dict(theta0=1.0, delta_theta_star=0.7, delta_theta_c=0.95, a=1.5, b=4.0)
    for th in (0.3, 1.7, 0.05, 1.95, 1.0, 0.1, 1.9, 3.0):
        b("f4", bf, theta=th, **f4)
    f5 = dict(x_star=-0.65, x_c=-0.77, a=2.0, b=10.9)
    for x in (0.0, -0.65, -0.77, 0.3, -0.3, -0.7, -0.9):
        b("f5", bf, x=x, **f5)
    for th in (0.5, 0.7, 0.2):
        b("f6", bf2, theta=th, a=40.0, b=0.5)

    OUT.write_text(json.dumps({"generated_by": "oracle/build_scalar_kats.py", "cases": cases, "boundary": boundary}, indent=1))
    print(f"{len(cases)} reference-test cases + {len(boundary)} boundary cases -> {OUT}")


if __name__ == "__main__":
    main()
