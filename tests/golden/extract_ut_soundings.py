#!/usr/bin/env python
"""Extract the sounding vectors of the reference's known-answer tests into a JSON fixture.

Run in the build container (where /root/reference exists):

    python tests/golden/extract_ut_soundings.py

It parses ``/root/reference/modules/unit_tests.py`` with ``ast`` (the module itself cannot
be imported: metpy/xarray are not installed), evaluates every simple assignment whose
right-hand side only needs numpy (``vert_array([...], 'hPa')``, ``np.array([...]) + 273.15``,
``xarray.DataArray(scalar, attrs=...)``), and writes the numeric results to
``tests/golden/ut_soundings.json`` as ``{test_function: {variable: [values]}}`` with NaN
encoded as null.  The expected answers (the ``assert_almost_equal`` pins) are written by hand
in ``tests/test_oracle_kat.py`` next to the line of unit_tests.py they come from.
"""

import ast
import json
import os
import sys

import numpy as np

SRC = "/root/reference/modules/unit_tests.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ut_soundings.json")


class _XarrayStub:
    @staticmethod
    def DataArray(value, *a, **k):
        return np.asarray(value, dtype=np.float64)


def _vert_array(x, units):
    return np.atleast_1d(np.asarray(x, dtype=np.float64))


def _jsonable(v):
    a = np.asarray(v, dtype=np.float64)
    if a.ndim == 0:
        f = float(a)
        return None if np.isnan(f) else f
    return [None if np.isnan(f) else float(f) for f in a.ravel().tolist()]


def main():
    tree = ast.parse(open(SRC).read())
    funcs = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
    out = {}
    for name, fn in funcs.items():
        if not (name.startswith("test_") or name == "multiple_intersections"):
            continue
        env = {"np": np, "xarray": _XarrayStub, "vert_array": _vert_array}
        # helper used by three tests: returns (levels, temperatures, dewpoints)
        env["multiple_intersections"] = lambda: (None, None, None)
        got = {}
        for node in ast.walk(fn):
            if not isinstance(node, ast.Assign) or len(node.targets) != 1:
                continue
            tgt = node.targets[0]
            if not isinstance(tgt, ast.Name):
                continue
            try:
                val = eval(compile(ast.Expression(node.value), SRC, "eval"), env)
                arr = np.asarray(val, dtype=np.float64)
            except Exception:
                continue
            if arr.size == 0:
                continue
            env[tgt.id] = val
            got[tgt.id] = _jsonable(arr)
        if got:
            out[name] = got
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(f"wrote {OUT}: {len(out)} functions, "
          f"{sum(len(v) for v in out.values())} vectors", file=sys.stderr)


if __name__ == "__main__":
    main()
