"""Generate tests/golden/lut_suite_golden.npz: LUT-mode results of the CPU oracle (oracle/, the NumPy
restatement of the reference) for a few seeded columns of each layout.

The reference itself cannot run in this environment (no xarray / MetPy / netCDF; its own LUT-mode pins,
test_data.nc + historic_results/*.nc, are absent from the checkout), so these vectors pin the ORACLE in LUT
mode against accidental change and give the CUDA path a committed fixture -- they are not an independent
check of the oracle (its exact-ODE mode is pinned against the reference's known answers, test_oracle_kat.py).

    python tests/golden/make_lut_goldens.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import parcel as op          # noqa: E402
from oracle import tables as otab        # noqa: E402
from xarray_parcel_b200 import synth     # noqa: E402

FIELDS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature", "lfc_pressure",
          "lfc_temperature", "el_pressure", "el_temperature"]


def main():
    tb = otab.load_tables()
    opts = op.Options(op.MoistLapseLUT(tb), lcl_mode="converged", metpy_compat="1.4.1")
    out = {}
    for name, (p, t, td) in (("era5", synth.era5_columns(96, seed=20260118, nan_columns=0.03)),
                             ("model70", synth.model_level_columns(96, 70, seed=20260119))):
        P = p.numpy().astype(np.float64)
        T, D = t.numpy().astype(np.float64), td.numpy().astype(np.float64)
        P2 = np.broadcast_to(P[:, None], T.shape) if P.ndim == 1 else P
        res = op.suite(P2, T, D, opts)
        out[name + "_pressure"] = p.numpy()
        out[name + "_temperature"] = t.numpy()
        out[name + "_dewpoint"] = td.numpy()
        for kind in ("sb", "ml", "mu"):
            for f in FIELDS:
                out[f"{name}_{kind}_{f}"] = res[f"{kind}_{f}"]
    np.savez_compressed(os.path.join(HERE, "lut_suite_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "lut_suite_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
