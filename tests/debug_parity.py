"""Debug helper (GPU box; test infrastructure -- it uses the oracle, so it lives under tests/): compare the CUDA kernel with the host-compiled kernel code (tests/hostsim)
and the oracle on one configuration; dump the worst columns to gpurun_out/ for offline replay."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))       # run as: python tests/debug_parity.py
import hostsim_util as hs
from xarray_parcel_b200 import _lib, synth
from oracle import tables as otab

ctx = _lib.get_context(0); ctx.tables_build()
idx, cur = ctx.tables_get()
pl, tl = otab.default_grids()
tb = otab.AdiabatTables(pl, tl, idx, cur)
o = dict(virtual_temperature_correction=True, lcl_interp="linear", pos_cape_neg_cin=False, metpy_compat="1.6.2")
p, t, td = synth.model_level_columns(6000, 70, seed=11)
res = ctx.cape_cin(p.cuda(), t.cuda(), td.cuda(), kinds=("sb", "ml", "mu"), options=_lib.make_options(**o))
P, T, D = [x.numpy().astype(np.float64) for x in (p, t, td)]
dump = {}
for kind in ("sb", "ml", "mu"):
    h = hs.cape_cin(P, T, D, tb, kind=kind, vtc=True, lcl_interp="linear", pos_cape_neg_cin=False, metpy_compat=162)
    for f in ("cape", "cin", "lcl_pressure", "lfc_pressure", "el_pressure", "lcl_temperature"):
        a = res[kind][f].double().cpu().numpy(); b = h[f]
        d = np.abs(a - b) / np.maximum(np.abs(b), 1.0); d[np.isnan(d)] = 0
        nanmis = int((np.isnan(a) != np.isnan(b)).sum())
        bad = np.where(d > 1e-5)[0]
        print(kind, f, "max rel", d.max(), "n>1e-5:", bad.size, "nan mismatches:", nanmis, bad[:5],
              [(a[i], b[i]) for i in bad[:3]])
        for i in bad[:3]:
            dump[f"{kind}_{f}_{i}"] = {"p": P[:, i].tolist(), "t": T[:, i].tolist(), "td": D[:, i].tolist(),
                                       "gpu": float(a[i]), "host": float(b[i])}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dump, open(os.path.join(ROOT, "gpurun_out", "debug_cols.json"), "w"))
