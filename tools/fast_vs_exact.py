"""GPU box: float32 fast paths against the float64 exact kernel on millions of columns.
Reports decision mismatches (LFC/EL existence, integer outputs) and worst value errors."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xarray_parcel_b200 import _lib, synth

ctx = _lib.get_context(0); ctx.tables_build()
FIELDS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature",
          "lfc_pressure", "lfc_temperature", "el_pressure", "el_temperature"]
report = {}
for name, gen in (("era5_3.1M", lambda s: synth.era5_columns(1440 * 721 * 3, seed=s, device="cuda", nan_columns=0.002)),
                  ("model70_2M", lambda s: synth.model_level_columns(2_000_000, 70, seed=s, device="cuda"))):
    for seed in (1, 2):
        p, t, td = gen(seed)
        fast = ctx.cape_cin(p, t, td, kinds=("sb", "ml", "mu"))
        n_exact = ctx.last_exact_count()
        exact = ctx.cape_cin(p, t, td, kinds=("sb", "ml", "mu"), options=_lib.make_options(exact_only=True))
        rep = {"columns": t.shape[1], "exact_path_columns": n_exact}
        for kind in ("sb", "ml", "mu"):
            for f in FIELDS:
                a, b = fast[kind][f].double(), exact[kind][f].double()
                nanmis = int((torch.isnan(a) != torch.isnan(b)).sum())
                ok = ~torch.isnan(a) & ~torch.isnan(b)
                d = (a - b).abs()[ok]
                rel = d / b.abs()[ok].clamp(min=1.0)
                rep[f"{kind}_{f}"] = {"nan_mismatch": nanmis, "max_abs": float(d.max()) if d.numel() else 0.0,
                                      "max_rel": float(rel.max()) if rel.numel() else 0.0}
            rep[f"{kind}_level_shift_mismatch"] = int((fast[kind]["level_shift"] != exact[kind]["level_shift"]).sum())
        report[f"{name}_seed{seed}"] = rep
        worst = {k: v for k, v in rep.items() if isinstance(v, dict) and (v["nan_mismatch"] or v["max_rel"] > 6e-4)}
        print(name, seed, "exact-path columns", n_exact, "level_shift mismatches",
              [rep[f"{k}_level_shift_mismatch"] for k in ("sb", "ml", "mu")], "issues:", worst)
        print("   max cape abs err", max(rep[f"{k}_cape"]["max_abs"] for k in ("sb", "ml", "mu")),
              "max lfc_p rel", max(rep[f"{k}_lfc_pressure"]["max_rel"] for k in ("sb", "ml", "mu")),
              "max el_p rel", max(rep[f"{k}_el_pressure"]["max_rel"] for k in ("sb", "ml", "mu")))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", "fast_vs_exact.json"), "w"), indent=1)
