import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xarray_parcel_b200 import _lib, synth
ctx = _lib.get_context(0); ctx.tables_build()
p, t, td = synth.model_level_columns(70001, 37, seed=8)
for profile in (False, True):
    dev = ctx.cape_cin(p.cuda(), t.cuda(), td.cuda(), kinds=("sb", "ml", "mu"), profile=profile)
    print("profile", profile, "dev exact count", ctx.last_exact_count())
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dev3 = ctx.cape_cin(p.cuda(), t.cuda(), td.cuda(), kinds=("sb", "ml", "mu"), profile=profile)
        s.synchronize()
    host = ctx.cape_cin(p, t, td, kinds=("sb", "ml", "mu"), profile=profile)
    host2 = ctx.cape_cin(p, t, td, kinds=("sb", "ml", "mu"), profile=profile)
    for kind in ("sb", "ml", "mu"):
        for f in ["cape", "lcl_pressure"] + (["profile_temperature"] if profile else []):
            a, a3, b, b2 = dev[kind][f].cpu(), dev3[kind][f].cpu(), host[kind][f], host2[kind][f]
            I = lambda x: x.view(torch.int32)
            print(" ", kind, f, "dev!=host", int((I(a) != I(b)).sum()), "dev!=dev_on_stream", int((I(a) != I(a3)).sum()),
                  "host!=host2", int((I(b) != I(b2)).sum()), "nan dev/host", int(torch.isnan(a).sum()), int(torch.isnan(b).sum()))
