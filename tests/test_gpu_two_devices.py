"""One process, two GPUs: the C ABI allows ``xp_create(device)`` per device (``_lib.get_context(i)``).  The > 48 KB
dynamic shared memory opt-in of the fast kernels is a per-DEVICE function attribute -- round 1 cached it per process and
the second device's launch failed; it is now set on every launch.  Needs two visible GPUs
(``gpurun --gpus 2``); skipped otherwise."""

import threading

import pytest
import torch

from xarray_parcel_b200 import _lib, synth

pytestmark = pytest.mark.gpu
FIELDS = ["cape", "cin", "lcl_pressure", "lfc_pressure", "el_pressure"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_contexts_on_two_devices_in_one_process():
    p, t, td = synth.era5_columns(200_000, seed=77)                      # float32, shared axis -> suite_fast_kernel
    pm, tm, tdm = synth.model_level_columns(50_000, 70, seed=78)         # per-column pressure -> suite_fast_pcol_kernel
    res = {}

    def work(dev):
        with torch.cuda.device(dev):
            ctx = _lib.get_context(dev)
            ctx.tables_build()
            d = torch.device("cuda", dev)
            a = ctx.cape_cin(p.to(d), t.to(d), td.to(d), kinds=("sb", "ml", "mu"))
            b = ctx.cape_cin(pm.to(d), tm.to(d), tdm.to(d), kinds=("sb", "ml", "mu"))
            torch.cuda.synchronize(d)
            assert ctx.last_exact_count() >= 0                          # the fast path ran on this device
            res[dev] = ({k: {f: v[f].cpu() for f in FIELDS} for k, v in a.items()},
                        {k: {f: v[f].cpu() for f in FIELDS} for k, v in b.items()})

    # device 1 FIRST (a process-wide cache would have been primed by device 0), then both from two threads at once
    work(1)
    work(0)
    first = {d: res[d] for d in (0, 1)}
    th = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for dev in (0, 1):
        for i in (0, 1):
            for kind in ("sb", "ml", "mu"):
                for f in FIELDS:
                    a, b = first[0][i][kind][f], res[dev][i][kind][f]
                    assert torch.equal(a.view(torch.int32), b.view(torch.int32)), (dev, i, kind, f)
                    assert torch.equal(a.view(torch.int32), first[1][i][kind][f].view(torch.int32)), (i, kind, f)
