"""Time the stand-alone column primitives (xp_layers.cu / xp_levels.cu / xp_derived.cu) on device-resident inputs with
CUDA events and print achieved GB/s of algorithmic bytes (inputs read once + outputs written once).

    python tools/bench_primitives.py [--columns 2000000] [--levels 70] [--reps 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xarray_parcel_b200 import _lib, synth  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=2_000_000)
    ap.add_argument("--levels", type=int, default=70)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    ctx = _lib.get_context(0)
    N, L = args.columns, args.levels
    p, t, td = [x.cuda() for x in synth.model_level_columns(N, L, seed=7, device="cuda", nan_columns=0, allnan_columns=0)]
    at = (p[0] - 100.0).contiguous()
    lev_t = torch.full((N,), 280.0, device="cuda")
    fb = 4 * N                                                      # bytes of one [N] float32 field
    k_ml = float((p >= (p[0] - 100.0)[None, :]).sum(0).float().mean()) + 1     # levels the mixed-layer loop touches
    cases = {
        "mixed_layer(2 fields, 100 hPa)": (lambda: ctx.mixed_layer(p, [t, td], depth=100.0), (L + 3 * k_ml) * fb + 2 * fb),
        "mixed_parcel(100 hPa)": (lambda: ctx.mixed_parcel(p, t, td, depth=100.0), (L + 3 * k_ml) * fb + 6 * fb),
        "layer_bounds": (lambda: ctx.layer_bounds(p, N, depth=300.0, interpolate=False), 2 * L * fb + 2 * fb),
        "interp_levels(2 fields, ln p)": (lambda: ctx.interp_levels(p, [t, td], at, log=True), 3 * L * fb + 3 * fb),
        "insert_level(2 fields)": (lambda: ctx.insert_level(p, at, [t, td], [lev_t, lev_t]), 3 * L * fb + 3 * (L + 1) * fb + 3 * fb),
        "shift_out_nans(3 fields)": (lambda: ctx.shift_out_nans(p, [p, t, td]), 4 * L * fb + 3 * L * fb + fb),
        "trapz(2 fields)": (lambda: ctx.trapz(p, [t, td]), 3 * L * fb + 2 * fb),
        "find_intersections(ln p)": (lambda: ctx.find_intersections(p, t, td + 5.0, log_x=True), 3 * L * fb + 6 * (L - 1) * fb),
        "trap_around_zeros(ln p)": (lambda: ctx.trap_around_zeros(p, t - td - 8.0, log_x=True), 2 * L * fb + 5 * (2 * L - 1) * fb + L * N),
        "valid_data": (lambda: ctx.valid_data(p, N), L * fb),
    }
    out = {"columns": N, "levels": L, "dtype": "f32", "kernels": {}}
    for name, (fn, nbytes) in cases.items():
        ms = timed(fn, args.reps)
        out["kernels"][name] = {"ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3),
                                "GB_per_s": round(nbytes / 1e9 / (ms / 1e3), 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
