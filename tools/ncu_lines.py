"""Attribute the executed instructions and stall samples of an .ncu-rep kernel to CUDA source lines.

ncu's CSV source page is SASS-only; this joins it, instruction by instruction, with `nvdisasm -g` line info of a
cubin compiled from the SAME sources with the same flags (the build is deterministic).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep xarray_parcel_b200/csrc/xp_fast.cu \
        --kernel suite_fast_kernelILj7ELi1ELi512ELi0 [--columns 3114720] [--top 60]
"""
import argparse
import csv
import io
import os
import re
import subprocess
import tempfile
from collections import defaultdict

FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("source")
    ap.add_argument("--kernel", required=True, help="substring of the mangled kernel name")
    ap.add_argument("--columns", type=int, default=3114720)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--nvcc-flag", action="append", default=[])
    a = ap.parse_args()
    tmp = tempfile.mkdtemp()
    cubin = os.path.join(tmp, "k.cubin")
    subprocess.run(["nvcc"] + FLAGS + a.nvcc_flag + ["-cubin", "-o", cubin, a.source], check=True)
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
    lines, cur, inside = [], ("?", 0), False
    for l in dis:
        if l.startswith(".text."):
            inside = a.kernel in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines.append(cur)
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(int(r[iex]), int(r[ismp])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
    if len(data) != len(lines):
        raise SystemExit(f"SASS length mismatch: report {len(data)} vs cubin {len(lines)} (different build?)")
    agg = defaultdict(lambda: [0, 0, 0])
    for (e, s), key in zip(data, lines):
        g = agg[key]
        g[0] += e; g[1] += s; g[2] += 1
    tot_e = sum(d[0] for d in data)
    tot_s = sum(d[1] for d in data) or 1
    warps = a.columns / 32
    src_cache = {}

    def text(f, n):
        if f not in src_cache:
            p = os.path.join(os.path.dirname(a.source), f)
            src_cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
        s = src_cache[f]
        return s[n - 1].strip()[:100] if 0 < n <= len(s) else ""

    print(f"{tot_e / warps:.0f} executed instructions per warp; {tot_s} samples")
    print("samples%  inst/warp  sass  file:line  source")
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1][1])[: a.top]:
        print(f"{100 * g[1] / tot_s:6.2f}% {g[0] / warps:8.0f} {g[2]:5d}  {key[0]}:{key[1]}  {text(*key)}")


if __name__ == "__main__":
    main()
