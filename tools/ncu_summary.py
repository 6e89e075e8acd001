"""Summarise an .ncu-rep (one kernel) on the command line: the raw metrics the profiles/*.md files quote and,
from the source page, how the executed warp instructions split by how often a SASS instruction runs per warp
(the loop tiers) and by opcode.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--columns 3114720] [--loop-min 20] [--dump loop.txt]
"""
import argparse
import csv
import io
import subprocess
from collections import Counter, defaultdict

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
       "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--columns", type=int, default=3114720)
    ap.add_argument("--loop-min", type=float, default=20.0)
    ap.add_argument("--dump", default="")
    a = ap.parse_args()
    rows = ncu_csv(a.rep, "raw")
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "")[:100])
        for k in RAW:
            if k in d:
                print(f"  {k} = {d[k]}")
    rows = ncu_csv(a.rep, "source")
    hdr = rows[1]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(r[isrc], int(r[iex]), int(r[ismp])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
    tot = sum(d[1] for d in data)
    smp = sum(d[2] for d in data) or 1
    warps = a.columns / 32
    print(f"executed warp instructions {tot} = {tot / warps:.0f} per warp of 32 columns; {len(data)} SASS instructions")
    tiers = defaultdict(lambda: [0, 0, 0])
    for s, e, m in data:
        per = e / warps
        key = round(per) if per >= 1 else round(per, 1)
        t = tiers[key]
        t[0] += 1; t[1] += e; t[2] += m
    print("tiers (executions per warp : SASS instructions, share of executed, share of samples):")
    for k in sorted(tiers, key=lambda k: -tiers[k][1])[:14]:
        t = tiers[k]
        print(f"  x{k}: {t[0]} instrs, {100 * t[1] / tot:.1f}% of executed ({t[1] / warps:.0f}/warp), {100 * t[2] / smp:.1f}% of samples")
    ops = Counter()
    for s, e, m in data:
        f = s.split()
        op = f[1] if f[0].startswith("@") else f[0]
        ops[op.split(".")[0]] += e
    print("opcodes:", ", ".join(f"{op} {100 * c / tot:.1f}%" for op, c in ops.most_common(24)))
    if a.dump:
        with open(a.dump, "w") as f:
            for s, e, m in data:
                if e / warps >= a.loop_min:
                    f.write(f"{e / warps:6.1f} {m:5d}  {s}\n")


if __name__ == "__main__":
    main()
