"""Host <-> device copy bandwidth of this box, on ONE GPU or on all ranks of a torchrun launch AT THE SAME TIME.

    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py [--json F]

The end-to-end path of the suite (xp_suite with host buffers) moves 922 MB in and 411 MB out per 3.1 M ERA5 columns
and is PCIe-bound on one GPU; on N GPUs it is bound by whatever the HOST side of the box delivers to N links at once
(root complexes, host DRAM, the NUMA placement of the pinned buffers).  This probe measures exactly that ceiling with
plain pinned-memory copies of the same sizes: every rank copies H2D alone, D2H alone and both directions at once,
inside a barrier, so that the aggregate is the concurrent figure.  Rank 0 prints per-rank and aggregate GB/s; bench.py's
`e2e` at the same N is to be read against `duplex_aggregate_gbs`."""

import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--write-combined", action="store_true", help="input staging buffer with cudaHostAllocWriteCombined")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the host thread next to the GPU (NVML affinity)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpus = None
    if not a.no_bind:
        try:
            import sys
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from xarray_parcel_b200.partition import bind_host_thread_to_device
            cpus = bind_host_thread_to_device(local)
        except Exception:
            cpus = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_in, n_out = 922 * 1024 * 1024 // 4, 411 * 1024 * 1024 // 4
    if a.write_combined:
        # write-combined pinned memory: not snooped, faster for the device to read, slow for the CPU to read back
        import ctypes
        rt = ctypes.CDLL("libcudart.so")
        ptr = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n_in * 4), ctypes.c_uint(4)) == 0
        buf = (ctypes.c_float * n_in).from_address(ptr.value)
        h_in = torch.frombuffer(buf, dtype=torch.float32)
    else:
        h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
    h_in.fill_(1.0)
    d_in = torch.empty(n_in, dtype=torch.float32, device="cuda")
    h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
    d_out = torch.zeros(n_out, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = {"h2d": 1e9, "d2h": 1e9, "both": 1e9}
    for it in range(4):
        barrier(); t0 = time.perf_counter()
        d_in.copy_(h_in, non_blocking=True); barrier(); t1 = time.perf_counter()
        h_out.copy_(d_out, non_blocking=True); barrier(); t2 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        barrier(); t3 = time.perf_counter()
        if it:                                                   # first round warms the page tables up
            best["h2d"] = min(best["h2d"], t1 - t0); best["d2h"] = min(best["d2h"], t2 - t1)
            best["both"] = min(best["both"], t3 - t2)
    # the barriers make every interval the time of the SLOWEST rank: the aggregate is bytes of all ranks / that time
    mine = torch.tensor([best["h2d"], best["d2h"], best["both"]], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    if rank == 0:
        tmax = torch.stack(allv).max(0).values.tolist()
        res = {"n_gpus": world, "bytes_in_per_gpu": n_in * 4, "bytes_out_per_gpu": n_out * 4,
               "write_combined_input": bool(a.write_combined),
               "host_cpus_rank0": len(cpus) if cpus else None,
               "h2d_aggregate_gbs": world * n_in * 4 / tmax[0] / 1e9,
               "d2h_aggregate_gbs": world * n_out * 4 / tmax[1] / 1e9,
               "duplex_aggregate_gbs": world * (n_in + n_out) * 4 / tmax[2] / 1e9,
               "duplex_ms": tmax[2] * 1e3,
               "columns_per_s_ceiling_of_the_era5_suite": world * 3114720 / tmax[2],
               "per_rank_seconds_h2d_d2h_both": [v.tolist() for v in allv]}
        print(json.dumps(res))
        if a.json:
            with open(a.json, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
