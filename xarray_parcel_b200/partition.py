"""Column-block partitioning of the parcel path across the GPUs of one box.

Columns are independent, so there is NO collective inside the hot path: rank r of `world` owns a
contiguous block of the flattened column axis, runs the same kernels on it, and the only exchange is
the (optional) gather of the per-column outputs.  The reference expresses the same thing with dask
chunks (`map_blocks` PF:667, `apply_ufunc(dask='parallelized')` PF:585-592).
"""

import torch
import torch.distributed as dist


def column_block(n_columns, rank, world, align=32):
    """[start, stop) of the columns owned by `rank`: blocks differ by at most `align` columns and all
    but the last start on a multiple of `align` (keeps 128-byte coalescing for float32 rows)."""
    if world <= 1:
        return 0, int(n_columns)
    per = -(-int(n_columns) // world)
    per = -(-per // align) * align
    start = min(rank * per, int(n_columns))
    stop = min(start + per, int(n_columns))
    return start, stop


def shard_columns(arrays, rank, world, align=32):
    """Slice level-major [L, N] (or shared 1-D [L]) tensors to this rank's column block (views)."""
    out = []
    n = max(a.shape[1] for a in arrays if a.dim() == 2)
    start, stop = column_block(n, rank, world, align)
    for a in arrays:
        out.append(a if a.dim() == 1 else a[:, start:stop])
    return out, (start, stop)


def gather_columns(local, n_columns, group=None, align=32):
    """All-gather a per-column result [..., n_local] of every rank into [..., n_columns] (rank order =
    column order).  Uses the process group's backend (NCCL on GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    per = column_block(n_columns, 0, world, align)[1]
    pad = torch.zeros(local.shape[:-1] + (per,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.cat(parts, dim=-1)
    return out[..., :n_columns]


def bind_host_thread_to_device(device):
    """Pin the calling process to the CPU cores (and thereby the NUMA node) next to CUDA device `device`.

    The host-memory path (``xp_suite(mem=XP_MEM_HOST)``) is PCIe-bound; with one process per GPU on a
    multi-socket box its staging buffers should live on the socket the GPU hangs off, or the copies of
    several ranks meet on the inter-socket link.  Call this BEFORE allocating (pinned) host buffers: first
    touch then places them on the local node.  Returns the CPU list, or None if NVML is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                handle = None
        if handle is None:
            bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
