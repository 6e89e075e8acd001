"""Chunk-streaming ingestion for the parcel suite (the I/O adjacency of the hot path).

The reference is driven chunk by chunk by dask (``map_blocks`` / ``apply_ufunc(dask='parallelized')``, PF:585-592, 667,
with the ``.chunk/.persist`` choreography of PF:561-579).  Here a chunked SOURCE -- dask / zarr arrays, ``numpy.memmap``
files, lazily indexed xarray variables, or any iterable of column blocks -- is consumed block by block without ever
materialising the whole field on the host or on the device:

* a loader thread materialises block i+1 (``.compute()`` / ``__array__``: disk or network I/O, decompression) while
* ``workers`` lifting threads, each with its OWN library context on the device (its own three-stream H2D / kernel /
  D2H pipeline and page-locked staging mirrors, xp_api.cu ``run_host``), process blocks i, i-1, ... concurrently, so the
  upload of one block overlaps the kernels and the download of another;
* results come back in source order, one ``Dataset`` of host arrays per block.

The column slices handed to the library are strided VIEWS of the caller's arrays (level stride = full row length): no
host-side re-packing.  File-backed sources that are tested: ``numpy.memmap`` / ``.npy`` files and netCDF CLASSIC /
64-bit-offset files (``iter_netcdf3_blocks``: memory-mapped through ``scipy.io.netcdf_file``, packed int16 variables
with ``scale_factor`` / ``add_offset`` / ``_FillValue`` unpacked block by block -- the form of ERA5 downloads and of
``nccopy -k nc6`` output).  netCDF-4 files such as the reference's ``test_data.nc`` are HDF5 containers: there is no HDF5
library in the build image, so those have to be converted (``nccopy -k nc6 in.nc out.nc``) first.
"""

import os
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib

__all__ = ["iter_column_blocks", "iter_netcdf3_blocks", "suite_blocks"]


def _materialise(x):
    """Anything array-like -> a NumPy array (float32 / float64) or a torch CPU tensor, without copying when possible."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "compute"):                 # dask array / delayed
        x = x.compute()
    if hasattr(x, "values") and not isinstance(x, np.ndarray):      # xarray.DataArray / Variable
        x = x.values
    a = np.asarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return a


def iter_column_blocks(pressure, temperature, dewpoint, block_columns):
    """Cut level-major fields [L, N] (pressure [L] or [L, N]) into column blocks of ``block_columns``.  The inputs may
    be lazy (dask, zarr, memmap): every block is only SLICED here, not read -- the loader thread of ``suite_blocks``
    materialises it."""
    n = temperature.shape[1]
    for s in range(0, n, int(block_columns)):
        e = min(n, s + int(block_columns))
        p = pressure if getattr(pressure, "ndim", len(getattr(pressure, "shape", ()))) == 1 else pressure[:, s:e]
        yield p, temperature[:, s:e], dewpoint[:, s:e]


_OPEN_NETCDF = {}


class _Unpacked:
    """A netCDF variable sliced lazily: ``_materialise`` reads (and unpacks: scale_factor / add_offset, missing -> NaN)
    only the block it is asked for."""

    def __init__(self, var, index, shape, dtype):
        self.var, self.index, self.shape, self.ndim, self.dtype = var, index, tuple(shape), len(shape), dtype

    def __getitem__(self, idx):
        assert isinstance(idx, tuple) and len(idx) == 2 and idx[0] == slice(None), "column blocks only: x[:, a:b]"
        return _Unpacked(self.var, self.index, self.shape, self.dtype).narrow(idx[1])

    def narrow(self, cols):
        self.cols = cols
        a, b, _ = cols.indices(self.shape[1])
        self.shape = (self.shape[0], max(0, b - a))
        return self

    def compute(self):
        raw = self.var[self.index]
        raw = raw.reshape(raw.shape[0], -1)                     # [L, columns]: trailing dimensions are contiguous
        cols = getattr(self, "cols", slice(None))
        raw = raw[:, cols]
        att = self.var._attributes
        out = np.array(raw, dtype=self.dtype)
        fill = [att[k] for k in ("_FillValue", "missing_value") if k in att]
        if "scale_factor" in att or "add_offset" in att:
            out = out * self.dtype(att.get("scale_factor", 1.0)) + self.dtype(att.get("add_offset", 0.0))
        for f in fill:
            out[np.asarray(raw) == np.asarray(f).reshape(-1)[0]] = np.nan
        return out


def iter_netcdf3_blocks(path, block_columns, vert_dim="model_level_number", names=None, dtype=np.float32,
                        surface_first=True):
    """Column blocks of a netCDF classic / 64-bit-offset file (memory-mapped, nothing is read until a block is
    materialised by ``suite_blocks``).  ``names`` maps 'pressure' / 'temperature' / 'dewpoint' to variable names of the
    file (default: those names; the reference's Datasets use them, PF:263).  Temperature and dewpoint must have the
    dimensions (..., vert_dim, y, x) or (..., vert_dim, column): the dimensions after ``vert_dim`` are flattened to
    columns and every index of the leading ones (time steps) is a field of its own; pressure may be the 1-D coordinate
    of ``vert_dim`` or have the shape of the temperature.  ``surface_first=False``: the vertical axis of the file runs
    from the model top to the surface (ERA5 pressure-level files) and is reversed on read.  Yields
    ``(pressure, temperature, dewpoint)`` like ``iter_column_blocks``, in file order (leading index major, column
    minor)."""
    from scipy.io import netcdf_file
    nm = {"pressure": "pressure", "temperature": "temperature", "dewpoint": "dewpoint"}
    nm.update(names or {})
    f = _OPEN_NETCDF.get(os.path.abspath(path))
    if f is None:
        # memory-mapped; kept open for the life of the process: blocks are read lazily by pool threads, possibly after
        # this generator is exhausted, and scipy cannot close a mapped file while views of it exist
        f = _OPEN_NETCDF[os.path.abspath(path)] = netcdf_file(path, "r", mmap=True, maskandscale=False)
    t_var, d_var, p_var = (f.variables[nm[k]] for k in ("temperature", "dewpoint", "pressure"))
    dims = list(t_var.dimensions)
    assert vert_dim in dims, f"{vert_dim!r} is not a dimension of {nm['temperature']!r} {tuple(dims)}"
    assert tuple(d_var.dimensions) == tuple(dims), "temperature and dewpoint must share their dimensions"
    iv = dims.index(vert_dim)
    lead = t_var.shape[:iv]
    L = t_var.shape[iv]
    n = int(np.prod(t_var.shape[iv + 1:], dtype=np.int64))
    p_is_1d = tuple(p_var.dimensions) == (vert_dim,)
    assert p_is_1d or tuple(p_var.dimensions) == tuple(dims), "pressure: the vertical coordinate or a full field"
    flip = (slice(None, None, -1),) if not surface_first else (slice(None),)

    class _Var:                                 # a variable with the leading index applied and the vertical axis oriented
        def __init__(self, var):
            self._attributes = var._attributes
            self.var = var

        def __getitem__(self, index):
            return self.var[(index or ()) + flip]

    p1 = None
    if p_is_1d:
        p1 = _Unpacked(_Var(p_var), None, (L, 1), np.float64).compute().reshape(L).astype(dtype)
    for index in np.ndindex(*lead):
        for s in range(0, n, int(block_columns)):
            cols = slice(s, min(n, s + int(block_columns)))
            blk = [_Unpacked(_Var(v), tuple(index), (L, n), dtype).narrow(cols) for v in (t_var, d_var)]
            pb = p1 if p_is_1d else _Unpacked(_Var(p_var), tuple(index), (L, n), dtype).narrow(cols)
            yield pb, blk[0], blk[1]


def _to_tensor(a):
    if isinstance(a, torch.Tensor):
        return a
    if not a.flags.writeable:                 # read-only memmaps: torch wants writable memory (it is only read)
        a = np.require(a, requirements=["W"]) if a.flags.owndata else np.array(a)
    return torch.from_numpy(a)


def suite_blocks(blocks, kinds=("sb", "ml", "mu"), workers=2, queue_depth=2, device=None, options=None,
                 specific_humidity=False, fields=None):
    """Generator: lift every block of ``blocks`` (an iterable of (pressure, temperature, dewpoint), level-major) and
    yield ``{kind: {field: torch CPU tensor [n_i]}}`` per block, in order.  See the module docstring for the pipeline.

    ``workers`` contexts are created on ``device`` (each holds its own copy of the lookup tables, built on the GPU in
    ~12 ms); at most ``workers + queue_depth`` blocks are alive at any time."""
    dev = _lib.get_context(device).device
    ctxs = []
    for _ in range(max(1, int(workers))):
        c = _lib.Context(dev)
        c.tables_build()
        ctxs.append(c)
    free = deque(ctxs)
    lock = threading.Lock()
    opts = options if options is not None else _lib.make_options()

    def lift(block):
        p, t, td = [_to_tensor(_materialise(x)) for x in block]      # I/O happens here, in a pool thread
        if p.dtype != t.dtype:
            p = p.to(t.dtype)
        if td.dtype != t.dtype:
            td = td.to(t.dtype)
        with lock:
            ctx = free.popleft()
        try:
            with torch.cuda.device(dev):
                out = ctx.alloc_outputs(t, kinds, False, False, fields=fields)
                return ctx.cape_cin(p, t, td, kinds=kinds, options=opts, out=out, specific_humidity=specific_humidity)
        finally:
            with lock:
                free.append(ctx)

    pending = deque()
    try:
        with ThreadPoolExecutor(max_workers=len(ctxs)) as pool:
            for block in blocks:
                pending.append(pool.submit(lift, block))
                while len(pending) >= len(ctxs) + int(queue_depth):
                    yield pending.popleft().result()
            while pending:
                yield pending.popleft().result()
    finally:
        for c in ctxs:
            c.close()
