"""Chunk-streaming ingestion: a chunked / lazy / file-backed source is lifted block by block, with the
read of one block, the upload of another and the kernels / download of a third overlapping (xarray_parcel_b200/
streaming.py) -- and gives bit-identical results to one call over the whole field."""

import os

import numpy as np
import pytest

from xarray_parcel_b200 import streaming, synth


class LazyBlock:
    """A dask-like lazy array: slicing is free, ``compute()`` materialises (and counts)."""
    computed = 0

    def __init__(self, a):
        self.a, self.shape, self.ndim = a, a.shape, a.ndim

    def __getitem__(self, idx):
        return LazyBlock(self.a[idx])

    def compute(self):
        LazyBlock.computed += 1
        return np.array(self.a)


def test_iter_column_blocks_slices_without_reading():
    p, t, td = synth.era5_columns(1000, seed=1)
    LazyBlock.computed = 0
    T, D = LazyBlock(t.numpy()), LazyBlock(td.numpy())
    blocks = list(streaming.iter_column_blocks(p.numpy(), T, D, 300))
    assert [b[1].shape[1] for b in blocks] == [300, 300, 300, 100]
    assert LazyBlock.computed == 0                                   # nothing was materialised by the slicing
    assert all(b[0] is blocks[0][0] for b in blocks)                 # the shared 1-D pressure axis is passed through
    got = streaming._materialise(blocks[3][1])
    assert LazyBlock.computed == 1 and np.array_equal(got, t.numpy()[:, 900:])
    # NumPy inputs: the blocks are views (no host-side re-packing; the library reads them with their level stride)
    tb = next(iter(streaming.iter_column_blocks(p.numpy(), t.numpy(), td.numpy(), 300)))[1]
    assert tb.base is not None and tb.strides[0] == t.numpy().strides[0]


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["memmap", "lazy"])
def test_streamed_suite_is_bit_identical_to_one_call(tmp_path, source):
    import torch
    import xarray_parcel_b200.parcel_functions as pf
    pf.load_moist_adiabat_lookups()
    n = 230_000
    p, t, td = synth.era5_columns(n, seed=12)
    whole = pf.parcel_suite(p.numpy(), t.numpy(), td.numpy())
    if source == "memmap":                                           # file-backed fields, read block by block
        paths = {}
        for name, a in (("t", t), ("td", td)):
            paths[name] = os.path.join(tmp_path, name + ".npy")
            np.save(paths[name], a.numpy())
        T = np.load(paths["t"], mmap_mode="r")
        D = np.load(paths["td"], mmap_mode="r")
    else:
        LazyBlock.computed = 0
        T, D = LazyBlock(t.numpy()), LazyBlock(td.numpy())
    blocks = streaming.iter_column_blocks(p.numpy(), T, D, 50_000)
    parts = list(pf.parcel_suite_chunks(blocks, workers=2))
    assert [len(d["surface_cape"]) for d in parts] == [50_000] * 4 + [30_000]
    if source == "lazy":
        assert LazyBlock.computed == 10                              # every block of T and Td was read exactly once
    for k in whole:
        a = np.concatenate([d[k] for d in parts])
        b = np.asarray(whole[k])
        assert a.dtype == b.dtype and np.array_equal(a.view(np.int32), b.view(np.int32)), k
    torch.cuda.synchronize()


def _write_netcdf3(path, p, T, D, packed):
    """A classic-format file shaped like a reanalysis download: (time, level, lat, lon), optionally with the dewpoint
    packed as int16 with scale_factor / add_offset / _FillValue."""
    from scipy.io import netcdf_file
    nt, L, ny, nx = T.shape
    f = netcdf_file(path, "w", version=2)
    for d, n in (("time", nt), ("level", L), ("lat", ny), ("lon", nx)):
        f.createDimension(d, n)
    v = f.createVariable("pressure", "f4", ("level",)); v[:] = p
    v = f.createVariable("t", "f4", ("time", "level", "lat", "lon")); v[:] = T
    if packed:
        sc, off = 0.004, 250.0                       # int16 covers 119 .. 381 K
        v = f.createVariable("d2", "i2", ("time", "level", "lat", "lon"))
        v.scale_factor, v.add_offset, v._FillValue = sc, off, np.int16(-32767)
        q = np.round((D - off) / sc)
        assert np.abs(q).max() < 32767
        q = q.astype(np.int16)
        q[0, 3, 2, 1] = -32767
        v[:] = q
        D = q.astype(np.float32) * np.float32(sc) + np.float32(off)
        D[0, 3, 2, 1] = np.nan
    else:
        v = f.createVariable("d2", "f4", ("time", "level", "lat", "lon")); v[:] = D
    f.flush()
    return D


def _fields(n_time, ny, nx, seed):
    p, t, td = synth.era5_columns(n_time * ny * nx, seed=seed)
    L = t.shape[0]
    T = t.numpy().reshape(L, n_time, ny, nx).transpose(1, 0, 2, 3).copy()
    D = td.numpy().reshape(L, n_time, ny, nx).transpose(1, 0, 2, 3).copy()
    return p.numpy(), T, D


@pytest.mark.parametrize("packed", [False, True])
def test_netcdf3_blocks_are_lazy_views_of_the_file(tmp_path, packed):
    p, T, D = _fields(2, 6, 5, seed=3)
    L = T.shape[1]
    path = os.path.join(tmp_path, "fields.nc")
    D = _write_netcdf3(path, p, T, D, packed)
    blocks = list(streaming.iter_netcdf3_blocks(path, 13, vert_dim="level", names={"temperature": "t", "dewpoint": "d2"}))
    assert [b[1].shape for b in blocks] == [(L, 13), (L, 13), (L, 4)] * 2          # two time steps of 30 columns
    for step in range(2):
        got_t = np.concatenate([streaming._materialise(b[1]) for b in blocks[3 * step:3 * step + 3]], axis=1)
        got_d = np.concatenate([streaming._materialise(b[2]) for b in blocks[3 * step:3 * step + 3]], axis=1)
        assert got_t.dtype == np.float32 and np.array_equal(got_t, T[step].reshape(L, -1))
        assert np.array_equal(got_d, D[step].reshape(L, -1), equal_nan=True)
    assert np.array_equal(blocks[0][0], p) and blocks[0][0].ndim == 1              # the shared pressure axis
    assert np.isnan(streaming._materialise(blocks[0][2])).sum() == (1 if packed else 0)   # _FillValue -> NaN
    # a file whose vertical axis runs from the top down (pressure-level reanalysis files) is reversed on read
    top_down = list(streaming.iter_netcdf3_blocks(path, 100, vert_dim="level", surface_first=False,
                                                  names={"temperature": "t", "dewpoint": "d2"}))
    assert np.array_equal(streaming._materialise(top_down[1][1]), T[1].reshape(L, -1)[::-1])
    assert np.array_equal(top_down[0][0], p[::-1])
    with pytest.raises(AssertionError, match="not a dimension"):
        next(streaming.iter_netcdf3_blocks(path, 13, names={"temperature": "t", "dewpoint": "d2"}))


@pytest.mark.gpu
def test_streamed_netcdf3_file_is_bit_identical_to_one_call(tmp_path):
    import xarray_parcel_b200.parcel_functions as pf
    pf.load_moist_adiabat_lookups()
    p, T, D = _fields(3, 120, 150, seed=21)
    L = T.shape[1]
    path = os.path.join(tmp_path, "fields.nc")
    D = _write_netcdf3(path, p, T, D, packed=True)
    blocks = streaming.iter_netcdf3_blocks(path, 7000, vert_dim="level", names={"temperature": "t", "dewpoint": "d2"})
    parts = list(pf.parcel_suite_chunks(blocks, workers=2))
    assert len(parts) == 9                                                         # 3 time steps x 3 blocks of 18 000 columns
    t_all = np.ascontiguousarray(np.concatenate([T[s].reshape(L, -1) for s in range(3)], axis=1))
    d_all = np.ascontiguousarray(np.concatenate([D[s].reshape(L, -1) for s in range(3)], axis=1))
    whole = pf.parcel_suite(p, t_all, d_all)
    for k in whole:
        a = np.concatenate([d[k] for d in parts])
        b = np.asarray(whole[k])
        assert a.dtype == b.dtype and np.array_equal(a.view(np.int32), b.view(np.int32)), k
