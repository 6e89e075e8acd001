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
