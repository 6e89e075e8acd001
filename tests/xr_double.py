"""A small stand-in for the ``xarray`` package -- TEST INFRASTRUCTURE ONLY.

xarray cannot be installed in the build image (no network, not in the wheelhouse), yet the drop-in boundary of the
reference is "accepts and returns xarray objects" (modules/parcel_functions.py returns ``xarray.Dataset``s with
variable attrs, renames them with a prefix PF:1508-1512 and re-labels the vertical coordinate PF:968-970).  This module
implements, with xarray's documented semantics, exactly the subset of the API that
``xarray_parcel_b200.parcel_functions`` touches in its ``is_xr`` branch, so that branch can be executed:

    DataArray(data, dims, coords, name, attrs): .data .values .dims .shape .ndim .dtype .coords .attrs .name,
        da[dim] (coordinate), isel({dim: i}, drop), broadcast_like(other), transpose(*dims), assign_coords,
        where(cond), arithmetic with scalars / aligned DataArrays, np.asarray(da)
    Dataset({name: DataArray}): ds[name], ds.name, .data_vars, .keys(), .attrs, rename, drop_vars, items()
    broadcast(*arrays), merge([datasets])

``install()`` registers it as ``sys.modules['xarray']`` and reloads the host layer; ``uninstall()`` undoes that.
Dimension order follows xarray: ``broadcast`` / ``broadcast_like`` order the result like the FIRST / the OTHER
operand, new dimensions are inserted by ``numpy.broadcast_to`` views.
"""

import importlib
import sys
import types

import numpy as np


def _as_np(x):
    if hasattr(x, "detach"):                       # torch tensor
        return x.detach().cpu().numpy()
    return np.asarray(x)


class DataArray:
    __array_priority__ = 50

    def __init__(self, data, dims=None, coords=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            dims = data.dims if dims is None else dims
            coords = data.coords if coords is None else coords
            name = data.name if name is None else name
            attrs = data.attrs if attrs is None else attrs
            data = data.data
        self.data = data if hasattr(data, "shape") else np.asarray(data)
        nd = len(self.data.shape)
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(nd))
        self.dims = (dims,) if isinstance(dims, str) else tuple(dims)
        if len(self.dims) != nd:
            raise ValueError(f"different number of dimensions on data and dims: {nd} vs {len(self.dims)}")
        self.name = name
        self.attrs = dict(attrs or {})
        self.coords = {}
        for k, v in (coords or {}).items():
            if isinstance(v, DataArray):
                c = DataArray(v.data, v.dims, None, k, v.attrs)
            elif isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], (str, tuple, list)):
                c = DataArray(np.asarray(v[1]), v[0], None, k)
            else:
                arr = np.asarray(v)
                c = DataArray(arr, (k,) if arr.ndim == 1 else (), None, k)
            for d, n in zip(c.dims, c.shape):
                if d in self.dims and self.shape[self.dims.index(d)] != n:
                    raise ValueError(f"conflicting sizes for dimension {d!r}")
            if all(d in self.dims for d in c.dims):          # xarray drops nothing silently; neither do we
                self.coords[k] = c
            else:
                raise ValueError(f"coordinate {k} has dimensions {c.dims} not on the array {self.dims}")

    # ---- basic properties
    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def values(self):
        return _as_np(self.data)

    @property
    def sizes(self):
        return dict(zip(self.dims, self.shape))

    def __array__(self, dtype=None, copy=None):
        a = self.values
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"<xr_double.DataArray {self.name!r} {dict(zip(self.dims, self.shape))}>"

    def _replace(self, data, dims=None, coords=None):
        dims = self.dims if dims is None else dims
        coords = self.coords if coords is None else coords
        return DataArray(data, dims, {k: v for k, v in coords.items() if all(d in dims for d in v.dims)},
                         self.name, self.attrs)

    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self.coords:
                return self.coords[key]
            if key in self.dims:                   # dimension without a coordinate: positional labels
                return DataArray(np.arange(self.shape[self.dims.index(key)]), (key,), None, key)
            raise KeyError(key)
        raise TypeError("the test double supports da['dim'] only; use isel for positional indexing")

    # ---- selection / reshaping
    def isel(self, indexers=None, drop=False, **kw):
        indexers = dict(indexers or {}, **kw)
        out = self
        for dim, idx in indexers.items():
            ax = out.dims.index(dim)
            sl = [slice(None)] * out.ndim
            sl[ax] = idx
            data = out.data[tuple(sl)]
            scalar = not isinstance(idx, slice) and np.ndim(idx) == 0
            dims = tuple(d for d in out.dims if d != dim) if scalar else out.dims
            coords = {}
            for k, c in out.coords.items():
                if dim in c.dims:
                    if scalar and drop:
                        continue                   # xarray: drop=True drops the now-scalar coordinates
                    coords[k] = c.isel({dim: idx})
                else:
                    coords[k] = c
            out = DataArray(data, dims, coords, out.name, out.attrs)
        return out

    def transpose(*args):
        self, dims = args[0], args[1:]
        if not dims:
            dims = self.dims[::-1]
        if set(dims) != set(self.dims):
            raise ValueError(f"{dims} must be a permutation of {self.dims}")
        perm = [self.dims.index(d) for d in dims]
        data = self.data.permute(*perm) if hasattr(self.data, "permute") else np.transpose(self.data, perm)
        return self._replace(data, tuple(dims))

    def _expand_to(self, dims, sizes):
        """View of self with dimensions `dims` (a superset of self.dims), in that order."""
        own = [d for d in dims if d in self.dims]
        x = self.transpose(*own) if tuple(own) != self.dims else self
        data = x.values if not hasattr(x.data, "expand") else x.data
        shape = [sizes[d] if d in self.dims else 1 for d in dims]
        data = data.reshape(shape)
        full = tuple(sizes[d] for d in dims)
        data = data.expand(*full) if hasattr(data, "expand") else np.broadcast_to(data, full)
        return data

    def broadcast_like(self, other):
        sizes = dict(other.sizes)
        for d, n in self.sizes.items():
            if d in sizes and sizes[d] != n:
                raise ValueError(f"cannot broadcast dimension {d}: {n} vs {sizes[d]}")
            sizes.setdefault(d, n)
        dims = tuple(other.dims) + tuple(d for d in self.dims if d not in other.dims)
        coords = dict(other.coords)
        coords.update(self.coords)
        return DataArray(self._expand_to(dims, sizes), dims, coords, self.name, self.attrs)

    def assign_coords(self, coords=None, **kw):
        c = dict(self.coords)
        c.update(dict(coords or {}, **kw))
        return DataArray(self.data, self.dims, c, self.name, self.attrs)

    def where(self, cond, other=np.nan):
        c = cond.broadcast_like(self).transpose(*self.dims).values if isinstance(cond, DataArray) else np.asarray(cond)
        o = other.broadcast_like(self).transpose(*self.dims).values if isinstance(other, DataArray) else other
        return self._replace(np.where(c, self.values, o))

    def copy(self):
        return self._replace(self.values.copy())

    def astype(self, dt):
        return self._replace(self.values.astype(dt))

    def load(self):
        return self

    # ---- arithmetic (aligned by dimension name, no index alignment: the tests use identical coordinates)
    def _binary(self, other, fn, reflexive=False):
        if isinstance(other, DataArray):
            a, b = broadcast(self, other)
            x, y = a.values, b.values
            res = fn(y, x) if reflexive else fn(x, y)
            return a._replace(res)
        res = fn(other, self.values) if reflexive else fn(self.values, other)
        return self._replace(res)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __rsub__(self, o): return self._binary(o, np.subtract, True)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __truediv__(self, o): return self._binary(o, np.divide)
    def __lt__(self, o): return self._binary(o, np.less)
    def __le__(self, o): return self._binary(o, np.less_equal)
    def __gt__(self, o): return self._binary(o, np.greater)
    def __ge__(self, o): return self._binary(o, np.greater_equal)
    def __invert__(self): return self._replace(~self.values)
    def __neg__(self): return self._replace(-self.values)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = {}
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if not isinstance(v, DataArray):
            v = DataArray(v)
        self._vars[k] = DataArray(v.data, v.dims, v.coords, k, v.attrs)

    def __getitem__(self, k):
        if isinstance(k, (list, tuple)):
            return Dataset({n: self._vars[n] for n in k}, attrs=self.attrs)
        if k in self._vars:
            return self._vars[k]
        for v in self._vars.values():              # a coordinate shared by the variables
            if k in v.coords:
                return v.coords[k]
        raise KeyError(k)

    def __getattr__(self, k):
        if k.startswith("_"):
            raise AttributeError(k)
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __contains__(self, k):
        return k in self._vars

    def __iter__(self):
        return iter(self._vars)

    def __len__(self):
        return len(self._vars)

    def keys(self):
        return self._vars.keys()

    def items(self):
        return self._vars.items()

    def values(self):
        return self._vars.values()

    @property
    def data_vars(self):
        return dict(self._vars)

    @property
    def dims(self):
        out = {}
        for v in self._vars.values():
            out.update(v.sizes)
        return out

    def rename(self, mapping):
        return Dataset({mapping.get(k, k): v for k, v in self._vars.items()}, attrs=self.attrs)

    def drop_vars(self, names):
        names = [names] if isinstance(names, str) else list(names)
        return Dataset({k: v for k, v in self._vars.items() if k not in names}, attrs=self.attrs)

    def isel(self, indexers=None, drop=False, **kw):
        ix = dict(indexers or {}, **kw)
        return Dataset({k: v.isel({d: i for d, i in ix.items() if d in v.dims}, drop=drop)
                        for k, v in self._vars.items()}, attrs=self.attrs)

    def load(self):
        return self

    def __repr__(self):
        return f"<xr_double.Dataset {list(self._vars)}>"


def broadcast(*arrays):
    """xarray.broadcast: every result has the union of the dimensions, ordered by first appearance."""
    dims, sizes = [], {}
    for a in arrays:
        for d, n in a.sizes.items():
            if d not in sizes:
                dims.append(d)
                sizes[d] = n
            elif sizes[d] != n:
                raise ValueError(f"cannot broadcast dimension {d}: {n} vs {sizes[d]}")
    coords = {}
    for a in arrays:
        coords.update(a.coords)
    out = []
    for a in arrays:
        out.append(DataArray(a._expand_to(tuple(dims), sizes), tuple(dims), coords, a.name, a.attrs))
    return tuple(out)


def merge(objects):
    out = Dataset()
    for o in objects:
        for k, v in (o.items() if isinstance(o, Dataset) else [(o.name, o)]):
            out[k] = v
    return out


_saved = {}


def install():
    """Register this module as ``xarray`` and reload the host layer so that it binds to it."""
    mod = types.ModuleType("xarray")
    mod.DataArray, mod.Dataset, mod.broadcast, mod.merge = DataArray, Dataset, broadcast, merge
    mod.__xr_double__ = True
    _saved["xarray"] = sys.modules.get("xarray")
    sys.modules["xarray"] = mod
    import xarray_parcel_b200.parcel_functions as pf
    importlib.reload(pf)
    assert pf._xr is mod
    return pf


def uninstall():
    prev = _saved.pop("xarray", None)
    if prev is None:
        sys.modules.pop("xarray", None)
    else:
        sys.modules["xarray"] = prev
    import xarray_parcel_b200.parcel_functions as pf
    importlib.reload(pf)
