"""xarray_parcel_b200 -- B200-native column parcel lifting (drop-in for the hot path of
traupach/xarray_parcel's ``modules/parcel_functions.py``).

    import xarray_parcel_b200.parcel_functions as parcel
    parcel.load_moist_adiabat_lookups()
    cape_cin, profile = parcel.surface_based_cape_cin(pressure, temperature, dewpoint)

The compute path is hand-written CUDA for sm_100a behind a C ABI (include/xparcel.h,
xarray_parcel_b200/libxparcel.so).  There is no CPU fallback.
"""

from . import _build, _lib, synth  # noqa: F401
from . import parcel_functions  # noqa: F401

__version__ = "0.1.0"
