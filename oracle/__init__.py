"""CPU oracle for the column parcel-lifting path of traupach/xarray_parcel.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``xarray_parcel_b200``) imports, links or executes anything under
``oracle/``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
there only as the checker / the timed CPU baseline.

What it is: a NumPy (float64, whole-array) restatement of
``/root/reference/modules/parcel_functions.py`` (PF) for the hot path named in
``BASELINE.json`` (LCL -> parcel profile -> LCL insertion -> LFC/EL -> CAPE/CIN
for surface-based, mixed-layer and most-unstable parcels), written from the
reference's xarray formulation operation by operation, every function citing
the PF lines it follows.

Third-party arithmetic: the reference delegates its thermodynamics to MetPy
(PyPI ``metpy``; not vendored in /root/reference and not installed here; the
reference pins it only through notebook print-outs: MetPy 1.4.1 and 1.6.2).
``oracle/thermo.py`` restates the published MetPy formulas (Bolton 1980 etc.).

Pinning status
--------------
* exact-ODE mode (``moist_lapse='ode'``): PINNED against every known-answer
  test the reference ships for this path (``modules/unit_tests.py``; vectors
  extracted to ``tests/golden/ut_soundings.json`` by
  ``tests/golden/extract_ut_soundings.py``; checked in
  ``tests/test_oracle_kat.py``).
* lookup-table mode (``moist_lapse='lut'``, the mode the reference runs in
  production and the mode the CUDA path implements): the reference's own pins
  for it are ``test_data.nc`` + ``historic_results/*.nc``, which are absent
  from /root/reference (``.MISSING_LARGE_BLOBS``) and unreadable here (no
  HDF5).  Beyond ``unit_tests.py:106-112,166-188`` (2 decimals) the LUT mode
  is therefore **parity unpinned**: the oracle in LUT mode *is* the reference
  for GPU parity.
"""
