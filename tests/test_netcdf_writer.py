"""The in-tree netCDF-classic writer of the host driver (the image has no netCDF library) against an independent reader."""
import ctypes as C

import numpy as np
from scipy.io import netcdf_file

from openswpc_b200 import _lib


def test_cdf1_selftest_roundtrip(tmp_path):
    lib = _lib.load()
    lib.swpc3d_host_nc_selftest.argtypes = [C.c_char_p]
    p = tmp_path / "t.nc"
    assert lib.swpc3d_host_nc_selftest(str(p).encode()) == 0
    assert p.read_bytes()[:4] == b"CDF\x01"
    with netcdf_file(str(p), "r", mmap=False) as f:
        assert f.dimensions == {"x": 3, "z": 2, "t": None}
        assert f.generated_by == b"SWPC" and f.hdrver == 6 and np.float32(f.dt) == np.float32(0.25)
        np.testing.assert_array_equal(f.variables["x"][:], [0.5, 1.5, 2.5])
        assert f.variables["x"].units == b"km"
        np.testing.assert_array_equal(f.variables["rho"][:], [[1, 2, 3], [4, 5, 6]])
        assert f.variables["Vx"].dimensions == ("t", "z", "x") and f.variables["Vx"].shape == (2, 2, 3)
        np.testing.assert_array_equal(f.variables["t"][:], [0.0, 0.25])
        np.testing.assert_array_equal(f.variables["Vx"][1], [[10, 11, 12], [13, 14, -15]])
        np.testing.assert_array_equal(f.variables["Vx"].actual_range, [-15.0, 14.0])
