"""Shared by the snapshot tests: the input.inf snapshot block and the file-vs-oracle comparison."""
import numpy as np
from scipy.io import netcdf_file

import oracle_lib as OL

ALL_ON = "\n ".join(f"{s}_{t}%sw = .true." for s in OL.SNAP_SECTIONS for t in OL.SNAP_TYPES)


def snap_extra(idec=2, jdec=2, kdec=2, ntdec_s=5, sw=ALL_ON, more=""):
    return (f"snp_format = 'netcdf'\n {sw}\n idec = {idec}\n jdec = {jdec}\n kdec = {kdec}\n ntdec_s = {ntdec_s}\n"
            f" z0_xy = 4.0\n x0_yz = 1.0\n y0_xz = -1.5\n {more}")


def check_file(path, o, q, title, dt, ntdec_s):
    sec, typ = divmod(q, 3)
    horiz = sec in (0, 3, 4)
    recs, its = OL.snap_records(o, q)
    n1, n2, nv = OL.snap_dims(o, q)
    x, y, z = OL.snap_coords(o)
    c1 = y if sec == 2 else x
    c2 = z if sec in (1, 2) else y
    d1, d2 = ("y" if sec == 2 else "x"), ("z" if sec in (1, 2) else "y")
    vnames = {0: ["div", "rot_x", "rot_y", "rot_z"], 1: ["Vx", "Vy", "Vz"], 2: ["Ux", "Uy", "Uz"]}[typ]
    with netcdf_file(str(path), "r", mmap=False) as f:
        assert f.dimensions == {d1: n1, d2: n2, "t": None}
        assert f.generated_by == b"SWPC" and f.hdrver == 6 and f.title == title.encode()
        assert f.coordinate == OL.SNAP_SECTIONS[sec].encode() and f.datatype == [b"ps", b"v3", b"u3"][typ]
        assert f.ns1 == n1 and f.ns2 == n2 and f.nsnp == nv and f.nmed == (6 if horiz else 3)
        assert np.float32(f.dt) == np.float32(dt) * np.float32(ntdec_s)
        np.testing.assert_array_equal(f.variables[d1][:], c1)
        np.testing.assert_array_equal(f.variables[d2][:], c2)
        for m, name in enumerate(["rho", "lambda", "mu"] + (["topo", "lon", "lat"] if horiz else [])):
            np.testing.assert_array_equal(f.variables[name][:], OL.snap_medium(o, q, m), err_msg=f"{path.name}:{name}")
        assert len(its) == f.variables["t"].shape[0] > 1
        np.testing.assert_array_equal(f.variables["t"][:], np.array([np.float32(it) * np.float32(dt) for it in its], dtype=np.float32))
        for v, name in enumerate(vnames):
            var = f.variables[name]
            assert var.dimensions == ("t", d2, d1)
            np.testing.assert_array_equal(var[:], recs[:, v], err_msg=f"{path.name}:{name}")
            np.testing.assert_array_equal(var.actual_range, [min(recs[:, v].min(), 0), max(recs[:, v].max(), 0)])
        assert np.abs(recs).max() > 0, path.name
        mx = OL.snap_max(o, q)
        if sec in (3, 4) and typ != 0:
            for v, name in enumerate(["max-V", "max-H", "max-A"]):
                np.testing.assert_array_equal(f.variables[name][:], mx[v], err_msg=f"{path.name}:{name}")
            assert mx.max() > 0
        else:
            assert "max-V" not in f.variables
