"""Snapshot products (m_snap.f90, snp_format='netcdf'): the files the product's driver writes from the device slices,
read back with an independent netCDF reader and compared record by record, bit for bit, with the oracle."""
import numpy as np
import pytest

import oracle_lib as OL
from helpers import write_case
from snap_helpers import check_file, snap_extra
from openswpc_b200.swpc3d import Swpc3d
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("dec", [(2, 2, 2, 5), (1, 1, 1, 4), (3, 2, 4, 7)])
def test_all_snapshot_products(tmp_path, dec):
    nt = 42
    inf = write_case(tmp_path, nt=nt, title="snp", extra=snap_extra(*dec))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    run.snap_open(tmp_path / "gpu")
    run.run(1, nt)
    run.snap_close()
    for q in range(15):
        sec, typ = divmod(q, 3)
        p = tmp_path / "gpu" / f"snp.3d.{OL.SNAP_SECTIONS[sec]}.{OL.SNAP_TYPES[typ]}.nc"
        assert p.exists(), p.name
        check_file(p, o, q, "snp", run["dt"], dec[3])


def test_snapshot_subset_and_cerjan(tmp_path):
    nt = 30
    inf = write_case(tmp_path, nt=nt, title="sub", abc_type="cerjan", vmodel="lhm_land",
                     extra=snap_extra(sw="xz_v%sw = .true.\n fs_u%sw = .true.\n ob_ps%sw = .true."))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    run.snap_open(tmp_path / "g")
    run.run(1, nt)
    run.snap_close()
    names = sorted(p.name for p in (tmp_path / "g").glob("*.nc"))
    assert names == ["sub.3d.fs.u.nc", "sub.3d.ob.ps.nc", "sub.3d.xz.v.nc"]
    for q in (1 * 3 + 1, 3 * 3 + 2, 4 * 3 + 0):
        sec, typ = divmod(q, 3)
        check_file(tmp_path / "g" / f"sub.3d.{OL.SNAP_SECTIONS[sec]}.{OL.SNAP_TYPES[typ]}.nc", o, q, "sub", run["dt"], 5)


def native_expected(o, q, title, exedate, dt, ntdec_s, na, dec, clon, clat, phi):
    """write_snp_header (m_snap.f90:846-892) + medium arrays + records, from the oracle's slices"""
    import struct

    sec, typ = divmod(q, 3)
    horiz = sec in (0, 3, 4)
    n1, n2, nv = OL.snap_dims(o, q)
    x, y, z = OL.snap_coords(o)
    a1 = y if sec == 2 else x
    a2 = z if sec in (1, 2) else y
    e1 = dec[1] if sec == 2 else dec[0]
    e2 = dec[2] if sec in (1, 2) else dec[1]
    f32 = np.float32
    b = b"STREAMIO" + b"SWPC_3D " + struct.pack("<i", 6) + title.ljust(80).encode() + struct.pack("<i", exedate)
    b += OL.SNAP_SECTIONS[sec].encode() + [b"ps", b"v3", b"u3"][typ]
    b += struct.pack("<ii", n1, n2) + np.array([a1[0], a2[0], a1[1] - a1[0], a2[1] - a2[0], f32(dt) * f32(ntdec_s)], dtype="<f4").tobytes()
    b += struct.pack("<iiii", na // e1, na // e2, 6 if horiz else 3, nv)
    b += np.array([clon, clat, phi, 1, 1, 1], dtype="<f4").tobytes()
    for m in range(4 if horiz else 3):
        b += OL.snap_medium(o, q, m).astype("<f4").tobytes()
    recs, _ = OL.snap_records(o, q)
    return b + recs.astype("<f4").tobytes()


def test_native_snapshot_files(tmp_path):
    nt, dec = 21, (2, 3, 2)
    inf = write_case(tmp_path, nt=nt, title="nat", extra=snap_extra(*dec, 5).replace("'netcdf'", "'native'"))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.set_exedate(1_700_000_000, 540)
    run.attach_device(0)
    run.snap_open(tmp_path / "g")
    run.run(1, nt)
    run.snap_close()
    for q in range(15):
        sec, typ = divmod(q, 3)
        got = (tmp_path / "g" / f"nat.3d.{OL.SNAP_SECTIONS[sec]}.{OL.SNAP_TYPES[typ]}.snp").read_bytes()
        ref = native_expected(o, q, "nat", 1_700_000_000, run["dt"], 5, run["na"], dec, run["clon"], run["clat"], run["phi"])
        assert got == ref, (q, len(got), len(ref))


def test_unknown_snapshot_format_refused(tmp_path):
    inf = write_case(tmp_path, nt=4, extra="snp_format = 'hdf5'\n xy_v%sw = .true.")
    with pytest.raises(Exception, match="snp_format"):
        Swpc3d(inf, base_dir=tmp_path, nm=3)


def test_snapshots_follow_the_topography(tmp_path):
    """The fs / ob products sample k = kfs(i,j) + 1 / kob(i,j) + 1 (m_snap.f90:1016-1036, 1476-1478): on a GMT-grid model with
    relief those depths change from column to column, with land and sea side by side; all 15 products."""
    from helpers import write_grd

    lon = 139.40 + 0.01 * np.arange(72)
    lat = 35.50 + 0.01 * np.arange(46)
    LO, LA = np.meshgrid(lon, lat)
    write_grd(tmp_path / "g1.grd", lon, lat, 700.0 * np.sin((LO - 139.76) * 60.0) * np.cos((LA - 35.72) * 55.0) + 100.0)
    write_grd(tmp_path / "g2.grd", lon, lat, 3200.0 + 900.0 * np.cos((LO - 139.7) * 25.0) + 400.0 * np.sin((LA - 35.7) * 30.0))
    (tmp_path / "grd.lst").write_text("'g1.grd' 2.1 3.0 1.6 100 50 0\n'g2.grd'  2.5 5.0 2.9 300 150 1\n")
    vm = "vmodel_type = 'grd'\n fn_grdlst = 'grd.lst'\n dir_grd = '.'\n"
    nt = 40
    dec = (2, 2, 2, 4)
    inf = write_case(tmp_path, nt=nt, title="topo", nx=52, ny=44, nz=48, zbeg=-2.0, dt=0.01, vmodel="raw:" + vm,
                     sources=["0.3 -0.2 2.1 0.02 0.3 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"], extra=snap_extra(*dec))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    kfs, kob = o.imap(0, "kfs")[4:-4, 4:-4], o.imap(0, "kob")[4:-4, 4:-4]
    assert kob.max() - kob.min() >= 2 and (kob > kfs).any() and (kob == kfs).any()
    o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    run.snap_open(tmp_path / "gpu")
    run.run(1, nt)
    run.snap_close()
    for q in range(15):
        sec, typ = divmod(q, 3)
        check_file(tmp_path / "gpu" / f"topo.3d.{OL.SNAP_SECTIONS[sec]}.{OL.SNAP_TYPES[typ]}.nc", o, q, "topo", run["dt"], dec[3])
