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


def test_snapshot_native_format_refused(tmp_path):
    inf = write_case(tmp_path, nt=4, extra="snp_format = 'native'\n xy_v%sw = .true.")
    with pytest.raises(Exception, match="snp_format"):
        Swpc3d(inf, base_dir=tmp_path, nm=3)
