"""Plane-wave mode (pw_mode, m_source.f90:316-466): the host driver's initial condition against the oracle's (CPU), and
the device run with the PML edge extrapolation (m_absorb_p.f90:137-243, :332-424) against the oracle (GPU)."""
import numpy as np
import pytest

from helpers import write_case
from openswpc_b200.swpc3d import Swpc3d, Swpc3dHostError
from oracle_lib import Oracle

FIELDS = ("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")


def pw_extra(ps="p", strike=20.0, dip=15.0, rake=40.0, ztop=6.0, zlen=4.0):
    return (f"pw_mode = .true.\n pw_ztop = {ztop}\n pw_zlen = {zlen}\n pw_ps = '{ps}'\n pw_strike = {strike}\n pw_dip = {dip}\n"
            f" pw_rake = {rake}")


@pytest.mark.parametrize("ps", ["p", "S"])
@pytest.mark.parametrize("npxy", [(1, 1), (2, 2)])
def test_planewave_initial_condition(tmp_path, ps, npxy):
    inf = write_case(tmp_path, nt=10, nproc_x=npxy[0], nproc_y=npxy[1], vmodel="lhm_land", stftype="cosine", extra=pw_extra(ps))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    for q in range(o.nranks):
        h = Swpc3d(inf, base_dir=tmp_path, nm=3, myid=q)
        assert h["nsrc"] == 0
        for n in ("M0", "fcut", "fmax"):
            assert np.float32(h[n]) == np.float32(o.cfg(n)), n
        for f in FIELDS:
            a, b = h.array("init_" + f), o.field(q, f)
            np.testing.assert_array_equal(a, b, err_msg=f)
            assert np.abs(b).max() > 0, f


def test_planewave_asserts(tmp_path):
    for extra, msg in ((pw_extra(zlen=-1.0), "pw_zlen"), (pw_extra(ps="x"), "pw_ps"), (pw_extra(ztop=1e4), "pw_ztop")):
        inf = write_case(tmp_path, nt=4, extra=extra)
        with pytest.raises(Swpc3dHostError, match=msg):
            Swpc3d(inf, base_dir=tmp_path, nm=3)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(ps="p", nm=3), dict(ps="s", nm=0, abc_type="cerjan"), dict(ps="s", nm=3, field_dtype=np.float32)])
def test_planewave_run_matches_oracle(tmp_path, case):
    nt = 40
    fd = case.get("field_dtype", np.float64)
    inf = write_case(tmp_path, nt=nt, vmodel="lhm_land", abc_type=case.get("abc_type", "pml"), extra=pw_extra(case["ps"]))
    o = Oracle(inf, base_dir=tmp_path, nm=case["nm"], mp="sp" if fd == np.float32 else "dp")
    vm_ref = o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=case["nm"], field_dtype=fd)
    run.attach_device(0)
    vm = run.run(1, nt)
    np.testing.assert_array_equal(vm, vm_ref)
    assert vm.max() > 0
    got = run.download_fields()
    nz = run["nz"]
    for f in FIELDS:
        ref = o.field(0, f)
        np.testing.assert_array_equal(got[f][:, :, 3:3 + nz], ref[:, :, 3:3 + nz].astype(fd), err_msg=f)
    assert run["nst"] > 0 and run.write_sac(tmp_path / "gpu") == 3 * run["nst"]
    np.testing.assert_array_equal(run.wav(), o.wav(0))
