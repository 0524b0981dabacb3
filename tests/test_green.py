"""Green's-function mode (m_green.f90): host setup against the oracle without a GPU; on the GPU, the traces
and the exported SAC files against the oracle, bit for bit."""
import numpy as np
import pytest

from helpers import write_case
from openswpc_b200.swpc3d import Swpc3d, Swpc3dHostError
from oracle_lib import Oracle

GLST = """# x y z gid
 1.3  -0.7  5.2  17
-3.1   2.4  2.6  4
 4.4   3.9  9.0  123456
 0.2   0.1  0.3  5
 9.9  -8.8  4.0  6
 30.0  0.0  4.0  7
"""


def _case(d, nt=40, ranks=(1, 1), cmp="z", bforce=False, extra="", stnm="st02"):
    (d / "glst.xyz").write_text(GLST)
    ex = (f"green_mode = .true.\n green_stnm = '{stnm}'\n green_cmp = '{cmp}'\n green_trise = 0.6\n green_bforce = {'.true.' if bforce else '.false.'}\n"
          f" fn_glst = 'glst.xyz'\n green_fmt = 'xyz'\n green_maxdist = 12.0\n" + extra)
    return write_case(d, nt=nt, nproc_x=ranks[0], nproc_y=ranks[1], nx=52, ny=44, extra=ex, title="gr")


def _hosts(inf, d, n):
    hs = [Swpc3d(inf, base_dir=d, nm=3, myid=q) for q in range(n)]
    if n > 1:   # the broadcast of m_green.f90:161-183, done here by hand
        owner = [h.green_query() for h in hs]
        src = [q for q in owner if q[0]][-1]
        for h in hs:
            h.green_set_source(*src[1:])
    return hs


@pytest.mark.parametrize("ranks", [(1, 1), (2, 2)])
def test_green_setup_matches_oracle(tmp_path, ranks):
    inf = _case(tmp_path, ranks=ranks)
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    hs = _hosts(inf, tmp_path, o.nranks)
    total = 0
    for q, h in enumerate(hs):
        g = o.green(q)
        assert h["green_mode"] == 1 and h["nsrc"] == 0 and h["ng"] == g["ng"] and h["green_ncmp"] == g["ncmp"] == 6 and h["green_ntw"] == g["ntw"]
        np.testing.assert_array_equal(h.array("green_ijk"), g["ijk"])
        np.testing.assert_array_equal(h.array("green_gid"), g["gid"])
        assert np.float32(h["M0"]) == np.float32(o.cfg("M0")) == 1 and np.float32(h["fmax"]) == np.float32(o.cfg("fmax"))
        for n in ("gxc", "gze"):   # absorb__setup ran with fcut of the (absent) source grid
            np.testing.assert_array_equal(h.array(n), o.profile(q, n))
        total += g["ng"]
    # gid 5 is in the air/sea column above kob, gid 6 farther than green_maxdist, gid 7 outside the model
    assert total == 3
    found, ijk, _, _ = hs[-1].green_query() if o.nranks == 1 else [h.green_query() for h in hs if h.green_query()[0]][0]
    g0 = o.green(0)
    assert found and ijk == [g0["isrc"], g0["jsrc"], g0["ksrc"]]


def test_green_errors(tmp_path):
    inf = _case(tmp_path, stnm="nowhere")
    with pytest.raises(Swpc3dHostError, match="green_stnm"):
        Swpc3d(inf, base_dir=tmp_path, nm=3)
    bad = tmp_path / "bad.inf"
    bad.write_text(inf.read_text().replace("green_cmp = 'z'", "green_cmp = 'q'").replace("nowhere", "st02"))
    with pytest.raises(Swpc3dHostError, match="green_cmp"):
        Swpc3d(bad, base_dir=tmp_path, nm=3)


@pytest.mark.gpu
@pytest.mark.parametrize("cmp,bforce,fmt", [("z", False, "sac"), ("x", True, "sac"), ("y", True, "csf")])
def test_green_traces_and_files_match_oracle(tmp_path, cmp, bforce, fmt):
    nt = 60
    inf = _case(tmp_path, nt=nt, cmp=cmp, bforce=bforce, extra=f"wav_format = '{fmt}'\n")
    inf.write_text(inf.read_text().replace(" wav_format = 'sac'\n", f" wav_format = '{fmt}'\n", 1))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.lib.ora_set_exedate(o.h, 1_700_000_000, 540)
    vm_ref = o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.set_exedate(1_700_000_000, 540)
    run.attach_device(0)
    vm = run.run(1, nt)
    np.testing.assert_array_equal(vm, vm_ref)
    got, nz = run.download_fields(), run["nz"]
    for f, a in got.items():
        np.testing.assert_array_equal(a[3:-3, 3:-3, 3:3 + nz], o.field(0, f)[3:-3, 3:-3, 3:3 + nz], err_msg=f)
    n = run.write_green(tmp_path / "gpu")
    g = o.green(0)
    gf = run.array("green_gf")
    ref = -g["gf"] if cmp == "z" else g["gf"]      # green__export flips the z component (m_green.f90:563-565)
    assert np.abs(ref).max() > 0
    np.testing.assert_array_equal(gf, ref)
    if fmt == "sac":
        n_ref = o.write_green_sac(tmp_path / "ref")
        assert n == n_ref == g["ng"] * g["ncmp"] and g["ncmp"] == (9 if bforce else 6)
        for f in sorted((tmp_path / "ref" / "green" / "st02").glob("*.sac")):
            h = tmp_path / "gpu" / "green" / "st02" / f.name
            assert h.exists(), f.name
            assert h.read_bytes() == f.read_bytes(), f.name
    else:
        files = list((tmp_path / "gpu" / "green" / "st02").glob("*.csf"))
        assert n == 1 and len(files) == 1 and files[0].name == "gr__st02__y__000000__.csf"
        raw = files[0].read_bytes()
        ntr, npts = np.frombuffer(raw[4:12], dtype=np.int32)
        assert raw[:4] == b"CSFD" and ntr == g["ng"] * g["ncmp"] and npts == g["ntw"] and len(raw) == 12 + ntr * (632 + 4 * npts)
        np.testing.assert_array_equal(np.frombuffer(raw[12 + 632:12 + 632 + 4 * npts], dtype=np.float32), ref[0])
