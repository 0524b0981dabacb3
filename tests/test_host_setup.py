"""The product's host-side setup chain (C++, include/swpc3d_host.h) against the oracle (C): integers bit-exact,
float arrays bit-exact (both evaluate the reference's expressions in the declared kinds with the same libm).
No GPU needed: swpc3d_host_create never touches the device."""
from pathlib import Path

import numpy as np
import pytest

from helpers import o_source_details, write_case
from openswpc_b200.swpc3d import Swpc3d, Swpc3dHostError
from oracle_lib import Oracle

INTS = ["ibeg", "iend", "jbeg", "jend", "nxp", "nyp", "ibeg_k", "iend_k", "jbeg_k", "jend_k", "kbeg_k", "kend_k", "nsrc", "nst"]

CASES = {
    "pml_lhm_ocean": dict(),
    "pml_lhm_land_3x2": dict(vmodel="lhm_land", nproc_x=3, nproc_y=2, nx=50, ny=44),
    "cerjan_uni_2x2": dict(abc_type="cerjan", vmodel="uni", nproc_x=2, nproc_y=2),
    "benchmark": dict(benchmark=True, nx=64, ny=64, nz=80, na=20),
    "bodyforce": dict(bf_mode=True, sources=["0.3 -0.2 4.1 0.05 0.6 1e12 2e12 -3e12"]),
    "dc_sources": dict(stf_format="xym0dc", stftype="herrmann", sources=["1.3 -0.7 5.2 0.0 0.8 2e15 30.0 45.0 90.0", "-1.0 2.0 3.3 0.2 0.4 1e15 210 80 -170"]),
    "mw_sources": dict(stf_format="xymwij", stftype="texp", sources=["1.3 -0.7 5.2 0.0 0.8 4.5 0.7 -0.3 0.5 0.4 -0.6 0.8"]),
    "dsdc_xy_2x2": dict(stf_format="xydsdc", nproc_x=2, nproc_y=2, sources=["1.3 -0.7 5.2 0.0 0.8 1.5 2.0e6 30.0 45.0 90.0", "-0.2 0.1 7.3 0.2 0.4 0.5 1e6 210 80 -170"]),
    "dsdc_ll": dict(stf_format="lldsdc", sources=["139.77 35.73 5.2 0.0 0.8 1.5 2.0e6 30.0 45.0 90.0"]),
    "psmeca": dict(stf_format="psmeca", sources=["139.765 35.715 5.0 1.21 -0.92 -0.29 0.56 -1.34 0.27 23", "139.75 35.70 6.5 -2.0 1.1 0.9 0.3 0.2 -0.7 22"]),
    "sdep_fit_bd0": dict(sdep_fit="bd0", sources=["0.3 -0.2 9.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"]),
    "ll_sources": dict(stf_format="llm0ij", stftype="cosine", sources=["139.77 35.73 5.2 0.0 0.8 2e15 0.7 -0.3 0.5 0.4 -0.6 0.8"]),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("nm", [3, 0])
def test_host_setup_matches_oracle(tmp_path, name, nm):
    kw = CASES[name]
    inf = write_case(tmp_path, nt=30, **kw)
    o = Oracle(inf, base_dir=tmp_path, nm=nm)
    for q in range(o.nranks):
        h = Swpc3d(inf, base_dir=tmp_path, nm=nm, myid=q)
        r = o.rank(q)
        for n in INTS:
            assert h[n] == r[n], (n, h[n], r[n])
        for n in ("fmax", "fcut", "M0", "UC", "zeta", "d2", "dt", "xbeg", "ybeg", "zbeg"):
            assert np.float32(h[n]) == np.float32(o.cfg(n)), n
        for n in ("rho", "lam", "mu", "taup", "taus"):
            np.testing.assert_array_equal(h.array(n), o.field(q, n).astype(np.float32), err_msg=n)
        j0, j1, i0, i1 = 3, 3 + r["nyp"], 3, 3 + r["nxp"]
        for n in ("kfs", "kob", "kbeg_a"):
            a, b = h.array(n), o.imap(q, n)
            # kfs/kob are defined on [ibeg-1, iend+2] (m_medium.f90:354-355); compare that window
            np.testing.assert_array_equal(a[j0 - 1:j1 + 2, i0 - 1:i1 + 2], b[j0 - 1:j1 + 2, i0 - 1:i1 + 2], err_msg=n)
        for n in ("kfs_top", "kfs_bot", "kob_top", "kob_bot"):
            np.testing.assert_array_equal(h.array(n)[j0:j1, i0:i1], o.imap(q, n)[j0:j1, i0:i1], err_msg=n)
        if o.cfg("abc_type") == "pml":
            for n in ("gxc", "gxe", "gyc", "gye", "gzc", "gze"):
                np.testing.assert_array_equal(h.array(n), o.profile(q, n), err_msg=n)
        else:
            for n in ("gx_c", "gx_b", "gy_c", "gy_b", "gz_c", "gz_b"):
                np.testing.assert_array_equal(h.array(n), o.profile(q, n), err_msg=n)
        ijk, mo = o.sources(q)
        np.testing.assert_array_equal(h.array("src_ijk"), ijk)
        np.testing.assert_array_equal(h.array("mo"), mo)
        if len(mo):
            mij, prm = o_source_details(o, q)
            np.testing.assert_array_equal(h.array("mij"), mij)
            np.testing.assert_array_equal(h.array("srcprm"), prm)
        sijk, names = o.stations(q)
        np.testing.assert_array_equal(h.array("st_ijk"), sijk)
        assert h.station_names() == names
        if nm > 0:
            np.testing.assert_array_equal(h.array("ts"), o.ts())
            for cf in ("c1", "c2", "d1"):
                np.testing.assert_array_equal(h.array(cf), o.coef(cf))
        h.close()
    # the allreduce of vmin/vmax (m_medium.f90:424-425) is the caller's: min/max over ranks == oracle's global values
    hs = [Swpc3d(inf, base_dir=tmp_path, nm=nm, myid=q) for q in range(o.nranks)]
    assert np.float32(min(h["vmin_local"] for h in hs)) == np.float32(o.cfg("vmin"))
    assert np.float32(max(h["vmax_local"] for h in hs)) == np.float32(o.cfg("vmax"))


def test_example_input_header_values():
    """example/example.out:9-13 (the reference's own known answers) through the product's setup chain."""
    ref = Path("/root/reference")
    if not ref.exists():
        pytest.skip("reference tree not present (GPU box)")
    h = Swpc3d(ref / "example" / "input.inf", base_dir=ref, nm=3, myid=0)
    assert f"{h['c']:.3f}" == "0.645"
    assert f"{h['r']:.3f}" == "12.488"
    assert f"{h['vmin']:.3f}" == "3.122"
    assert f"{h['vmax']:.3f}" == "7.977"
    assert f"{h['fmax']:.3f}" == "0.500"
    assert (h["ibeg"], h["iend"], h["jbeg"], h["jend"]) == (1, 192, 1, 192)
    np.testing.assert_array_equal(h.array("src_ijk"), [[192, 192, 24]])
    np.testing.assert_array_equal(h.array("st_ijk"), [[192, 192, 21], [172, 182, 21]])
    assert h["ntw"] == 200


def test_error_behaviour(tmp_path):
    inf = write_case(tmp_path, nt=10, extra="vmodel_type = 'user'")
    # the first matching key wins (m_readini.f90:78-92): the case's own vmodel_type line comes first
    Swpc3d(inf, base_dir=tmp_path, nm=3).close()
    bad = tmp_path / "bad.inf"
    bad.write_text(inf.read_text().replace("vmodel_type = 'lhm'", "vmodel_type = 'user'"))
    with pytest.raises(Swpc3dHostError, match="vmodel_type"):
        Swpc3d(bad, base_dir=tmp_path, nm=3)
    with pytest.raises(Swpc3dHostError, match="cannot open parameter file"):
        Swpc3d(tmp_path / "nope.inf", base_dir=tmp_path)
    with pytest.raises(Swpc3dHostError, match="myid"):
        Swpc3d(inf, base_dir=tmp_path, nm=3, myid=7)
    out = tmp_path / "outside.inf"
    (tmp_path / "far.dat").write_text("500.0 0.0 4.0 0.1 1.0 1e15 1 1 1 0 0 0\n")
    out.write_text(inf.read_text().replace('fn_stf = "source.dat"', 'fn_stf = "far.dat"'))
    h = Swpc3d(out, base_dir=tmp_path, nm=3)   # far outside of every sleeve: simply not owned
    assert h["nsrc"] == 0
