"""Host side of the swpc_psv path: the C-ABI library exports every declared symbol, and the C++ setup chain (psv_driver.cpp)
reproduces the oracle's setup arrays exactly -- integers bit-exact, float arrays bit-exact (same kinds, same order)."""
import re
from pathlib import Path

import numpy as np
import pytest

from psv_oracle import MAPS, MEDIUM, PsvOracle, psv_case_text, write_psv_files

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from openswpc_b200 import _lib, swpc_psv

    lib = _lib.load()
    for hdr, names in (("swpcpsv_b200.h", _lib.PSV_SYMBOLS), ("swpcpsv_host.h", swpc_psv.HOST_SYMBOLS)):
        declared = set(re.findall(r"\b(swpcpsv_\w+)\s*\(", (ROOT / "include" / hdr).read_text()))
        assert declared == set(names), declared ^ set(names)
        for s in names:
            assert hasattr(lib, s), s


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from openswpc_b200._lib import Swpc3dError
    from openswpc_b200.psv_device import PsvGeometry, PsvRank

    with pytest.raises(Swpc3dError, match="no CUDA device"):
        PsvRank(PsvGeometry(nx=64, nz=64, nproc_x=1, myid=0, ibeg=1, iend=64, ibeg_k=11, iend_k=54, kend_k=54, na=10), dx=0.5, dz=0.5, dt=0.01, nm=0,
                abc_type="pml")


@pytest.mark.parametrize("case", [dict(), dict(abc="cerjan", nm=0), dict(nproc_x=3, nx=100), dict(bf_mode=True, sources=["0.3 0.0 4.2 0.05 0.6 3e9 0.0 -4e9"]),
                                  dict(stf_format="xym0dc", sources=["0.3 0.0 4.2 0.05 0.6 1e15 30 60 -80", "-5 0 7 0.1 0.5 2e14 120 40 10"]),
                                  dict(stf_format="llmwij", sources=["139.77 35.70 4.2 0.05 0.6 4.5 0.7 0.0 -0.3 0.0 0.5 0.0"], vmodel="lhm", extra=" fn_lhm = 'lhm.dat'"),
                                  dict(sp=True)])
def test_host_setup_matches_oracle(tmp_path, case):
    from openswpc_b200.swpc_psv import SwpcPsv

    case = dict(case)
    nm, sp, sources = case.pop("nm", 3), case.pop("sp", False), case.pop("sources", None)
    write_psv_files(tmp_path, sources=sources)
    (tmp_path / "lhm.dat").write_text("# depth rho vp vs qp qs\n0.0 2.3 5.5 3.14 600 300\n3.0 2.4 6.0 3.55 600 300\n16.0 2.8 6.7 3.83 600 300\n")
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=40, products="v,u", **case))
    o = PsvOracle(inf, base_dir=tmp_path, nm=nm, sp=sp)
    for q in range(o.nranks):
        h = SwpcPsv(inf, base_dir=tmp_path, nm=nm, myid=q, field_dtype=np.float32 if sp else np.float64)
        r = o.rank(q)
        for n in ("ibeg", "iend", "ibeg_k", "iend_k", "kend_k", "nxp"):
            assert h[n] == r[n], n
        assert h["nsrc"] == r["nsrc"] and h["nst"] == r["nst"] and h["ntw"] == o.cfg("ntw")
        for n in ("fcut", "fmax", "M0", "UC", "zeta", "d2", "vmin_local" if o.nranks > 1 else "vmin"):
            if n == "vmin_local":
                continue
            assert np.float32(h[n]) == np.float32(o.cfg(n)), (n, h[n], o.cfg(n))
        for n in MEDIUM:
            assert np.array_equal(h[n].reshape(o.shape2(q)), o.field(q, n).astype(np.float32)), n
        for n in MAPS:
            assert np.array_equal(h[n], o.map(q, n)), n
        names = ("gxc", "gxe", "gzc", "gze") if o.cfg("abc_type") == "pml" else ("gx_c", "gx_b", "gz_c", "gz_b")
        for n in names:
            assert np.array_equal(h[n], o.profile(q, n)), n
        ik, val = o.sources(q)
        assert np.array_equal(h["src_ik"].reshape(-1, 2), ik)
        if len(ik):
            assert np.array_equal(h["mo"], val[:, 0])
            assert np.array_equal(h["m3"].reshape(-1, 3)[:, :2], val[:, 1:3])
            if not o.cfg("bf_mode"):
                assert np.array_equal(h["m3"].reshape(-1, 3)[:, 2], val[:, 3])
            assert np.array_equal(h["srcprm"].reshape(-1, 2), val[:, 4:6].astype(np.float32))
        sik, snames = o.stations(q)
        assert np.array_equal(h["st_ik"].reshape(-1, 2), sik) and h.station_names() == snames
        assert np.array_equal(h["ts"], o.cfg("ts"))


def test_scope_errors_are_explicit(tmp_path):
    from openswpc_b200.swpc_psv import SwpcPsv, SwpcPsvError

    write_psv_files(tmp_path)
    for vm, msg in (("user", "vmodel_type"), ("grd", "no layer in the list"), ("lgm", "not found")):
        inf = tmp_path / "input.inf"
        inf.write_text(psv_case_text(nt=4, vmodel=vm))
        with pytest.raises(SwpcPsvError, match=msg):
            SwpcPsv(inf, base_dir=tmp_path)


@pytest.mark.parametrize("ps,dip", [("p", 0.0), ("s", 25.0), ("P", -30.0)])
def test_planewave_initial_condition_matches_oracle(tmp_path, ps, dip):
    """pw_setup (swpc_psv/m_source.f90:656-785): the five initial fields over the memory box, fcut / fmax / M0, no source grid."""
    from openswpc_b200.swpc_psv import SwpcPsv

    write_psv_files(tmp_path)
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=10, nproc_x=2, nx=100, extra=f" pw_mode = .true.\n pw_ztop = 12.0\n pw_zlen = 6.0\n pw_ps = '{ps}'\n pw_dip = {dip}\n pw_strike = 33.0\n pw_rake = 12.0"))
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    for q in range(o.nranks):
        h = SwpcPsv(inf, base_dir=tmp_path, nm=3, myid=q)
        assert h["nsrc"] == 0 == o.rank(q)["nsrc"]
        for n in ("fcut", "fmax", "M0"):
            assert np.float32(h[n]) == np.float32(o.cfg(n)), n
        for n in ("Vx", "Vz", "Sxx", "Szz", "Sxz"):
            a, ref = h["init_" + n].reshape(o.shape2(q)), o.field(q, n)
            assert np.array_equal(a, ref), n
        assert np.abs(o.field(q, "Vz")).max() > 0 and np.abs(o.field(q, "Szz")).max() > 0
        for n in ("gxc", "gze"):
            assert np.array_equal(h[n], o.profile(q, n)), n


PSV_LAYERS = "# depth rho vp vs qp qs [rmed]\n0.5 2.3 5.5 3.14 600 300 r1.nc\n3.0 2.4 6.0 3.55 400 200 r2.nc\n9.0 2.1 1.2 0.6 100 50 r1.nc\n12.0 2.8 6.7 3.83 600 300 none.nc\n"
PSV_MODELS = {
    "lgm": " fn_lhm = 'layers.dat'\n",
    "lgm_flat": " fn_lhm = 'layers.dat'\n earth_flattening = .true.\n munk_profile = .true.\n",
    "uni_rmed": " dir_rmed = '.'\n fn_rmed0 = 'r0.nc'\n rhomin = 1.0\n",
    "uni_rmed_nofile": " dir_rmed = '.'\n fn_rmed0 = 'nope.nc'\n",
    "lhm_rmed": " fn_lhm_rmed = 'layers.dat'\n dir_rmed = '.'\n rhomin = 2.0\n",
    "lhm_rmed_flat": " fn_lhm_rmed = 'layers.dat'\n dir_rmed = '.'\n earth_flattening = .true.\n",
    "lgm_rmed": " fn_lhm_rmed = 'layers.dat'\n dir_rmed = '.'\n",
}


@pytest.mark.parametrize("name", list(PSV_MODELS))
@pytest.mark.parametrize("nproc_x", [1, 3])
def test_psv_model_builders_match_oracle(tmp_path, name, nproc_x):
    """swpc_psv/m_vmodel_{lgm,uni_rmed,lhm_rmed,lgm_rmed}.f90 + rdrmed__2d: the product's C++ builders against the oracle's, cell
    for cell.  r2.nc is strong enough for the vmax / vmin / rhomin corrections of vcheck; the third layer is below vcut and
    takes the parameters of the one beneath it; none.nc does not exist (no perturbation)."""
    from helpers import write_rmed2d
    from openswpc_b200.swpc_psv import SwpcPsv

    rng = np.random.default_rng(11)
    for fn, shape, amp in (("r0.nc", (30, 41), 0.08), ("r1.nc", (24, 20), 0.05), ("r2.nc", (150, 130), 0.6)):
        write_rmed2d(tmp_path / fn, (amp * rng.standard_normal(shape)).astype(np.float32))
    write_psv_files(tmp_path)
    (tmp_path / "layers.dat").write_text(PSV_LAYERS)
    vt = name.replace("_flat", "").replace("_nofile", "")
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=20, nx=100, nz=90, nproc_x=nproc_x, vmodel=vt, zbeg=-3.0, extra=PSV_MODELS[name] + " vcut = 1.5\n"))
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    for q in range(o.nranks):
        h = SwpcPsv(inf, base_dir=tmp_path, nm=3, myid=q)
        for n in MEDIUM:
            assert np.array_equal(h[n].reshape(o.shape2(q)), o.field(q, n).astype(np.float32)), (q, n)
        for n in MAPS:
            assert np.array_equal(h[n], o.map(q, n)), (q, n)
        assert np.float32(h["zeta"]) == np.float32(o.cfg("zeta"))
        if o.nranks == 1:
            assert np.float32(h["vmin"]) == np.float32(o.cfg("vmin")) and np.float32(h["vmax"]) == np.float32(o.cfg("vmax"))
    if name in ("uni_rmed", "lhm_rmed", "lhm_rmed_flat"):   # (lgm_rmed assigns whole planes, m_vmodel_lgm_rmed.f90:225-229: laterally uniform)
        mu, na = o.field(0, "mu"), o.cfg("na")                # (nxm, nzm)
        assert np.ptp(mu[na + 5:-na - 5], axis=0).max() > 0  # laterally varying inside the absorber


def test_rdrmed2d_wraps_like_the_reference(tmp_path):
    """rdrmed__2d (m_rdrmed.f90:19-70): periodic in x, k <= 0 wraps upward, rows below the section repeat it cyclically."""
    import ctypes as C

    from helpers import write_rmed2d
    from oracle_lib import lib as load_oracle

    lib = load_oracle()
    nzc, nxc = 7, 5
    xi = np.arange(nzc * nxc, dtype=np.float32).reshape(nzc, nxc) + 1
    write_rmed2d(tmp_path / "s.nc", xi)
    ib, ie, kb, ke = -3, 12, -2, 17
    vol = np.zeros((ie - ib + 1, ke - kb + 1), dtype=np.float32)
    err = C.create_string_buffer(512)
    lib.ora_rdrmed2d.argtypes = [C.c_int] * 4 + [C.c_char_p, C.POINTER(C.c_float), C.c_char_p, C.c_size_t]
    assert lib.ora_rdrmed2d(ib, ie, kb, ke, str(tmp_path / "s.nc").encode(), vol.ctypes.data_as(C.POINTER(C.c_float)), err, 512) == 0, err.value
    wrap = lambda v, n: np.where(v % n <= 0, v % n + n, v % n)
    ii = wrap(np.arange(ib, ie + 1), nxc) - 1
    for k in range(kb, ke + 1):
        kk = k + nzc if k <= 0 else (k if k <= nzc else int(wrap(np.array(k), nzc)))
        np.testing.assert_array_equal(vol[:, k - kb], xi[kk - 1][ii], err_msg=str(k))


@pytest.mark.parametrize("vt,opts", [("grd", ""), ("grd", " is_ocean = .false.\n"), ("grd", " topo_flatten = .true.\n earth_flattening = .true.\n"),
                                     ("grd_rmed", " rhomin = 2.2\n")])
@pytest.mark.parametrize("nproc_x", [1, 2])
def test_psv_grd_models_match_oracle(tmp_path, vt, opts, nproc_x):
    """swpc_psv/m_vmodel_grd.f90 / m_vmodel_grd_rmed.f90: GMT grids sampled along the section (y = 0, no clamping to the absorber
    edge), layers filled below the interpolated interfaces; the random-media variant reads 2-D sections relative to a
    reference interface.  sdep_fit = bd1 takes the source depth from the layer flagged pid = 1."""
    from helpers import write_grd, write_rmed2d
    from openswpc_b200.swpc_psv import SwpcPsv

    lon = 139.40 + 0.01 * np.arange(72)
    lat = 35.50 + 0.01 * np.arange(46)
    LO, LA = np.meshgrid(lon, lat)
    write_grd(tmp_path / "g1.grd", lon, lat, 600.0 * np.sin((LO - 139.76) * 40.0) * np.cos((LA - 35.72) * 35.0) + 150.0)
    write_grd(tmp_path / "g2.grd", lon, lat, 3200.0 + 900.0 * np.cos((LO - 139.7) * 25.0) + 400.0 * np.sin((LA - 35.7) * 30.0), zdtype=">f8")
    write_grd(tmp_path / "g3.grd", lon, lat, 9500.0 + 1500.0 * np.sin((LO - 139.8) * 12.0 + (LA - 35.7) * 9.0))
    rng = np.random.default_rng(5)
    for fn, shape, amp in (("r1.nc", (24, 20), 0.05), ("r2.nc", (150, 130), 0.6)):
        write_rmed2d(tmp_path / fn, (amp * rng.standard_normal(shape)).astype(np.float32))
    if vt == "grd":
        (tmp_path / "grd.lst").write_text("# file rho vp vs qp qs pid\n'g1.grd' 2.1 2.4 1.0 100 50 0\n'g2.grd'  2.5 5.0 2.9 300 150 0\n g3.grd  2.9 6.8 3.9 500 250 1\n")
        extra = " fn_grdlst = 'grd.lst'\n dir_grd = '.'\n"
    else:
        (tmp_path / "grd.lst").write_text("'g1.grd' 2.1 2.4 1.0 100 50 0 'r1.nc' 0\n'g2.grd'  2.5 5.0 2.9 300 150 0 'r2.nc' 2\n g3.grd  2.9 6.8 3.9 500 250 1 none.nc 1\n")
        extra = " fn_grdlst_rmed = 'grd.lst'\n dir_grd = '.'\n dir_rmed = '.'\n"
    write_psv_files(tmp_path, sources=["0.3 0.0 4.2 0.05 0.6 1e15 0.7 0.0 -0.3 0.0 0.5 0.0", "-6.0 0.0 1.0 0.1 0.5 4e14 0.2 0.0 0.9 0.0 0.1 0.0"])
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=20, nx=100, nz=90, nproc_x=nproc_x, vmodel=vt, zbeg=-2.0, extra=extra + opts + " vcut = 1.5\n sdep_fit = 'bd1'\n phi = 90.0\n"))
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    for q in range(o.nranks):
        h = SwpcPsv(inf, base_dir=tmp_path, nm=3, myid=q)
        for n in MEDIUM:
            assert np.array_equal(h[n].reshape(o.shape2(q)), o.field(q, n).astype(np.float32)), (q, n)
        for n in MAPS:
            assert np.array_equal(h[n], o.map(q, n)), (q, n)
        ik, _ = o.sources(q)
        assert np.array_equal(h["src_ik"].reshape(-1, 2), ik)
        sik, _ = o.stations(q)
        assert np.array_equal(h["st_ik"].reshape(-1, 2), sik)
    kob = np.concatenate([o.map(q, "kob")[3:-3] for q in range(o.nranks)])
    if "topo_flatten" not in opts:
        assert kob.max() > kob.min()          # real bathymetry along the section


@pytest.mark.parametrize("nproc_x", [1, 2])
def test_psv_stabilize_pml(tmp_path, nproc_x):
    """stabilize_absorber of swpc_psv (m_medium.f90:309-366): thin low-velocity layers inside the absorber are replaced by the
    material above them and shear velocities are floored at 0.4 vmax (the GLOBAL maximum: on several ranks the product
    applies it once the caller has reduced vmin / vmax, as the reference does after its all-reduce)."""
    from openswpc_b200.swpc_psv import SwpcPsv

    write_psv_files(tmp_path)
    (tmp_path / "lvz.dat").write_text("# depth rho vp vs qp qs\n0.0 2.3 5.5 3.14 600 300\n3.0 2.4 6.0 3.55 400 200\n6.0 2.2 4.0 2.10 200 100\n"
                                      "8.0 2.8 6.7 3.83 600 300\n15.0 3.2 7.8 4.46 600 300\n")
    text = {sw: psv_case_text(nt=10, nx=100, nz=90, nproc_x=nproc_x, vmodel="lhm", zbeg=-3.0, extra=f" fn_lhm = 'lvz.dat'\n stabilize_pml = {sw}\n")
            for sw in (".true.", ".false.")}
    inf = tmp_path / "input.inf"
    inf.write_text(text[".true."])
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    plain = PsvOracle(None, base_dir=tmp_path, nm=3, text=text[".false."])
    changed = 0
    for q in range(o.nranks):
        h = SwpcPsv(inf, base_dir=tmp_path, nm=3, myid=q)
        if o.nranks > 1:
            h.set_minmax(o.cfg("vmin"), o.cfg("vmax"))
        for n in MEDIUM:
            assert np.array_equal(h[n].reshape(o.shape2(q)), o.field(q, n).astype(np.float32)), (q, n)
        changed += int((o.field(q, "mu") != plain.field(q, "mu")).sum())
    assert changed > 0
