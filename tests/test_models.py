"""the heavier velocity-model builders (lgm, uni_rmed, lhm_rmed, lgm_rmed, stabilize_pml) -- the product's
host-side C++ (openswpc_b200/csrc/host/models.hpp) against the oracle's C restatement (oracle/ora_models.c), bit for bit,
and the random-media reader against an independent numpy statement of m_rdrmed.f90:73-134.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from helpers import write_case, write_grd, write_rmed
from openswpc_b200.swpc3d import Swpc3d
from oracle_lib import Oracle

LAYERS_RMED = """# depth rho vp vs Qp Qs rmed
  0.0   2.3   5.5   3.14   600   300  'r1.nc'
  3.0   2.4   6.0   3.55   400   200  r2.nc
  9.0   2.8   6.7   3.83   600   300  r1.nc
 15.0   3.2   7.8   4.46   600   300  missing.nc
"""
# a thin low-velocity layer for the stabiliser to remove
LAYERS_LVZ = """# depth rho vp vs Qp Qs
  0.0   2.3   5.5   3.14   600   300
  3.0   2.4   6.0   3.55   400   200
  6.0   2.2   4.0   2.10   200   100
  8.0   2.8   6.7   3.83   600   300
 15.0   3.2   7.8   4.46   600   300
"""


def _volumes(d, seed=7):
    rng = np.random.default_rng(seed)
    xs = {}
    for name, shape, amp in (("r1.nc", (20, 12, 16), 0.05), ("r2.nc", (60, 50, 70), 0.6), ("r0.nc", (24, 16, 20), 0.08)):
        xs[name] = (amp * rng.standard_normal(shape)).astype(np.float32)
        write_rmed(d / name, xs[name])
    return xs


CASES = {
    "lgm": "vmodel_type = 'lgm'\n fn_lhm = 'lhm_land.dat'\n",
    "lgm_ocean_flat": "vmodel_type = 'lgm'\n fn_lhm = 'lhm_ocean.dat'\n earth_flattening = .true.\n",
    "lgm_stabilize": "vmodel_type = 'lgm'\n fn_lhm = 'lvz.dat'\n stabilize_pml = .true.\n",
    "lhm_stabilize": "vmodel_type = 'lhm'\n fn_lhm = 'lvz.dat'\n stabilize_pml = .true.\n",
    "uni_rmed": "vmodel_type = 'uni_rmed'\n vp0 = 5.0\n vs0 = 2.9\n rho0 = 2.6\n qp0 = 300\n qs0 = 150\n topo0 = 0.4\n dir_rmed = '.'\n fn_rmed0 = 'r0.nc'\n rhomin = 1.0\n",
    "uni_rmed_nofile": "vmodel_type = 'uni_rmed'\n vp0 = 5.0\n dir_rmed = '.'\n fn_rmed0 = 'nope.nc'\n",
    "lhm_rmed": "vmodel_type = 'lhm_rmed'\n fn_lhm_rmed = 'layers_rmed.dat'\n dir_rmed = '.'\n rhomin = 2.0\n",
    "lgm_rmed": "vmodel_type = 'lgm_rmed'\n fn_lhm_rmed = 'layers_rmed.dat'\n dir_rmed = '.'\n",
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("ranks", [(1, 1), (2, 2)])
def test_model_builders_match_oracle(tmp_path, name, ranks):
    _volumes(tmp_path)
    (tmp_path / "layers_rmed.dat").write_text(LAYERS_RMED)
    (tmp_path / "lvz.dat").write_text(LAYERS_LVZ)
    inf = write_case(tmp_path, nt=10, vmodel="raw:" + CASES[name], nproc_x=ranks[0], nproc_y=ranks[1], nx=52, ny=44, nz=48, extra="vcut = 1.5")
    inf.write_text(inf.read_text().replace(" vcut = 0.0\n", " vcut = 1.5\n"))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    hs = [Swpc3d(inf, base_dir=tmp_path, nm=3, myid=q) for q in range(o.nranks)]
    vmin, vmax = min(h["vmin_local"] for h in hs), max(h["vmax_local"] for h in hs)
    assert np.float32(vmin) == np.float32(o.cfg("vmin")) and np.float32(vmax) == np.float32(o.cfg("vmax"))
    for q, h in enumerate(hs):
        h.set_minmax(vmin, vmax)   # the allreduce of m_medium.f90:424-425; triggers the deferred stabiliser on multi-rank runs
        for n in ("rho", "lam", "mu", "taup", "taus"):
            np.testing.assert_array_equal(h.array(n), o.field(q, n).astype(np.float32), err_msg=f"{name} rank {q} {n}")
        r = o.rank(q)
        j0, j1, i0, i1 = 3, 3 + r["nyp"], 3, 3 + r["nxp"]
        for n in ("kfs", "kob"):
            np.testing.assert_array_equal(h.array(n)[j0 - 1:j1 + 2, i0 - 1:i1 + 2], o.imap(q, n)[j0 - 1:j1 + 2, i0 - 1:i1 + 2], err_msg=n)
        for n in ("kfs_top", "kfs_bot", "kob_top", "kob_bot"):
            np.testing.assert_array_equal(h.array(n)[j0:j1, i0:i1], o.imap(q, n)[j0:j1, i0:i1], err_msg=n)
        h.close()
    if name in ("uni_rmed", "lhm_rmed"):   # the perturbation really is lateral
        rho = o.field(0, "rho")
        assert np.ptp(rho[10:-10, 10:-10, 30]) > 0
    if name.endswith("stabilize"):        # and the stabiliser really changed the absorber
        inf2 = tmp_path / "nostab.inf"
        inf2.write_text(inf.read_text().replace("stabilize_pml = .true.", "stabilize_pml = .false."))
        o2 = Oracle(inf2, base_dir=tmp_path, nm=3)
        assert not np.array_equal(o.field(0, "mu"), o2.field(0, "mu"))
        na = 6
        np.testing.assert_array_equal(o.field(0, "mu")[3 + na + 3:-(3 + na + 3), 3 + na + 3:-(3 + na + 3), :48 - na - 3],
                                      o2.field(0, "mu")[3 + na + 3:-(3 + na + 3), 3 + na + 3:-(3 + na + 3), :48 - na - 3])


def test_rdrmed3d_cyclic_read(tmp_path):
    """m_rdrmed.f90:73-134 restated in numpy: periodic in x / y, k <= 0 wraps upward, planes below nzc repeat mod nzc."""
    xs = _volumes(tmp_path)
    xi = xs["r1.nc"]            # (nz, ny, nx) = (20, 12, 16)
    nzc, nyc, nxc = xi.shape
    ib, ie, jb, je, kb, ke = -2, 40, -2, 30, -2, 47
    lib = oracle_lib.lib("dp")
    vol = np.zeros((je - jb + 1, ie - ib + 1, ke - kb + 1), dtype=np.float32)
    err = C.create_string_buffer(512)
    lib.ora_rdrmed3d.argtypes = [C.c_int] * 6 + [C.c_char_p, C.POINTER(C.c_float), C.c_char_p, C.c_size_t]
    rc = lib.ora_rdrmed3d(ib, ie, jb, je, kb, ke, str(tmp_path / "r1.nc").encode(), vol.ctypes.data_as(C.POINTER(C.c_float)), err, 512)
    assert rc == 0, err.value
    wrap = lambda v, n: np.where(np.fmod(v, n) <= 0, np.fmod(v, n) + n, np.fmod(v, n))
    ii = wrap(np.arange(ib, ie + 1), nxc) - 1
    jj = wrap(np.arange(jb, je + 1), nyc) - 1
    exp = np.zeros_like(vol)
    for k in range(kb, ke + 1):
        if k <= nzc:
            kk = k + nzc if k <= 0 else k
            exp[:, :, k - kb] = xi[kk - 1][np.ix_(jj, ii)]
        else:
            exp[:, :, k - kb] = exp[:, :, (k % nzc) - kb]
    np.testing.assert_array_equal(vol, exp)
    rc = lib.ora_rdrmed3d(ib, ie, jb, je, kb, ke, str(tmp_path / "lhm_land.dat").encode(), vol.ctypes.data_as(C.POINTER(C.c_float)), err, 512)
    assert rc != 0


def _grids(d):
    """Three interfaces on a 0.01-degree grid around the model centre: a surface with land and sea, and two deeper ones."""
    lon = 139.40 + 0.01 * np.arange(72)
    lat = 35.50 + 0.01 * np.arange(46)
    LO, LA = np.meshgrid(lon, lat)
    surf = 600.0 * np.sin((LO - 139.76) * 40.0) * np.cos((LA - 35.72) * 35.0) + 150.0       # m, positive down: sea where > 0
    mid = 3200.0 + 900.0 * np.cos((LO - 139.7) * 25.0) + 400.0 * np.sin((LA - 35.7) * 30.0)
    deep = 9500.0 + 1500.0 * np.sin((LO - 139.8) * 12.0 + (LA - 35.7) * 9.0)
    write_grd(d / "g1.grd", lon, lat, surf)
    write_grd(d / "g2.grd", lon, lat, mid, zdtype=">f8")
    write_grd(d / "g3.grd", lon, lat, deep)
    (d / "grd.lst").write_text("# file rho vp vs qp qs pid\n'g1.grd' 2.1 2.4 1.0 100 50 0\n'g2.grd'  2.5 5.0 2.9 300 150 0\n g3.grd  2.9 6.8 3.9 500 250 1\n")
    return lon, lat, surf


@pytest.mark.parametrize("ranks", [(1, 1), (2, 2)])
@pytest.mark.parametrize("opts", ["", " is_ocean = .false.\n", " topo_flatten = .true.\n earth_flattening = .true.\n"])
def test_vmodel_grd_matches_oracle(tmp_path, ranks, opts):
    """vmodel_grd (m_vmodel_grd.f90) + m_bicubic: classic-netCDF grids, bicubic interpolation onto the columns, layer filling."""
    _grids(tmp_path)
    vm = "vmodel_type = 'grd'\n fn_grdlst = 'grd.lst'\n dir_grd = '.'\n" + opts
    inf = write_case(tmp_path, nt=10, vmodel="raw:" + vm, nproc_x=ranks[0], nproc_y=ranks[1], nx=52, ny=44, nz=48, zbeg=-2.0, sdep_fit="bd1",
                     stations=["0.0 0.0 0.0 st01 obb", "-3.2 2.1 2.0 st02 dep", "4.3 -3.9 0.0 st03 fsb", "2.2 4.4 0.0 st04 bd1"])
    inf.write_text(inf.read_text().replace(" vcut = 0.0\n", " vcut = 1.5\n"))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    hs = [Swpc3d(inf, base_dir=tmp_path, nm=3, myid=q) for q in range(o.nranks)]
    for q, h in enumerate(hs):
        for n in ("rho", "lam", "mu", "taup", "taus"):
            np.testing.assert_array_equal(h.array(n), o.field(q, n).astype(np.float32), err_msg=f"rank {q} {n}")
        r = o.rank(q)
        j0, j1, i0, i1 = 3, 3 + r["nyp"], 3, 3 + r["nxp"]
        for n in ("kfs", "kob"):
            np.testing.assert_array_equal(h.array(n)[j0 - 1:j1 + 2, i0 - 1:i1 + 2], o.imap(q, n)[j0 - 1:j1 + 2, i0 - 1:i1 + 2], err_msg=n)
        np.testing.assert_array_equal(h.array("src_ijk"), o.sources(q)[0])     # sdep_fit = bd1 uses the plate-boundary depth of layer 3
        np.testing.assert_array_equal(h.array("st_ijk"), o.stations(q)[0])
        h.close()
    kfs, kob = o.imap(0, "kfs")[4:-4, 4:-4], o.imap(0, "kob")[4:-4, 4:-4]
    if "topo_flatten" not in opts:
        assert kob.max() > kob.min()                              # real topography / bathymetry
    else:
        assert kob.max() == kob.min()                             # topo_flatten: the first interface is moved to z = 0 everywhere
    if "is_ocean = .false." not in opts and "topo_flatten" not in opts:
        assert (kob > kfs).any() and (kob == kfs).any()            # sea columns and land columns


@pytest.mark.parametrize("ranks", [(1, 1), (2, 2)])
@pytest.mark.parametrize("opts", ["", " earth_flattening = .true.\n rhomin = 2.2\n"])
def test_vmodel_grd_rmed_matches_oracle(tmp_path, ranks, opts):
    """vmodel_grd_rmed (m_vmodel_grd_rmed.f90): the grd layers, each perturbed by a random-media volume sampled at the depth
    below its reference interface (reflyr; 0 = the top of the volume, 2 = below the second grid), cyclic in depth.  r2.nc is
    strong enough to trigger the vmax / vmin / rhomin corrections of vcheck; 'none.nc' does not exist (no perturbation)."""
    _grids(tmp_path)
    xs = _volumes(tmp_path)
    (tmp_path / "grd_rmed.lst").write_text("# file rho vp vs qp qs pid rmed reflyr\n'g1.grd' 2.1 2.4 1.0 100 50 0 'r1.nc' 0\n"
                                           "'g2.grd'  2.5 5.0 2.9 300 150 0 'r2.nc' 2\n g3.grd  2.9 6.8 3.9 500 250 1 none.nc 1\n")
    vm = "vmodel_type = 'grd_rmed'\n fn_grdlst_rmed = 'grd_rmed.lst'\n dir_grd = '.'\n dir_rmed = '.'\n" + opts
    inf = write_case(tmp_path, nt=10, vmodel="raw:" + vm, nproc_x=ranks[0], nproc_y=ranks[1], nx=52, ny=44, nz=48, zbeg=-2.0)
    inf.write_text(inf.read_text().replace(" vcut = 0.0\n", " vcut = 1.5\n"))
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    hs = [Swpc3d(inf, base_dir=tmp_path, nm=3, myid=q) for q in range(o.nranks)]
    for q, h in enumerate(hs):
        for n in ("rho", "lam", "mu", "taup", "taus"):
            np.testing.assert_array_equal(h.array(n), o.field(q, n).astype(np.float32), err_msg=f"rank {q} {n}")
        for n in ("kfs", "kob"):
            r = o.rank(q)
            np.testing.assert_array_equal(h.array(n)[2:4 + r["nyp"] + 1, 2:4 + r["nxp"] + 1], o.imap(q, n)[2:4 + r["nyp"] + 1, 2:4 + r["nxp"] + 1], err_msg=n)
        h.close()
    # the perturbation is really there: the same case without random media differs in the first two layers only
    (tmp_path / "plain").mkdir()
    _grids(tmp_path / "plain")
    vm0 = "vmodel_type = 'grd'\n fn_grdlst = 'grd.lst'\n dir_grd = '.'\n" + opts
    inf0 = write_case(tmp_path / "plain", nt=10, vmodel="raw:" + vm0, nproc_x=ranks[0], nproc_y=ranks[1], nx=52, ny=44, nz=48, zbeg=-2.0)
    inf0.write_text(inf0.read_text().replace(" vcut = 0.0\n", " vcut = 1.5\n"))
    o0 = Oracle(inf0, base_dir=tmp_path / "plain", nm=3)
    a, b = o.field(0, "rho"), o0.field(0, "rho")
    changed = a != b
    assert changed.any() and not changed[b == np.float32(2.9)].any() and not changed[b < 1.5].any()
    if "rhomin" in opts:
        assert a[changed].min() >= np.float32(2.2)


def test_bicubic_reproduces_a_bicubic_surface(tmp_path):
    """The patch interpolates exactly (to rounding) any surface whose values, first and cross derivatives it samples exactly:
    a product of quadratics has central differences equal to its derivatives, so the interpolant must return it."""
    lon = 139.0 + 0.05 * np.arange(30)
    lat = 35.0 + 0.04 * np.arange(25)
    LO, LA = np.meshgrid(lon, lat)
    f = lambda x, y: 1000.0 * (1 + 0.5 * (x - 139.7) ** 2) * (2 - 0.3 * (y - 35.5) ** 2)
    write_grd(tmp_path / "g1.grd", lon, lat, f(LO, LA), zdtype=">f8")
    (tmp_path / "grd.lst").write_text("g1.grd 2.5 5.0 2.9 300 150 1\n")
    vm = "vmodel_type = 'grd'\n fn_grdlst = 'grd.lst'\n dir_grd = '.'\n"
    inf = write_case(tmp_path, nt=5, vmodel="raw:" + vm, nx=52, ny=44, nz=48, zbeg=-2.0, na=6)
    h = Swpc3d(inf, base_dir=tmp_path, nm=3)
    # bd(:,:,1) holds the interpolated depth [km] of the layer with pid = 1 at every column: recover it through sdep_fit = bd1
    # sources are not needed -- compare the layer's top index with the analytic surface instead
    mu = h.array("mu")
    nym, nxm, nzm = mu.shape
    xc, yc, zc = h.array("xc"), h.array("yc"), h.array("zc")
    import ctypes as C
    lib = oracle_lib.lib("dp")
    top = np.argmax(mu > 0, axis=2)          # first solid k per column (0-based into zc)
    bad = 0
    for j in range(3 + 6, nym - 3 - 6, 3):
        for i in range(3 + 6, nxm - 3 - 6, 3):
            lo, la = C.c_float(), C.c_float()
            lib.ora_geomap_c2g(C.c_float(xc[i]), C.c_float(yc[j]), C.c_float(139.7604), C.c_float(35.7182), C.c_float(0.0), C.byref(lo), C.byref(la))
            z = f(lo.value, la.value) / 1000.0
            # the first solid cell is the first one whose centre lies deeper than the interface (within one cell)
            bad += not (zc[top[j, i]] - 0.51 <= z <= zc[top[j, i]] + 0.51)
    assert bad == 0
