"""the heavier velocity-model builders (lgm, uni_rmed, lhm_rmed, lgm_rmed, stabilize_pml) -- the product's
host-side C++ (openswpc_b200/csrc/host/models.hpp) against the oracle's C restatement (oracle/ora_models.c), bit for bit,
and the random-media reader against an independent numpy statement of m_rdrmed.f90:73-134.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from helpers import write_case, write_rmed
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
