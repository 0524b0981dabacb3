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
    for extra, msg in ((" vmodel_type = 'grd'", "vmodel_type"),):
        inf = tmp_path / "input.inf"
        inf.write_text(psv_case_text(nt=4, extra=extra) if "vmodel" not in extra else psv_case_text(nt=4, vmodel="grd"))
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
