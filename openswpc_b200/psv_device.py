"""Thin object wrapper over the swpc_psv C ABI (include/swpcpsv_b200.h): one `PsvRank` == one MPI rank of the reference's
swpc_psv == one GPU-resident strip of columns.  Method names follow the reference's subroutines.

numpy conventions: a reference array `A(kbeg_m:kend_m, ibeg_m:iend_m)` (k fastest) is a C-ordered numpy array of shape
(nxm, nzm); a map `M(ibeg_m:iend_m)` has shape (nxm,).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import PsvGrid, check_psv as check

FIELDS = ("Vx", "Vz", "Sxx", "Szz", "Sxz")
ABC = {"pml": 1, "cerjan": 2}


@dataclass
class PsvGeometry:
    """The integers of src/swpc_psv/m_global.f90:41-69 for one rank."""
    nx: int
    nz: int
    nproc_x: int
    myid: int
    ibeg: int
    iend: int
    ibeg_k: int
    iend_k: int
    kend_k: int
    na: int
    ipad: int = 0
    kpad: int = 0

    @property
    def nxp(self):
        return self.iend - self.ibeg + 1

    @property
    def shape2(self):
        return (self.nxp + 6 + self.ipad, self.nz + 6 + self.kpad)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PsvRank:
    def __init__(self, geom: PsvGeometry, *, dx: float, dz: float, dt: float, nm: int, abc_type: str, ts=None,
                 field_dtype=np.float64, device: int = -1):
        self.lib = _lib.load()
        self.geom = geom
        self.dtype = np.dtype(field_dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("field_dtype must be float64 (MP=DP) or float32 (MP=SP)")
        g = PsvGrid()
        for n in ("nx", "nz", "nproc_x", "myid", "ibeg", "iend", "ipad", "kpad", "ibeg_k", "iend_k", "kend_k", "na"):
            setattr(g, n, int(getattr(geom, n)))
        g.nm = nm
        g.abc_type = ABC[abc_type]
        g.field_bytes = self.dtype.itemsize
        g.device = device
        g.dx, g.dz, g.dt = float(dx), float(dz), float(np.float32(dt))
        self.nm = nm
        self.abc_type = abc_type
        tsa = np.ascontiguousarray(ts if ts is not None else np.zeros(max(nm, 1)), dtype=np.float32)
        h = C.c_void_p()
        check(self.lib.swpcpsv_create(C.byref(g), _fp(tsa), C.byref(h)))
        self.h = h
        self.ntw = 0
        self.nst = 0
        self.sw = (0, 0, 0, 0)

    def close(self):
        if getattr(self, "h", None):
            self.lib.swpcpsv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads (main.f90:80-93)
    def upload_medium(self, rho, lam, mu, taup, taus, kfs, kob, kfs_top, kfs_bot, kob_top, kob_bot, kbeg_a=None):
        s2 = self.geom.shape2
        f = [np.ascontiguousarray(a, dtype=np.float32) for a in (rho, lam, mu, taup, taus)]
        m = [np.ascontiguousarray(a, dtype=np.int32) for a in (kfs, kob, kfs_top, kfs_bot, kob_top, kob_bot)]
        for a in f:
            assert a.shape == s2, (a.shape, s2)
        for a in m:
            assert a.shape == (s2[0],), (a.shape, s2)
        ka = None if kbeg_a is None else np.ascontiguousarray(kbeg_a, dtype=np.int32)
        check(self.lib.swpcpsv_upload_medium(self.h, *[_fp(a) for a in f], *[_ip(a) for a in m], _ip(ka) if ka is not None else None))

    def upload_fields(self, **fields):
        args = []
        keep = []
        for n in FIELDS:
            a = fields.get(n)
            if a is None:
                args.append(None)
            else:
                a = np.ascontiguousarray(a, dtype=self.dtype)
                assert a.shape == self.geom.shape2
                keep.append(a)
                args.append(a.ctypes.data_as(C.c_void_p))
        check(self.lib.swpcpsv_upload_fields(self.h, *args))

    def download_fields(self) -> dict:
        out = {n: np.empty(self.geom.shape2, dtype=self.dtype) for n in FIELDS}
        check(self.lib.swpcpsv_download_fields(self.h, *[out[n].ctypes.data_as(C.c_void_p) for n in FIELDS]))
        return out

    def download_memvars(self) -> dict:
        """Rxx, Rzz, Rxz in the reference layout (m fastest): shape (nxm, nzm, nm)."""
        if self.nm == 0:
            return {}
        out = {n: np.zeros(self.geom.shape2 + (self.nm,), dtype=np.float32) for n in ("Rxx", "Rzz", "Rxz")}
        check(self.lib.swpcpsv_download_memvars(self.h, *[_fp(out[n]) for n in ("Rxx", "Rzz", "Rxz")]))
        return out

    def zero_state(self):
        check(self.lib.swpcpsv_zero_state(self.h))

    def setup_pml(self, gxc, gxe, gzc, gze):
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (gxc, gxe, gzc, gze)]
        assert a[0].size == 4 * self.geom.nxp and a[2].size == 4 * self.geom.nz
        check(self.lib.swpcpsv_setup_pml(self.h, *[_fp(x) for x in a]))

    def setup_cerjan(self, gx_c, gx_b, gz_c, gz_b):
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (gx_c, gx_b, gz_c, gz_b)]
        assert a[0].size == self.geom.shape2[0] and a[2].size == self.geom.shape2[1]
        check(self.lib.swpcpsv_setup_cerjan(self.h, *[_fp(x) for x in a]))

    def set_sources(self, isrc, ksrc, mo, mxx, mzz, mxz, srcprm, stftype="kupper", bf_mode=False, tbeg=0.0):
        n = len(isrc)
        i = np.ascontiguousarray(isrc, dtype=np.int32)
        k = np.ascontiguousarray(ksrc, dtype=np.int32)
        d = [np.ascontiguousarray(a if a is not None else np.zeros(n), dtype=np.float64) for a in (mo, mxx, mzz, mxz)]
        prm = np.ascontiguousarray(srcprm, dtype=np.float32).reshape(-1)
        check(self.lib.swpcpsv_set_sources(self.h, n, _ip(i), _ip(k), *[_dp(a) for a in d], _fp(prm), stftype.encode(), int(bool(bf_mode)),
                                           C.c_float(tbeg)))

    def set_stations(self, ist, kst, ntdec_w, ntw, M0, UC, sw_v=True, sw_u=False, sw_stress=False, sw_strain=False):
        n = len(ist)
        i = np.ascontiguousarray(ist, dtype=np.int32)
        k = np.ascontiguousarray(kst, dtype=np.int32)
        self.nst, self.ntw = n, ntw
        self.sw = (int(sw_v), int(sw_u), int(sw_stress), int(sw_strain))
        check(self.lib.swpcpsv_set_stations(self.h, n, _ip(i), _ip(k), int(ntdec_w), int(ntw), C.c_float(M0), C.c_float(UC), *self.sw))

    def get_wav(self, which: int = 0) -> np.ndarray:
        """which: 0 velocity, 1 displacement (nst, 2, ntw); 2 stress, 3 strain (nst, 3, ntw)."""
        ncmp = 2 if which < 2 else 3
        out = np.zeros((self.nst, ncmp, self.ntw), dtype=np.float32)
        if self.nst and self.sw[which]:
            check(self.lib.swpcpsv_get_wav(self.h, which, _fp(out)))
        return out

    # ---- the hot path (main.f90:95-113)
    def update_stress(self):
        check(self.lib.swpcpsv_update_stress(self.h))

    def stressglut(self, it: int):
        check(self.lib.swpcpsv_stressglut(self.h, it))

    def comm_stress(self):
        check(self.lib.swpcpsv_comm_stress(self.h))

    def update_vel(self, it: int):
        check(self.lib.swpcpsv_update_vel(self.h, it))

    def comm_vel(self):
        check(self.lib.swpcpsv_comm_vel(self.h))

    def wav_store(self, it: int):
        check(self.lib.swpcpsv_wav_store(self.h, it))

    def step(self, it: int):
        check(self.lib.swpcpsv_step(self.h, it))

    def run(self, it0: int, it1: int):
        check(self.lib.swpcpsv_run(self.h, it0, it1))

    def sync(self):
        check(self.lib.swpcpsv_sync(self.h))

    def vmax(self, global_: bool = False) -> np.ndarray:
        out = np.zeros(2, dtype=np.float32)
        check((self.lib.swpcpsv_vmax_global if global_ else self.lib.swpcpsv_vmax)(self.h, _fp(out)))
        return out

    def comm_init(self, unique_id: bytes, nranks: int, rank: int):
        check(self.lib.swpcpsv_comm_init(self.h, unique_id, nranks, rank))

    def timer_start(self):
        check(self.lib.swpcpsv_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(self.lib.swpcpsv_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def set_option(self, key: str, value: int):
        check(self.lib.swpcpsv_set_option(self.h, key.encode(), int(value)))

    def info(self, key: str) -> float:
        v = C.c_double()
        check(self.lib.swpcpsv_get_info(self.h, key.encode(), C.byref(v)))
        return float(v.value)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_lib.load().swpcpsv_nccl_unique_id(buf))
    return buf.raw


def comm_local(ranks, which: int):
    """Emulated halo exchange between ranks living in this process (tests); which: 0 stress, 1 velocity."""
    arr = (C.c_void_p * len(ranks))(*[r.h for r in ranks])
    check(_lib.load().swpcpsv_comm_local(arr, len(ranks), which))


def step_local(ranks, it: int):
    """One iteration of main.f90:99-111 for several ranks on this process's GPU(s)."""
    for r in ranks:
        r.wav_store(it)
        r.update_stress()
        r.stressglut(it)
    comm_local(ranks, 0)
    for r in ranks:
        r.update_vel(it)
    comm_local(ranks, 1)
