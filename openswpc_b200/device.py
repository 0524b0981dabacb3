"""Thin object wrapper over the C ABI (include/swpc3d_b200.h): one `DeviceRank` == one MPI rank of the
reference == one GPU-resident subdomain.  Method names follow the reference's subroutines
(kernel__update_stress -> update_stress, global__comm_vel -> comm_vel, ...).

numpy conventions: a reference array `A(kbeg_m:kend_m, ibeg_m:iend_m, jbeg_m:jend_m)` (k fastest) is a
C-ordered numpy array of shape (nym, nxm, nzm); a map `M(ibeg_m:iend_m, jbeg_m:jend_m)` has shape (nym, nxm).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Grid, check

FIELDS = ("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")
ABC = {"pml": 1, "cerjan": 2}


@dataclass
class RankGeometry:
    """The integers of m_global.f90:52-86 for one rank."""
    nx: int
    ny: int
    nz: int
    nproc_x: int
    nproc_y: int
    myid: int
    ibeg: int
    iend: int
    jbeg: int
    jend: int
    ibeg_k: int
    iend_k: int
    jbeg_k: int
    jend_k: int
    kbeg_k: int
    kend_k: int
    na: int
    ipad: int = 0
    jpad: int = 0
    kpad: int = 0

    @property
    def nxp(self):
        return self.iend - self.ibeg + 1

    @property
    def nyp(self):
        return self.jend - self.jbeg + 1

    @property
    def shape3(self):
        return (self.nyp + 6 + self.jpad, self.nxp + 6 + self.ipad, self.nz + 6 + self.kpad)

    @property
    def shape2(self):
        return (self.nyp + 6 + self.jpad, self.nxp + 6 + self.ipad)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class DeviceRank:
    def __init__(self, geom: RankGeometry, *, dx: float, dy: float, dz: float, dt: float, nm: int, abc_type: str,
                 ts=None, field_dtype=np.float64, device: int = -1):
        self.lib = _lib.load()
        self.geom = geom
        self.dtype = np.dtype(field_dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("field_dtype must be float64 (MP=DP) or float32 (MP=SP)")
        g = Grid()
        for n in ("nx", "ny", "nz", "nproc_x", "nproc_y", "myid", "ibeg", "iend", "jbeg", "jend", "ipad", "jpad", "kpad",
                  "ibeg_k", "iend_k", "jbeg_k", "jend_k", "kbeg_k", "kend_k", "na"):
            setattr(g, n, int(getattr(geom, n)))
        g.nm = nm
        g.abc_type = ABC[abc_type]
        g.field_bytes = self.dtype.itemsize
        g.device = device
        g.dx, g.dy, g.dz, g.dt = float(dx), float(dy), float(dz), float(np.float32(dt))
        self.nm = nm
        self.abc_type = abc_type
        tsa = np.ascontiguousarray(ts if ts is not None else np.zeros(max(nm, 1)), dtype=np.float32)
        h = C.c_void_p()
        check(self.lib.swpc3d_create(C.byref(g), _fp(tsa), C.byref(h)))
        self.h = h
        self.ntw = 0
        self.nst = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.swpc3d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads (main.f90:80-113)
    def upload_medium(self, rho, lam, mu, taup, taus, kfs, kob, kfs_top, kfs_bot, kob_top, kob_bot, kbeg_a=None):
        s3, s2 = self.geom.shape3, self.geom.shape2
        f = [np.ascontiguousarray(a, dtype=np.float32) for a in (rho, lam, mu, taup, taus)]
        m = [np.ascontiguousarray(a, dtype=np.int32) for a in (kfs, kob, kfs_top, kfs_bot, kob_top, kob_bot)]
        for a in f:
            assert a.shape == s3, (a.shape, s3)
        for a in m:
            assert a.shape == s2, (a.shape, s2)
        ka = np.ascontiguousarray(kbeg_a, dtype=np.int32) if kbeg_a is not None else None
        check(self.lib.swpc3d_upload_medium(self.h, *map(_fp, f), *map(_ip, m), _ip(ka) if ka is not None else None))

    def upload_fields(self, **fields):
        args = []
        keep = []
        for n in FIELDS:
            a = fields.get(n)
            if a is None:
                args.append(None)
            else:
                a = np.ascontiguousarray(a, dtype=self.dtype)
                assert a.shape == self.geom.shape3
                keep.append(a)
                args.append(a.ctypes.data_as(C.c_void_p))
        check(self.lib.swpc3d_upload_fields(self.h, *args))

    def download_fields(self, names=FIELDS) -> dict:
        out = {n: np.zeros(self.geom.shape3, dtype=self.dtype) for n in names}
        args = [out[n].ctypes.data_as(C.c_void_p) if n in out else None for n in FIELDS]
        check(self.lib.swpc3d_download_fields(self.h, *args))
        return out

    def zero_state(self):
        check(self.lib.swpc3d_zero_state(self.h))

    def setup_pml(self, gxc, gxe, gyc, gye, gzc, gze):
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (gxc, gxe, gyc, gye, gzc, gze)]
        g = self.geom
        for x, n in zip(a, (g.nxp, g.nxp, g.nyp, g.nyp, g.nz, g.nz)):
            assert x.shape == (n, 4), (x.shape, n)
        check(self.lib.swpc3d_setup_pml(self.h, *map(_fp, a)))

    def setup_cerjan(self, gx_c, gx_b, gy_c, gy_b, gz_c, gz_b):
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (gx_c, gx_b, gy_c, gy_b, gz_c, gz_b)]
        check(self.lib.swpc3d_setup_cerjan(self.h, *map(_fp, a)))

    def set_sources(self, ijk, mo, mij, srcprm, stftype="kupper", bf_mode=False, tbeg=0.0):
        """ijk (nsrc,3) global indices; mo (nsrc,) already divided by M0; mij (nsrc,6) = mxx myy mzz myz mxz mxy
        (or fx fy fz in the first three columns for bf_mode); srcprm (nsrc,2)."""
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        n = ijk.shape[0]
        cols = [np.ascontiguousarray(ijk[:, q]) for q in range(3)]
        mo = np.ascontiguousarray(mo, dtype=np.float64).reshape(n)
        mij = np.ascontiguousarray(mij, dtype=np.float64).reshape(n, 6)
        m = [np.ascontiguousarray(mij[:, q]) for q in range(6)]
        prm = np.ascontiguousarray(srcprm, dtype=np.float32).reshape(n, 2)
        check(self.lib.swpc3d_set_sources(self.h, n, *map(_ip, cols), _dp(mo), *map(_dp, m), _fp(prm), stftype.encode(),
                                          int(bool(bf_mode)), float(tbeg)))

    def set_stations(self, ijk, ntdec_w, ntw, M0, UC):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        n = ijk.shape[0]
        cols = [np.ascontiguousarray(ijk[:, q]) for q in range(3)]
        check(self.lib.swpc3d_set_stations(self.h, n, *map(_ip, cols), int(ntdec_w), int(ntw), float(M0), float(UC)))
        self.nst, self.ntw = n, int(ntw)

    # ---- the reference's per-step subroutines
    def update_stress(self):
        check(self.lib.swpc3d_update_stress(self.h))

    def stressglut(self, it):
        check(self.lib.swpc3d_stressglut(self.h, it))

    def comm_stress(self):
        check(self.lib.swpc3d_comm_stress(self.h))

    def update_vel(self):
        check(self.lib.swpc3d_update_vel(self.h))

    def bodyforce(self, it):
        check(self.lib.swpc3d_bodyforce(self.h, it))

    def comm_vel(self):
        check(self.lib.swpc3d_comm_vel(self.h))

    def wav_store(self, it):
        check(self.lib.swpc3d_wav_store(self.h, it))

    def step(self, it):
        check(self.lib.swpc3d_step(self.h, it))

    def run(self, it0, it1):
        check(self.lib.swpc3d_run(self.h, it0, it1))

    def sync(self):
        check(self.lib.swpc3d_sync(self.h))

    def timer_start(self):
        check(self.lib.swpc3d_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(self.lib.swpc3d_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def vmax(self) -> np.ndarray:
        out = np.zeros(3, dtype=np.float32)
        check(self.lib.swpc3d_vmax(self.h, _fp(out)))
        return out

    def get_wav(self) -> np.ndarray:
        """(nst, 3, ntw) float32 [nm/s]"""
        out = np.zeros((max(self.nst, 1), 3, max(self.ntw, 1)), dtype=np.float32)
        if self.nst and self.ntw:
            check(self.lib.swpc3d_get_wav(self.h, _fp(out)))
        return out[: self.nst]

    # ---- multi-GPU
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().swpc3d_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, nranks: int, rank: int):
        check(self.lib.swpc3d_comm_init(self.h, uid, nranks, rank))

    def set_option(self, key: str, value: int):
        check(self.lib.swpc3d_set_option(self.h, key.encode(), int(value)))

    def info(self, key: str) -> float:
        v = C.c_double()
        check(self.lib.swpc3d_get_info(self.h, key.encode(), C.byref(v)))
        return v.value


def comm_local(ranks: list[DeviceRank], which: str):
    """Single-process emulation of global__comm_stress / global__comm_vel for ranks sharing this process."""
    arr = (C.c_void_p * len(ranks))(*[r.h for r in ranks])
    check(_lib.load().swpc3d_comm_local(arr, len(ranks), 0 if which == "stress" else 1))


def _set_wav_products(self, sw_v=True, sw_u=False, sw_stress=False, sw_strain=False):
    """which products wav__store keeps (m_wav.f90:67-70); call after set_stations"""
    check(self.lib.swpc3d_set_wav_products(self.h, int(sw_v), int(sw_u), int(sw_stress), int(sw_strain)))


def _get_wav_product(self, which: int) -> np.ndarray:
    """which: 0 velocity, 1 displacement (nst,3,ntw); 2 stress, 3 strain (nst,6,ntw)"""
    nc = 3 if which < 2 else 6
    out = np.zeros((max(self.nst, 1), nc, max(self.ntw, 1)), dtype=np.float32)
    if self.nst and self.ntw:
        check(self.lib.swpc3d_get_wav_product(self.h, which, _fp(out)))
    return out[: self.nst]


DeviceRank.set_wav_products = _set_wav_products
DeviceRank.get_wav_product = _get_wav_product
