"""ctypes binding of the CPU oracle (oracle/, TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The oracle restates the reference's swpc_3d path (see oracle/ora.h); it is the checker,
never the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _load(name: str) -> C.CDLL:
    path = ORACLE_DIR / "_build" / name
    if not path.exists():
        build_oracle()
    lib = C.CDLL(str(path))
    vp, ci, cd, cc = C.c_void_p, C.c_int, C.c_double, C.c_char_p
    lib.ora_create.restype = vp
    lib.ora_create.argtypes = [cc, cc, ci, ci, ci, ci]
    lib.ora_create_from_text.restype = vp
    lib.ora_create_from_text.argtypes = [cc, cc, ci, ci, ci, ci]
    lib.ora_destroy.argtypes = [vp]
    lib.ora_last_error.restype = cc
    for f in ("ora_update_stress", "ora_comm_stress", "ora_update_vel", "ora_comm_vel"):
        getattr(lib, f).argtypes = [vp]
    for f in ("ora_stressglut", "ora_bodyforce", "ora_wav_store", "ora_step"):
        getattr(lib, f).argtypes = [vp, ci]
    lib.ora_vmax.argtypes = [vp, C.POINTER(C.c_float)]
    lib.ora_run.restype = ci
    lib.ora_run.argtypes = [vp, ci, ci, C.POINTER(C.c_float), ci]
    lib.ora_nranks.restype = ci
    lib.ora_nranks.argtypes = [vp]
    lib.ora_rank_int.restype = ci
    lib.ora_rank_int.argtypes = [vp, ci, ci]
    lib.ora_cfg_value.restype = cd
    lib.ora_cfg_value.argtypes = [vp, ci]
    lib.ora_cfg_int.restype = ci
    lib.ora_cfg_int.argtypes = [vp, ci]
    lib.ora_cfg_str.restype = cc
    lib.ora_cfg_str.argtypes = [vp, ci]
    lib.ora_set_exedate.argtypes = [vp, ci, ci]
    lib.ora_get_field.restype = ci
    lib.ora_get_field.argtypes = [vp, ci, cc, C.POINTER(cd)]
    lib.ora_set_field.restype = ci
    lib.ora_set_field.argtypes = [vp, ci, cc, C.POINTER(cd)]
    lib.ora_get_map.restype = ci
    lib.ora_get_map.argtypes = [vp, ci, cc, C.POINTER(ci)]
    lib.ora_gather_field.restype = ci
    lib.ora_gather_field.argtypes = [vp, cc, C.POINTER(cd)]
    lib.ora_get_sources.restype = ci
    lib.ora_get_sources.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(cd)]
    lib.ora_get_stations.restype = ci
    lib.ora_get_stations.argtypes = [vp, ci, C.POINTER(ci), C.c_char_p]
    lib.ora_get_wav.restype = ci
    lib.ora_get_wav.argtypes = [vp, ci, C.POINTER(C.c_float)]
    lib.ora_get_profile.restype = ci
    lib.ora_get_profile.argtypes = [vp, ci, cc, C.POINTER(C.c_float)]
    lib.ora_write_sac.restype = ci
    lib.ora_write_sac.argtypes = [vp, cc]
    # Green's-function mode (ora_green.c)
    lib.ora_green_int.restype = ci
    lib.ora_green_int.argtypes = [vp, ci, ci]
    lib.ora_green_points.restype = ci
    lib.ora_green_points.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
    lib.ora_get_green.restype = ci
    lib.ora_get_green.argtypes = [vp, ci, C.POINTER(C.c_float)]
    lib.ora_write_green_sac.restype = ci
    lib.ora_write_green_sac.argtypes = [vp, cc]
    # small helpers
    lib.ora_x2i.restype = ci
    lib.ora_x2i.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.ora_i2x.restype = C.c_float
    lib.ora_i2x.argtypes = [ci, C.c_float, C.c_float]
    lib.ora_decomp1d.argtypes = [ci, ci, ci, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
    lib.ora_momentrate.restype = C.c_float
    lib.ora_momentrate.argtypes = [C.c_float, cc, C.POINTER(C.c_float)]
    lib.ora_damping_profile.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, ci, C.c_float, C.c_float,
                                        C.POINTER(C.c_float)]
    lib.ora_geomap_c2g.argtypes = [C.c_float] * 5 + [C.POINTER(C.c_float)] * 2
    lib.ora_geomap_g2c.argtypes = [C.c_float] * 5 + [C.POINTER(C.c_float)] * 2
    lib.ora_sdr2moment.argtypes = [C.c_float] * 3 + [C.POINTER(C.c_float)] * 6
    lib.ora_seawater_vel.restype = C.c_float
    lib.ora_seawater_vel.argtypes = [C.c_float, ci]
    return lib


_LIBS: dict[str, C.CDLL] = {}


def lib(mp: str = "dp") -> C.CDLL:
    name = "liboracle.so" if mp == "dp" else "liboracle_sp.so"
    if name not in _LIBS:
        _LIBS[name] = _load(name)
    return _LIBS[name]


RANK_INTS = ["ibeg", "iend", "jbeg", "jend", "nxp", "nyp", "ibeg_k", "iend_k", "jbeg_k", "jend_k", "kbeg_k", "kend_k",
             "nsrc", "nst", "idx", "idy", "nzm", "nxm", "nym", "ibeg_m", "jbeg_m", "kbeg_m"]
CFG_VALUES = ["vmin", "vmax", "fmax", "fcut", "M0", "UC", "zeta", "d2", "dt", "xbeg", "ybeg", "zbeg", "dx", "dy", "dz", "c",
              "r"]
CFG_INTS = ["nx", "ny", "nz", "nt", "na", "nm", "nproc_x", "nproc_y", "ntw", "ntdec_w", "ntdec_r", "bf_mode"]


class Oracle:
    """One emulated multi-rank swpc_3d run on the CPU."""

    def __init__(self, inf: str | os.PathLike | None = None, *, text: str | None = None, base_dir: str | os.PathLike = ".",
                 nm: int = 3, nproc_x: int = 0, nproc_y: int = 0, nt: int = 0, mp: str = "dp"):
        self.lib = lib(mp)
        if text is not None:
            h = self.lib.ora_create_from_text(text.encode(), str(base_dir).encode(), nm, nproc_x, nproc_y, nt)
        else:
            h = self.lib.ora_create(str(inf).encode(), str(base_dir).encode(), nm, nproc_x, nproc_y, nt)
        if not h:
            raise RuntimeError("oracle: " + self.lib.ora_last_error().decode())
        self.h = C.c_void_p(h)
        self.nranks = self.lib.ora_nranks(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ora_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- config
    def cfg(self, name: str):
        if name in CFG_VALUES:
            return self.lib.ora_cfg_value(self.h, CFG_VALUES.index(name))
        if name in CFG_INTS:
            return self.lib.ora_cfg_int(self.h, CFG_INTS.index(name))
        strs = ["title", "odir", "abc_type", "stftype", "vmodel_type"]
        if name in strs:
            return self.lib.ora_cfg_str(self.h, strs.index(name)).decode()
        raise KeyError(name)

    def ts(self):
        return np.array([self.lib.ora_cfg_value(self.h, 17 + m) for m in range(self.cfg("nm"))], dtype=np.float32)

    def coef(self, which: str):
        base = {"c1": 25, "c2": 33, "d1": 41}[which]
        return np.array([self.lib.ora_cfg_value(self.h, base + m) for m in range(self.cfg("nm"))], dtype=np.float32)

    def rank(self, q: int) -> dict:
        return {n: self.lib.ora_rank_int(self.h, q, i) for i, n in enumerate(RANK_INTS)}

    # ---- stepping
    def step(self, it: int):
        self.lib.ora_step(self.h, it)

    def run(self, it0: int, it1: int):
        n = max(1, (it1 - it0 + 1) // max(1, self.cfg("ntdec_r")) + 2)
        buf = (C.c_float * (3 * n))()
        k = self.lib.ora_run(self.h, it0, it1, buf, n)
        return np.frombuffer(buf, dtype=np.float32)[: 3 * k].reshape(k, 3).copy()

    def vmax(self):
        out = (C.c_float * 3)()
        self.lib.ora_vmax(self.h, out)
        return np.array(out[:], dtype=np.float32)

    # ---- data
    def field(self, q: int, name: str) -> np.ndarray:
        """(nym, nxm, nzm) array = Fortran (k,i,j) over the rank's memory box (margins included)."""
        r = self.rank(q)
        out = np.empty((r["nym"], r["nxm"], r["nzm"]), dtype=np.float64)
        rc = self.lib.ora_get_field(self.h, q, name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise KeyError(name)
        return out

    def set_field(self, q: int, name: str, arr: np.ndarray):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        rc = self.lib.ora_set_field(self.h, q, name.encode(), a.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise KeyError(name)

    def imap(self, q: int, name: str) -> np.ndarray:
        r = self.rank(q)
        out = np.empty((r["nym"], r["nxm"]), dtype=np.int32)
        rc = self.lib.ora_get_map(self.h, q, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int)))
        if rc:
            raise KeyError(name)
        return out

    def gather(self, name: str) -> np.ndarray:
        """(ny, nx, nz) array of the owned cells of all ranks."""
        out = np.zeros((self.cfg("ny"), self.cfg("nx"), self.cfg("nz")), dtype=np.float64)
        rc = self.lib.ora_gather_field(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise KeyError(name)
        return out

    def sources(self, q: int):
        n = self.rank(q)["nsrc"]
        ijk = np.zeros((max(n, 1), 3), dtype=np.int32)
        mo = np.zeros(max(n, 1), dtype=np.float64)
        self.lib.ora_get_sources(self.h, q, ijk.ctypes.data_as(C.POINTER(C.c_int)), mo.ctypes.data_as(C.POINTER(C.c_double)))
        return ijk[:n], mo[:n]

    def stations(self, q: int):
        n = self.rank(q)["nst"]
        ijk = np.zeros((max(n, 1), 3), dtype=np.int32)
        names = C.create_string_buffer(9 * max(n, 1))
        self.lib.ora_get_stations(self.h, q, ijk.ctypes.data_as(C.POINTER(C.c_int)), names)
        nm = [names.raw[9 * i: 9 * i + 9].split(b"\0")[0].decode() for i in range(n)]
        return ijk[:n], nm

    def wav(self, q: int) -> np.ndarray:
        """(nst, 3, ntw) float32 velocity traces [nm/s] of rank q."""
        n = self.rank(q)["nst"]
        ntw = self.cfg("ntw")
        out = np.zeros((max(n, 1), 3, max(ntw, 1)), dtype=np.float32)
        self.lib.ora_get_wav(self.h, q, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out[:n]

    def profile(self, q: int, name: str) -> np.ndarray:
        buf = (C.c_float * (4 * 70000))()
        n = self.lib.ora_get_profile(self.h, q, name.encode(), buf)
        if n < 0:
            raise KeyError(name)
        a = np.frombuffer(buf, dtype=np.float32)[:n].copy()
        return a.reshape(-1, 4) if name in ("gxc", "gxe", "gyc", "gye", "gzc", "gze") else a

    # ---- Green's-function mode
    GREEN_INTS = ["ng", "ncmp", "isrc", "jsrc", "ksrc", "is_src", "ntw"]

    def green(self, q: int) -> dict:
        """ints of ora_green_int + the owned grid points (ijk (ng,3), gid) + traces gf (ncmp*ng, ntw) of rank q."""
        d = {n: self.lib.ora_green_int(self.h, q, i) for i, n in enumerate(self.GREEN_INTS)}
        ng = max(d["ng"], 0)
        ijk = np.zeros((max(ng, 1), 3), dtype=np.int32)
        gid = np.zeros(max(ng, 1), dtype=np.int32)
        self.lib.ora_green_points(self.h, q, ijk.ctypes.data_as(C.POINTER(C.c_int)), gid.ctypes.data_as(C.POINTER(C.c_int)))
        gf = np.zeros((max(ng, 1) * d["ncmp"], max(d["ntw"], 1)), dtype=np.float32)
        if ng:
            self.lib.ora_get_green(self.h, q, gf.ctypes.data_as(C.POINTER(C.c_float)))
        d.update(ijk=ijk[:ng], gid=gid[:ng], gf=gf[: ng * d["ncmp"]])
        return d

    def write_green_sac(self, odir: str | os.PathLike) -> int:
        return self.lib.ora_write_green_sac(self.h, str(odir).encode())

    def write_sac(self, odir: str | os.PathLike) -> int:
        return self.lib.ora_write_sac(self.h, str(odir).encode())


# ---- snapshots (oracle/ora_snap.c)
SNAP_SECTIONS = ("xy", "xz", "yz", "fs", "ob")
SNAP_TYPES = ("ps", "v", "u")


def _snap_bind(o):
    L = o.lib
    vp, ci, fp = C.c_void_p, C.c_int, C.POINTER(C.c_float)
    L.ora_snap_info.argtypes = [vp, C.POINTER(ci)]
    L.ora_snap_coords.argtypes = [vp, fp, fp, fp]
    L.ora_snap_nrec.argtypes = [vp, ci]
    L.ora_snap_nrec.restype = ci
    L.ora_snap_rec.argtypes = [vp, ci, ci, fp, C.POINTER(ci)]
    L.ora_snap_max.argtypes = [vp, ci, fp]
    L.ora_snap_medium.argtypes = [vp, ci, ci, fp]


def snap_info(o) -> dict:
    _snap_bind(o)
    v = (C.c_int * 10)()
    o.lib.ora_snap_info(o.h, v)
    return dict(zip(["idec", "jdec", "kdec", "ntdec_s", "nxs", "nys", "nzs", "k0_xy", "i0_yz", "j0_xz"], list(v)))


def snap_dims(o, q):
    i = snap_info(o)
    sec, typ = divmod(q, 3)
    n1 = i["nys"] if sec == 2 else i["nxs"]
    n2 = i["nzs"] if sec in (1, 2) else i["nys"]
    return n1, n2, (4 if typ == 0 else 3)


def snap_records(o, q):
    """(nrec, nvar, n2, n1) array and the list of it0"""
    _snap_bind(o)
    n1, n2, nv = snap_dims(o, q)
    nrec = o.lib.ora_snap_nrec(o.h, q)
    out = np.zeros((nrec, nv, n2, n1), dtype=np.float32)
    its = []
    for r in range(nrec):
        it0 = C.c_int()
        o.lib.ora_snap_rec(o.h, q, r, out[r].ctypes.data_as(C.POINTER(C.c_float)), C.byref(it0))
        its.append(it0.value)
    return out, its


def snap_max(o, q):
    _snap_bind(o)
    n1, n2, _ = snap_dims(o, q)
    out = np.zeros((3, n2, n1), dtype=np.float32)
    rc = o.lib.ora_snap_max(o.h, q, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out if rc == 0 else None


def snap_medium(o, q, which):
    _snap_bind(o)
    n1, n2, _ = snap_dims(o, q)
    out = np.zeros((n2, n1), dtype=np.float32)
    o.lib.ora_snap_medium(o.h, q, which, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def snap_coords(o):
    _snap_bind(o)
    i = snap_info(o)
    x, y, z = (np.zeros(i[k], dtype=np.float32) for k in ("nxs", "nys", "nzs"))
    o.lib.ora_snap_coords(o.h, *(a.ctypes.data_as(C.POINTER(C.c_float)) for a in (x, y, z)))
    return x, y, z
