"""`Swpc3d` -- Python face of the host-side driver (include/swpc3d_host.h): `program swpc_3d` of the reference
(src/swpc_3d/main.f90) for one rank, with the time loop on the GPU.

    run = Swpc3d("example/input.inf", base_dir=".", nm=3)      # main.f90:55-78: setup chain on the CPU (setup only)
    run.attach_device(0)                                        # main.f90:80-113: `!$acc enter data`
    run.run(1, run["nt"], verbose=True)                         # main.f90:119-139 on the B200
    run.write_sac("./out")                                      # wav__write

Multi-GPU (one process per GPU, torchrun): see `openswpc_b200.distributed`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

_INTS = {"nx", "ny", "nz", "nt", "na", "nm", "nproc_x", "nproc_y", "myid", "ibeg", "iend", "jbeg", "jend", "nxp", "nyp", "ibeg_k",
         "iend_k", "jbeg_k", "jend_k", "kbeg_k", "kend_k", "nsrc", "nst", "ntw", "ntdec_w", "ntdec_r", "bf_mode", "nzm", "nxm",
         "nym", "exedate", "tz_minutes", "field_bytes", "green_mode", "ng", "green_ncmp", "green_ntw"}
_DOUBLES = {"dx", "dy", "dz", "dt", "xbeg", "ybeg", "zbeg", "tbeg", "vmin", "vmax", "vmin_local", "vmax_local", "fmax", "fcut", "M0",
            "UC", "zeta", "d2", "c", "r", "loop_seconds", "evlo", "evla", "evdp", "clon", "clat", "phi"}
_STRS = {"title", "odir", "abc_type", "stftype", "vmodel_type", "stf_format", "wav_format"}
_F32 = {"rho", "lam", "mu", "taup", "taus", "gxc", "gxe", "gyc", "gye", "gzc", "gze", "gx_c", "gx_b", "gy_c", "gy_b", "gz_c", "gz_b",
        "ts", "c1", "c2", "d1", "srcprm", "xc", "yc", "zc", "stlo", "stla", "wav", "wav_u", "wav_stress", "wav_strain", "green_gf"}
_I32 = {"kfs", "kob", "kfs_top", "kfs_bot", "kob_top", "kob_bot", "kbeg_a", "src_ijk", "st_ijk", "green_ijk", "green_gid"}
_F64 = {"mo", "mij"} | {"init_" + f for f in ("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")}

_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    vp, i32, cp = C.c_void_p, C.c_int32, C.c_char_p
    lib.swpc3d_host_last_error.restype = cp
    lib.swpc3d_host_create.argtypes = [cp, cp, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]
    lib.swpc3d_host_create_from_text.argtypes = [cp, cp, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]
    lib.swpc3d_host_destroy.argtypes = [vp]
    lib.swpc3d_host_get_int.argtypes = [vp, cp, C.POINTER(i32)]
    lib.swpc3d_host_get_double.argtypes = [vp, cp, C.POINTER(C.c_double)]
    lib.swpc3d_host_get_string.argtypes = [vp, cp, cp, i32]
    lib.swpc3d_host_set_minmax.argtypes = [vp, C.c_float, C.c_float]
    lib.swpc3d_host_set_exedate.argtypes = [vp, i32, i32]
    lib.swpc3d_host_get_array.argtypes = [vp, cp, vp, C.c_int64, C.POINTER(C.c_int64)]
    lib.swpc3d_host_station_name.argtypes = [vp, i32, cp]
    lib.swpc3d_host_attach_device.argtypes = [vp, i32]
    lib.swpc3d_host_handle.argtypes = [vp]
    lib.swpc3d_host_handle.restype = vp
    lib.swpc3d_host_run.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_float), i32, C.POINTER(i32)]
    lib.swpc3d_host_write_sac.argtypes = [vp, cp, C.POINTER(i32)]
    lib.swpc3d_host_banner.argtypes = [vp]
    lib.swpc3d_host_write_tim.argtypes = [vp, cp]
    lib.swpc3d_host_green_query.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.swpc3d_host_green_set_source.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.swpc3d_host_write_green.argtypes = [vp, cp, C.POINTER(i32)]
    lib.swpc3d_host_snap_open.argtypes = [vp, cp]
    lib.swpc3d_host_snap_close.argtypes = [vp]
    _bound = True


class Swpc3dHostError(RuntimeError):
    pass


class Swpc3d:
    def __init__(self, inf=None, *, text: str | None = None, base_dir=".", nm: int = 3, myid: int = 0, nproc_x: int = 0,
                 nproc_y: int = 0, nt: int = 0, field_dtype=np.float64):
        self.lib = _lib.load()
        _bind(self.lib)
        fb = np.dtype(field_dtype).itemsize
        h = C.c_void_p()
        if text is not None:
            rc = self.lib.swpc3d_host_create_from_text(text.encode(), str(base_dir).encode(), nm, myid, nproc_x, nproc_y, nt, fb, C.byref(h))
        else:
            rc = self.lib.swpc3d_host_create(os.fspath(inf).encode(), str(base_dir).encode(), nm, myid, nproc_x, nproc_y, nt, fb, C.byref(h))
        self._ck(rc)
        self.h = h
        self.field_dtype = np.dtype(field_dtype)

    def _ck(self, rc):
        if rc:
            raise Swpc3dHostError(self.lib.swpc3d_host_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.swpc3d_host_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scalars / arrays
    def __getitem__(self, name: str):
        if name in _INTS:
            v = C.c_int32()
            self._ck(self.lib.swpc3d_host_get_int(self.h, name.encode(), C.byref(v)))
            return v.value
        if name in _DOUBLES:
            v = C.c_double()
            self._ck(self.lib.swpc3d_host_get_double(self.h, name.encode(), C.byref(v)))
            return v.value
        if name in _STRS:
            b = C.create_string_buffer(512)
            self._ck(self.lib.swpc3d_host_get_string(self.h, name.encode(), b, 512))
            return b.value.decode()
        raise KeyError(name)

    def array(self, name: str) -> np.ndarray:
        dt = np.float32 if name in _F32 else np.int32 if name in _I32 else np.float64 if name in _F64 else None
        if dt is None:
            raise KeyError(name)
        n = C.c_int64()
        self._ck(self.lib.swpc3d_host_get_array(self.h, name.encode(), None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=dt)
        if n.value:
            self._ck(self.lib.swpc3d_host_get_array(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        nym, nxm, nzm = self["nym"], self["nxm"], self["nzm"]
        if name in ("rho", "lam", "mu", "taup", "taus") or (name.startswith("init_") and n.value):
            return out.reshape(nym, nxm, nzm)
        if name in ("kfs", "kob", "kfs_top", "kfs_bot", "kob_top", "kob_bot", "kbeg_a"):
            return out.reshape(nym, nxm)
        if name in ("gxc", "gxe", "gyc", "gye", "gzc", "gze"):
            return out.reshape(-1, 4)
        if name in ("src_ijk", "st_ijk", "green_ijk"):
            return out.reshape(-1, 3)
        if name == "green_gf":
            return out.reshape(-1, max(self["green_ntw"], 1))
        if name == "mij":
            return out.reshape(-1, 6)
        if name == "srcprm":
            return out.reshape(-1, 2)
        if name in ("wav", "wav_u"):
            return out.reshape(-1, 3, max(self["ntw"], 1))
        if name in ("wav_stress", "wav_strain"):
            return out.reshape(-1, 6, max(self["ntw"], 1))
        return out

    def station_names(self) -> list[str]:
        out = []
        for i in range(self["nst"]):
            b = C.create_string_buffer(9)
            self._ck(self.lib.swpc3d_host_station_name(self.h, i, b))
            out.append(b.value.decode())
        return out

    def set_minmax(self, vmin: float, vmax: float):
        self._ck(self.lib.swpc3d_host_set_minmax(self.h, vmin, vmax))

    def set_exedate(self, exedate: int, tz_minutes: int):
        self._ck(self.lib.swpc3d_host_set_exedate(self.h, exedate, tz_minutes))

    # ---- device
    def attach_device(self, device: int = -1):
        self._ck(self.lib.swpc3d_host_attach_device(self.h, device))

    @property
    def handle(self) -> C.c_void_p:
        hh = self.lib.swpc3d_host_handle(self.h)
        if not hh:
            raise Swpc3dHostError("no device attached")
        return C.c_void_p(hh)

    def banner(self):
        self._ck(self.lib.swpc3d_host_banner(self.h))

    def run(self, it0: int = 1, it1: int | None = None, verbose: bool = False) -> np.ndarray:
        """Time loop main.f90:119-139; returns the (nreports, 3) max-amplitude table of report__progress."""
        it1 = self["nt"] if it1 is None else it1
        ntr = max(1, self["ntdec_r"])
        cap = (it1 - it0 + 1) // ntr + 2
        vm = np.zeros((cap, 3), dtype=np.float32)
        nrec = C.c_int32()
        self._ck(self.lib.swpc3d_host_run(self.h, it0, it1, int(verbose), vm.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(nrec)))
        return vm[: nrec.value]

    def write_sac(self, odir=None) -> int:
        n = C.c_int32()
        self._ck(self.lib.swpc3d_host_write_sac(self.h, os.fspath(odir).encode() if odir is not None else None, C.byref(n)))
        return n.value

    # ---- Green's-function mode (m_green.f90)
    def write_tim(self, odir=None):
        """pwatch__report: <odir>/<title>.tim (stopwatch_mode, on by default)."""
        self._ck(self.lib.swpc3d_host_write_tim(self.h, os.fspath(odir).encode() if odir is not None else None))

    def green_query(self):
        """(found, ijk[3], xyz[3], lonlat[2]) of the pseudo source on this rank (wav__stquery)."""
        f, ijk, xyz, ll = C.c_int32(), (C.c_int32 * 3)(), (C.c_float * 3)(), (C.c_float * 2)()
        self._ck(self.lib.swpc3d_host_green_query(self.h, C.byref(f), ijk, xyz, ll))
        return bool(f.value), list(ijk), list(xyz), list(ll)

    def green_set_source(self, ijk, xyz, lonlat):
        """The broadcast of m_green.f90:176-183: hand every rank the owner's answer; reads the grid-point list."""
        self._ck(self.lib.swpc3d_host_green_set_source(self.h, (C.c_int32 * 3)(*ijk), (C.c_float * 3)(*xyz), (C.c_float * 2)(*lonlat)))

    def write_green(self, odir=None) -> int:
        """green__export: SAC / CSF files of the Green's-function traces; returns the file count."""
        n = C.c_int32()
        self._ck(self.lib.swpc3d_host_write_green(self.h, os.fspath(odir).encode() if odir is not None else None, C.byref(n)))
        return n.value

    def snap_open(self, odir=None):
        """Create the netCDF snapshot files (m_snap.f90 newfile_*_nc); call after attach_device (and attach_nccl)."""
        self._ck(self.lib.swpc3d_host_snap_open(self.h, os.fspath(odir).encode() if odir is not None else None))

    def snap_close(self):
        """snap__closefiles: running maxima + close."""
        self._ck(self.lib.swpc3d_host_snap_close(self.h))

    def wav(self) -> np.ndarray:
        """(nst, 3, ntw) float32 [nm/s] after write_sac() / fetch_wav()."""
        return self.array("wav")

    # ---- low-level access to the kernel ABI of the attached device
    def device_call(self, fn: str, *args):
        _lib.check(getattr(self.lib, fn)(self.handle, *args))

    def timer_start(self):
        self.device_call("swpc3d_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.device_call("swpc3d_timer_stop", C.byref(ms))
        return ms.value

    def set_option(self, key: str, value: int):
        self.device_call("swpc3d_set_option", key.encode(), int(value))

    def info(self, key: str) -> float:
        v = C.c_double()
        self.device_call("swpc3d_get_info", key.encode(), C.byref(v))
        return v.value

    def download_fields(self, names=("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")) -> dict:
        shape = (self["nym"], self["nxm"], self["nzm"])
        out = {n: np.zeros(shape, dtype=self.field_dtype) for n in names}
        allf = ("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")
        args = [out[n].ctypes.data_as(C.c_void_p) if n in out else None for n in allf]
        self.device_call("swpc3d_download_fields", *args)
        return out

