"""`SwpcPsv` -- Python face of the swpc_psv host-side driver (include/swpcpsv_host.h): `program swpc_psv` of the reference
(src/swpc_psv/main.f90) for one rank, with the time loop on the GPU.

    run = SwpcPsv("input.inf", base_dir=".", nm=3)      # main.f90:55-78: setup chain on the CPU (setup only)
    run.attach_device(0)                                 # main.f90:80-93: `!$acc enter data`
    run.run(1, run["nt"], verbose=True)                  # main.f90:95-113 on the B200
    run.write_wav("./out")                               # wav__write
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

_INTS = {"nx", "nz", "nt", "na", "nm", "nproc_x", "myid", "ibeg", "iend", "nxp", "ibeg_k", "iend_k", "kend_k", "nsrc", "nst", "ntw", "ntdec_w",
         "ntdec_r", "bf_mode", "nzm", "nxm", "exedate"}
_DOUBLES = {"dx", "dz", "dt", "xbeg", "zbeg", "tbeg", "vmin", "vmax", "vmin_local", "vmax_local", "fmax", "fcut", "M0", "UC", "zeta", "d2", "c", "r",
            "loop_seconds", "evlo", "evla", "evdp"}
_F32 = {"rho", "lam", "mu", "taup", "taus", "gxc", "gxe", "gzc", "gze", "gx_c", "gx_b", "gz_c", "gz_b", "ts", "srcprm", "wav0", "wav1", "wav2", "wav3"}
_I32 = {"kfs", "kob", "kfs_top", "kfs_bot", "kob_top", "kob_bot", "kbeg_a", "src_ik", "st_ik"}
_F64 = {"mo", "m3", "init_Vx", "init_Vz", "init_Sxx", "init_Szz", "init_Sxz"}

# every symbol include/swpcpsv_host.h declares
HOST_SYMBOLS = ["swpcpsv_host_create", "swpcpsv_host_create_from_text", "swpcpsv_host_destroy", "swpcpsv_host_last_error", "swpcpsv_host_get_int",
                "swpcpsv_host_get_double", "swpcpsv_host_set_minmax", "swpcpsv_host_set_exedate", "swpcpsv_host_get_array",
                "swpcpsv_host_station_name", "swpcpsv_host_attach_device", "swpcpsv_host_handle", "swpcpsv_host_run", "swpcpsv_host_write_wav",
                "swpcpsv_host_banner", "swpcpsv_host_snap_open", "swpcpsv_host_snap_close"]

_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    vp, i32, cp = C.c_void_p, C.c_int32, C.c_char_p
    lib.swpcpsv_host_last_error.restype = cp
    lib.swpcpsv_host_create.argtypes = [cp, cp, i32, i32, i32, i32, i32, C.POINTER(vp)]
    lib.swpcpsv_host_create_from_text.argtypes = [cp, cp, i32, i32, i32, i32, i32, C.POINTER(vp)]
    lib.swpcpsv_host_destroy.argtypes = [vp]
    lib.swpcpsv_host_get_int.argtypes = [vp, cp, C.POINTER(i32)]
    lib.swpcpsv_host_get_double.argtypes = [vp, cp, C.POINTER(C.c_double)]
    lib.swpcpsv_host_set_minmax.argtypes = [vp, C.c_float, C.c_float]
    lib.swpcpsv_host_set_exedate.argtypes = [vp, i32, i32]
    lib.swpcpsv_host_get_array.argtypes = [vp, cp, vp, C.c_int64, C.POINTER(C.c_int64)]
    lib.swpcpsv_host_station_name.argtypes = [vp, i32, cp]
    lib.swpcpsv_host_attach_device.argtypes = [vp, i32]
    lib.swpcpsv_host_handle.argtypes = [vp]
    lib.swpcpsv_host_handle.restype = vp
    lib.swpcpsv_host_run.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_float), i32, C.POINTER(i32)]
    lib.swpcpsv_host_write_wav.argtypes = [vp, cp, C.POINTER(i32)]
    lib.swpcpsv_host_banner.argtypes = [vp]
    lib.swpcpsv_host_snap_open.argtypes = [vp, cp]
    lib.swpcpsv_host_snap_close.argtypes = [vp]
    _bound = True


class SwpcPsvError(RuntimeError):
    pass


class SwpcPsv:
    def __init__(self, inf=None, *, text: str | None = None, base_dir=".", nm: int = 3, myid: int = 0, nproc_x: int = 0, nt: int = 0,
                 field_dtype=np.float64):
        self.lib = _lib.load()
        _bind(self.lib)
        fb = np.dtype(field_dtype).itemsize
        h = C.c_void_p()
        if text is not None:
            rc = self.lib.swpcpsv_host_create_from_text(text.encode(), str(base_dir).encode(), nm, myid, nproc_x, nt, fb, C.byref(h))
        else:
            rc = self.lib.swpcpsv_host_create(os.fspath(inf).encode(), str(base_dir).encode(), nm, myid, nproc_x, nt, fb, C.byref(h))
        self._ck(rc)
        self.h = h
        self.field_dtype = np.dtype(field_dtype)

    def _ck(self, rc):
        if rc != 0:
            raise SwpcPsvError(self.lib.swpcpsv_host_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.swpcpsv_host_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getitem__(self, name: str):
        if name in _INTS:
            v = C.c_int32()
            self._ck(self.lib.swpcpsv_host_get_int(self.h, name.encode(), C.byref(v)))
            return int(v.value)
        if name in _DOUBLES:
            v = C.c_double()
            self._ck(self.lib.swpcpsv_host_get_double(self.h, name.encode(), C.byref(v)))
            return float(v.value)
        if name in _F32 | _I32 | _F64:
            dt = np.float32 if name in _F32 else np.int32 if name in _I32 else np.float64
            n = C.c_int64()
            self._ck(self.lib.swpcpsv_host_get_array(self.h, name.encode(), None, 0, C.byref(n)))
            out = np.zeros(int(n.value), dtype=dt)
            if out.size:
                self._ck(self.lib.swpcpsv_host_get_array(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
            return out
        raise KeyError(name)

    def station_names(self):
        buf = C.create_string_buffer(9)
        out = []
        for i in range(self["nst"]):
            self._ck(self.lib.swpcpsv_host_station_name(self.h, i, buf))
            out.append(buf.value.decode())
        return out

    def set_minmax(self, vmin: float, vmax: float):
        self._ck(self.lib.swpcpsv_host_set_minmax(self.h, C.c_float(vmin), C.c_float(vmax)))

    def set_exedate(self, exedate: int, tz_minutes: int = 0):
        self._ck(self.lib.swpcpsv_host_set_exedate(self.h, exedate, tz_minutes))

    def attach_device(self, device: int = -1):
        self._ck(self.lib.swpcpsv_host_attach_device(self.h, device))

    @property
    def handle(self):
        return C.c_void_p(self.lib.swpcpsv_host_handle(self.h))

    def banner(self):
        self._ck(self.lib.swpcpsv_host_banner(self.h))

    def run(self, it0: int = 1, it1: int | None = None, verbose: bool = False) -> np.ndarray:
        """main.f90:95-113; returns the (nrec, 2) max-amplitude pairs of report__progress."""
        it1 = self["nt"] if it1 is None else it1
        cap = max((it1 - it0 + 1) // max(self["ntdec_r"], 1) + 2, 1)
        vm = np.zeros((cap, 2), dtype=np.float32)
        n = C.c_int32()
        self._ck(self.lib.swpcpsv_host_run(self.h, it0, it1, int(verbose), vm.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return vm[:n.value]

    def download_fields(self) -> dict:
        """Vx Vz Sxx Szz Sxz over the memory box, (nxm, nzm), from the attached device."""
        nxm, nzm = self["iend"] - self["ibeg"] + 7, self["nz"] + 6
        out = {n: np.zeros((nxm, nzm), dtype=self.field_dtype) for n in ("Vx", "Vz", "Sxx", "Szz", "Sxz")}
        _lib.check_psv(self.lib.swpcpsv_download_fields(self.handle, *[out[n].ctypes.data_as(C.c_void_p) for n in ("Vx", "Vz", "Sxx", "Szz", "Sxz")]))
        return out

    def snap_open(self, odir=None):
        """Create <title>.psv.xz.<ps|v|u>.<nc|snp> (m_snap.f90 newfile_xz[_nc]); call after attach_device, before run()."""
        self._ck(self.lib.swpcpsv_host_snap_open(self.h, None if odir is None else os.fspath(odir).encode()))

    def snap_close(self):
        self._ck(self.lib.swpcpsv_host_snap_close(self.h))

    def write_wav(self, odir=None) -> int:
        n = C.c_int32()
        self._ck(self.lib.swpcpsv_host_write_wav(self.h, None if odir is None else os.fspath(odir).encode(), C.byref(n)))
        return int(n.value)

    def wav(self, which: int = 0) -> np.ndarray:
        """traces fetched by the last write_wav: (nst, ncmp, ntw)."""
        a = self[f"wav{which}"]
        ncmp = 2 if which < 2 else 3
        return a.reshape(self["nst"], ncmp, self["ntw"]) if a.size else a
