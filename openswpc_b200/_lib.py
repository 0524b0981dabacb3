"""ctypes binding of libswpc3d_b200.so (include/swpc3d_b200.h).

The product path has no CPU fallback: importing works anywhere (so that the symbol table can be checked
without a GPU), but every compute entry point needs the CUDA library and a device and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libswpc3d_b200.so"
CSRC = PKG / "csrc"
ROOT = PKG.parent

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-ffp-contract=off", "-Xcompiler", "-fopenmp", "-shared", "-diag-suppress", "177"]


def build(fmad: bool = False, force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [CSRC / "abi.cu", CSRC / "psv_abi.cu", CSRC / "host" / "driver.cpp", CSRC / "host" / "psv_driver.cpp"]
    srcs = [s for s in srcs if s.exists()]
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((CSRC / "host").glob("*")) + [ROOT / "include" / "swpc3d_b200.h", ROOT / "include" / "swpcpsv_b200.h"]
    if LIB_PATH.exists() and not force:
        t = LIB_PATH.stat().st_mtime
        if all(d.stat().st_mtime <= t for d in deps if d.exists()):
            return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, f"-fmad={'true' if fmad else 'false'}", "-I", str(ROOT / "include"), "-o", str(LIB_PATH),
           *map(str, srcs), "-ldl", "-lgomp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    subprocess.run(cmd, check=True, env=env)
    return LIB_PATH


class Grid(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nx", "ny", "nz", "nproc_x", "nproc_y", "myid", "ibeg", "iend", "jbeg", "jend", "ipad", "jpad", "kpad",
        "ibeg_k", "iend_k", "jbeg_k", "jend_k", "kbeg_k", "kend_k", "na", "nm", "abc_type", "field_bytes", "device",
        "reserved")] + [("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("dt", C.c_float), ("reserved_f", C.c_float)]


class PsvGrid(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nx", "nz", "nproc_x", "myid", "ibeg", "iend", "ipad", "kpad", "ibeg_k", "iend_k", "kend_k", "na", "nm", "abc_type",
        "field_bytes", "device")] + [("dx", C.c_double), ("dz", C.c_double), ("dt", C.c_float), ("reserved_f", C.c_float)]


# every symbol include/swpcpsv_b200.h declares
PSV_SYMBOLS = [
    "swpcpsv_last_error", "swpcpsv_version", "swpcpsv_create", "swpcpsv_destroy", "swpcpsv_upload_medium", "swpcpsv_upload_fields",
    "swpcpsv_download_fields", "swpcpsv_download_memvars", "swpcpsv_zero_state", "swpcpsv_setup_pml", "swpcpsv_setup_cerjan",
    "swpcpsv_set_sources", "swpcpsv_set_stations", "swpcpsv_get_wav", "swpcpsv_update_stress", "swpcpsv_stressglut",
    "swpcpsv_comm_stress", "swpcpsv_update_vel", "swpcpsv_comm_vel", "swpcpsv_wav_store", "swpcpsv_step", "swpcpsv_run",
    "swpcpsv_sync", "swpcpsv_vmax", "swpcpsv_vmax_global", "swpcpsv_nccl_unique_id", "swpcpsv_comm_init", "swpcpsv_comm_local",
    "swpcpsv_timer_start", "swpcpsv_timer_stop", "swpcpsv_set_option", "swpcpsv_get_info",
    "swpcpsv_snap_setup", "swpcpsv_snap_step", "swpcpsv_snap_fetch", "swpcpsv_reduce_sum",
]

# every symbol include/swpc3d_b200.h declares
SYMBOLS = [
    "swpc3d_last_error", "swpc3d_version", "swpc3d_create", "swpc3d_destroy", "swpc3d_upload_medium", "swpc3d_upload_fields",
    "swpc3d_download_fields", "swpc3d_zero_state", "swpc3d_setup_pml", "swpc3d_setup_cerjan", "swpc3d_set_sources",
    "swpc3d_set_stations", "swpc3d_update_stress", "swpc3d_stressglut", "swpc3d_comm_stress", "swpc3d_update_vel",
    "swpc3d_bodyforce", "swpc3d_comm_vel", "swpc3d_wav_store", "swpc3d_step", "swpc3d_run", "swpc3d_sync", "swpc3d_vmax",
    "swpc3d_get_wav", "swpc3d_vmax_global", "swpc3d_set_wav_products", "swpc3d_get_wav_product", "swpc3d_snap_setup", "swpc3d_snap_step", "swpc3d_snap_fetch",
    "swpc3d_snap_fetch_max", "swpc3d_snap_fetch_begin", "swpc3d_snap_fetch_end", "swpc3d_reduce_sum", "swpc3d_nccl_unique_id", "swpc3d_comm_init", "swpc3d_comm_local", "swpc3d_set_option", "swpc3d_get_info", "swpc3d_timer_start", "swpc3d_timer_stop",
    "swpc3d_set_green", "swpc3d_green_store", "swpc3d_green_source", "swpc3d_get_green", "swpc3d_advance",
]

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("SWPC3D_LIB") and os.environ.get("SWPC3D_DEV") == "1":
        # kernel development only (both variables needed, and it says so on stderr): an experimental build of the same ABI
        LIB_PATH = Path(os.environ["SWPC3D_LIB"])
        print(f"openswpc_b200: SWPC3D_DEV=1, loading the experimental build {LIB_PATH}", file=sys.stderr, flush=True)
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the swpc3d_b200 path has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, fp, dp, ip, cp = C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_char_p
    lib.swpc3d_last_error.restype = cp
    lib.swpc3d_version.restype = cp
    lib.swpc3d_create.argtypes = [C.POINTER(Grid), fp, C.POINTER(vp)]
    lib.swpc3d_destroy.argtypes = [vp]
    lib.swpc3d_upload_medium.argtypes = [vp] + [fp] * 5 + [ip] * 7
    lib.swpc3d_upload_fields.argtypes = [vp] + [vp] * 9
    lib.swpc3d_download_fields.argtypes = [vp] + [vp] * 9
    lib.swpc3d_zero_state.argtypes = [vp]
    lib.swpc3d_setup_pml.argtypes = [vp] + [fp] * 6
    lib.swpc3d_setup_cerjan.argtypes = [vp] + [fp] * 6
    lib.swpc3d_set_sources.argtypes = [vp, i32, ip, ip, ip] + [dp] * 7 + [fp, cp, i32, C.c_float]
    lib.swpc3d_set_stations.argtypes = [vp, i32, ip, ip, ip, i32, i32, C.c_float, C.c_float]
    for f in ("swpc3d_update_stress", "swpc3d_comm_stress", "swpc3d_update_vel", "swpc3d_comm_vel", "swpc3d_sync"):
        getattr(lib, f).argtypes = [vp]
    for f in ("swpc3d_stressglut", "swpc3d_bodyforce", "swpc3d_wav_store", "swpc3d_step", "swpc3d_advance"):
        getattr(lib, f).argtypes = [vp, i32]
    lib.swpc3d_run.argtypes = [vp, i32, i32]
    lib.swpc3d_vmax.argtypes = [vp, fp]
    lib.swpc3d_get_wav.argtypes = [vp, fp]
    lib.swpc3d_vmax_global.argtypes = [vp, fp]
    lib.swpc3d_set_wav_products.argtypes = [vp, i32, i32, i32, i32]
    lib.swpc3d_get_wav_product.argtypes = [vp, i32, fp]
    lib.swpc3d_nccl_unique_id.argtypes = [C.c_char_p]
    lib.swpc3d_comm_init.argtypes = [vp, C.c_char_p, i32, i32]
    lib.swpc3d_comm_local.argtypes = [C.POINTER(vp), i32, i32]
    lib.swpc3d_snap_fetch_begin.argtypes = [vp, i32, i32, i32]
    lib.swpc3d_snap_fetch_end.argtypes = [vp, i32, i32, C.POINTER(fp)]
    lib.swpc3d_set_green.argtypes = [vp, i32, ip, ip, ip, i32, i32, i32, i32, i32] + [C.c_float] * 4 + [cp, i32, i32, C.c_float]
    lib.swpc3d_green_store.argtypes = [vp, i32]
    lib.swpc3d_green_source.argtypes = [vp, i32]
    lib.swpc3d_get_green.argtypes = [vp, fp]
    lib.swpc3d_timer_start.argtypes = [vp]
    lib.swpc3d_timer_stop.argtypes = [vp, fp]
    lib.swpc3d_set_option.argtypes = [vp, cp, i32]
    lib.swpc3d_get_info.argtypes = [vp, cp, dp]
    for s in SYMBOLS:
        if s not in ("swpc3d_last_error", "swpc3d_version"):
            getattr(lib, s).restype = C.c_int
    # ---- swpc_psv (include/swpcpsv_b200.h)
    lib.swpcpsv_last_error.restype = cp
    lib.swpcpsv_version.restype = cp
    lib.swpcpsv_create.argtypes = [C.POINTER(PsvGrid), fp, C.POINTER(vp)]
    lib.swpcpsv_destroy.argtypes = [vp]
    lib.swpcpsv_upload_medium.argtypes = [vp] + [fp] * 5 + [ip] * 7
    lib.swpcpsv_upload_fields.argtypes = [vp] + [vp] * 5
    lib.swpcpsv_download_fields.argtypes = [vp] + [vp] * 5
    lib.swpcpsv_download_memvars.argtypes = [vp, fp, fp, fp]
    lib.swpcpsv_setup_pml.argtypes = [vp] + [fp] * 4
    lib.swpcpsv_setup_cerjan.argtypes = [vp] + [fp] * 4
    lib.swpcpsv_set_sources.argtypes = [vp, i32, ip, ip] + [dp] * 4 + [fp, cp, i32, C.c_float]
    lib.swpcpsv_set_stations.argtypes = [vp, i32, ip, ip, i32, i32, C.c_float, C.c_float, i32, i32, i32, i32]
    lib.swpcpsv_get_wav.argtypes = [vp, i32, fp]
    for f in ("swpcpsv_zero_state", "swpcpsv_update_stress", "swpcpsv_comm_stress", "swpcpsv_comm_vel", "swpcpsv_sync", "swpcpsv_timer_start"):
        getattr(lib, f).argtypes = [vp]
    for f in ("swpcpsv_stressglut", "swpcpsv_update_vel", "swpcpsv_wav_store", "swpcpsv_step"):
        getattr(lib, f).argtypes = [vp, i32]
    lib.swpcpsv_run.argtypes = [vp, i32, i32]
    lib.swpcpsv_vmax.argtypes = [vp, fp]
    lib.swpcpsv_vmax_global.argtypes = [vp, fp]
    lib.swpcpsv_nccl_unique_id.argtypes = [C.c_char_p]
    lib.swpcpsv_comm_init.argtypes = [vp, C.c_char_p, i32, i32]
    lib.swpcpsv_comm_local.argtypes = [C.POINTER(vp), i32, i32]
    lib.swpcpsv_timer_stop.argtypes = [vp, fp]
    lib.swpcpsv_snap_setup.argtypes = [vp, vp]
    lib.swpcpsv_snap_step.argtypes = [vp, i32]
    lib.swpcpsv_snap_fetch.argtypes = [vp, i32, i32, fp]
    lib.swpcpsv_reduce_sum.argtypes = [vp, fp, C.c_int64, i32]
    lib.swpcpsv_set_option.argtypes = [vp, cp, i32]
    lib.swpcpsv_get_info.argtypes = [vp, cp, dp]
    for s in PSV_SYMBOLS:
        if s not in ("swpcpsv_last_error", "swpcpsv_version"):
            getattr(lib, s).restype = C.c_int
    _lib = lib
    return lib


class Swpc3dError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise Swpc3dError(load().swpc3d_last_error().decode())


def check_psv(rc: int) -> None:
    if rc != 0:
        raise Swpc3dError(load().swpcpsv_last_error().decode())
