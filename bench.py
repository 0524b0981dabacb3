#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native swpc_3d time-stepping path.

Metric (BASELINE.json): 3-D viscoelastic cell-updates/s (+ % of the HBM roofline).
  N = 1 : BASELINE configs[3] -- swpc_3d, NM=3 GZB, ADE-CFS PML (na=20), synthetic layered model (the 8-layer table
          of the reference's example/lhm.dat), 1024 x 1024 x 512 on one B200, float64 fields (reference default MP=DP).
          After the headline, on the same GPU and under the same clock sampler ("secondary", CUDA-event timed):
            psv       BASELINE configs[1]: swpc_psv 16384 x 8192, NM=3, PML
            elastic   BASELINE configs[2]: swpc_3d benchmark_mode half-space 512^3, NM=0
            f32       configs[3] with float32 fields (the reference's MP=SP build)
          and "weak_base": the per-GPU workload of the N>1 runs (lhm_rmed, 512 x 1024 x 1024, 1x1) -- the base the
          weak-scaling efficiency has to be taken against.
  N > 1 : BASELINE configs[4] shape -- weak scaling, 512 x 1024 x 1024 cells per GPU, x-y decomposition 2x1 / 4x1 / 4x2,
          NCCL send/recv halo exchange overlapped with the core sweeps; synthetic heterogeneous crust = the layered table
          with Gaussian random media per layer (vmodel lhm_rmed).  Before the timed region a small decomposed case runs
          over NCCL on the N GPUs and is compared bit for bit with the same decomposition emulated on rank 0's GPU
          ("parity"; the emulated exchange is what the GPU test-suite ties to the oracle); a mismatch exits non-zero.
A "step" is one iteration of main.f90:119-139 (stress sweep, stress glut, halo, velocity sweep, halo).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...

`--impl reference` times the reference's CPU implementation of the same path (the oracle port: the Fortran cannot be
built in this image) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LHM = """# depth  rho  vp  vs  Qp  Qs   (example/lhm.dat of the reference)
      0          2.3       5.5      3.14      600     300
      3          2.4       6.0      3.55      600     300
     18          2.8       6.7      3.83      600     300
     33          3.2       7.8      4.46      600     300
    100          3.3       8.0      4.57      600     300
    225          3.4       8.4      4.80      600     300
    325          3.5       8.6      4.91      600     300
    425          3.7       9.3      5.31      600     300
"""

# the reference's example/input.inf:55-83 snapshot block (6 products on, every 5 steps, netCDF), used by the e2e_snap leg
SNAP_BLOCK = """
 snp_format = 'netcdf'
 xy_ps%sw = .false.
 xz_ps%sw = .true.
 yz_ps%sw = .false.
 fs_ps%sw = .false.
 ob_ps%sw = .true.
 xy_v%sw = .false.
 xz_v%sw = .true.
 yz_v%sw = .false.
 fs_v%sw = .false.
 ob_v%sw = .true.
 xy_u%sw = .false.
 xz_u%sw = .true.
 yz_u%sw = .false.
 fs_u%sw = .false.
 ob_u%sw = .true.
 z0_xy = 7.0
 x0_yz = 0.0
 y0_xz = 0.0
 ntdec_s = 5
 idec = 2
 jdec = 2
 kdec = 2
"""


def write_workload(d: Path, nx: int, ny: int, nz: int, nt: int, npx: int, npy: int, dx=0.5, dt=0.025, na=20, hetero=False,
                   benchmark=False, extra="") -> Path:
    d.mkdir(parents=True, exist_ok=True)
    (d / "lhm.dat").write_text(LHM)
    if hetero:
        # configs[4]: heterogeneous crust = the layered table with a random-media volume per layer (vmodel lhm_rmed,
        # m_vmodel_lhm_rmed.f90: Vp, Vs * (1 + xi), rho * (1 + 0.8 xi)); xi ~ N(0, 0.03^2), Gaussian-smoothed, periodic,
        # two independent 96^3 volumes alternating between layers, seeds 20251017 / 20251018
        from openswpc_b200.rmed import smoothed_gaussian, write_rmed3d

        for q in range(2):
            write_rmed3d(d / f"rmed{q}.nc", smoothed_gaussian((96, 96, 96), 3.0, 0.03, 20251017 + q), dx)
        rows = [ln for ln in LHM.splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
        (d / "lhm_rmed.dat").write_text("# depth rho vp vs Qp Qs rmed\n" + "\n".join(f"{ln} rmed{q % 2}.nc" for q, ln in enumerate(rows)) + "\n")
    vm = " vmodel_type = 'lhm_rmed'\n fn_lhm_rmed = 'lhm_rmed.dat'\n dir_rmed = '.'\n rhomin = 1.0" if hetero else " vmodel_type = 'lhm'\n fn_lhm = 'lhm.dat'"
    (d / "source.dat").write_text("# x y z tbeg trise mo mxx myy mzz myz mxz mxy\n 0.0 0.0 10.0 0.1 4.0 1.e15 0.8165 0.8165 0.8165 0.0 0.0 0.0\n")
    st = []
    for a in range(8):
        for b in range(8):
            x = (a - 3.5) * nx * dx / 10.0
            y = (b - 3.5) * ny * dx / 10.0
            st.append(f"{x:.3f} {y:.3f} 0.0 s{a}{b} obb")
    (d / "stloc.xy").write_text("\n".join(st) + "\n")
    inf = f"""
 title = 'bench'
 odir = './out'
 ntdec_r = 10
 benchmark_mode = {'.true.' if benchmark else '.false.'}
 nproc_x = {npx}
 nproc_y = {npy}
 nx = {nx}
 ny = {ny}
 nz = {nz}
 nt = {nt}
 dx = {dx}
 dy = {dx}
 dz = {dx}
 dt = {dt}
 vcut = 1.5
 xbeg = {-nx * dx / 2}
 ybeg = {-ny * dx / 2}
 zbeg = -10.0
 tbeg = 0.0
 fq_min = 0.02
 fq_max = 2.00
 fq_ref = 1.0
 sw_wav_v = .true.
 ntdec_w = 5
 st_format = 'xy'
 fn_stloc = 'stloc.xy'
 wav_format = 'sac'
 stf_format = 'xym0ij'
 stftype = 'kupper'
 fn_stf = 'source.dat'
 abc_type = 'pml'
 na = {na}
{vm}
 munk_profile = .true.
{extra}
"""
    p = d / "input.inf"
    p.write_text(inf)
    return p


def write_psv_workload(d: Path, nx: int, nz: int, nt: int, dx=0.25, dt=0.0125, na=20) -> Path:
    """BASELINE configs[1] (SURVEY 8d input 2): swpc_psv, layered model, one moment-tensor source, 16 stations on a line."""
    d.mkdir(parents=True, exist_ok=True)
    (d / "lhm.dat").write_text(LHM)
    (d / "source.dat").write_text("# x y z tbeg trise mo mxx myy mzz myz mxz mxy\n 0.0 0.0 10.0 0.1 2.0 1.e15 0.7 0.0 -0.3 0.0 0.5 0.0\n")
    (d / "stloc.xy").write_text("\n".join(f"{(a - 7.5) * nx * dx / 20.0:.3f} 0.0 0.0 p{a:02d} obb" for a in range(16)) + "\n")
    p = d / "input.inf"
    p.write_text(f"""
 title = 'benchpsv'
 odir = './out'
 ntdec_r = 10
 nproc_x = 1
 nx = {nx}
 nz = {nz}
 nt = {nt}
 dx = {dx}
 dz = {dx}
 dt = {dt}
 na = {na}
 xbeg = {-nx * dx / 2}
 zbeg = -10.0
 tbeg = 0.0
 vcut = 1.5
 abc_type = 'pml'
 vmodel_type = 'lhm'
 fn_lhm = 'lhm.dat'
 fq_min = 0.02
 fq_max = 2.0
 fq_ref = 1.0
 fn_stf = 'source.dat'
 stftype = 'kupper'
 stf_format = 'xym0ij'
 fn_stloc = 'stloc.xy'
 st_format = 'xy'
 ntdec_w = 5
 sw_wav_v = .true.
""")
    return p


def cell_counts(run, region=None) -> tuple[int, int]:
    """(interior cells, absorber cells) of this rank (m_global.f90:334-376), optionally restricted to the owned columns
    li0..li1 x lj0..lj1 (local, inclusive) -- the core region of the boundary-first overlap."""
    nz = run["nz"]
    li0, li1, lj0, lj1 = region if region else (0, run["nxp"] - 1, 0, run["nyp"] - 1)
    ki0, ki1 = max(run["ibeg_k"] - run["ibeg"], li0), min(run["iend_k"] - run["ibeg"], li1)
    kj0, kj1 = max(run["jbeg_k"] - run["jbeg"], lj0), min(run["jend_k"] - run["jbeg"], lj1)
    nzk = max(0, run["kend_k"] - run["kbeg_k"] + 1)
    interior = max(0, ki1 - ki0 + 1) * max(0, kj1 - kj0 + 1) * nzk
    total = max(0, li1 - li0 + 1) * max(0, lj1 - lj0 + 1) * nz
    return interior, total - interior


def core_region(run, world: int) -> tuple[int, int, int, int]:
    """Owned columns swept on the launch stream inside the timed brackets when the exchange is overlapped: everything but the
    boundary slab towards each neighbour (abi.cu core_region)."""
    npx, npy, me = run["nproc_x"], run["nproc_y"], run["myid"]
    idx, idy = me % npx, me // npx
    li0, li1, lj0, lj1 = 0, run["nxp"] - 1, 0, run["nyp"] - 1
    wx = 8 if run["nxp"] >= 32 else 2   # x slabs are one tile column wide (abi.cu slab_width), y slabs the two planes the messages carry
    if world > 1:
        if idx > 0: li0 += wx
        if idx < npx - 1: li1 -= wx
        if idy > 0: lj0 += 2
        if idy < npy - 1: lj1 -= 2
    return li0, li1, lj0, lj1


def bytes_per_cell(nm: int, W: int) -> dict:
    """Algorithmic (compulsory) HBM bytes per cell per sweep, SURVEY 8d / DESIGN.md."""
    return {
        "stress_interior": 3 * W + 12 * W + (16 if nm > 0 else 8) + 6 * nm * 8,
        "stress_pml": 3 * W + 12 * W + 8 + 9 * 8,
        "vel_interior": 6 * W + 6 * W + 4,
        "vel_pml": 6 * W + 6 * W + 4 + 9 * 8,
    }


def psv_bytes_per_cell(nm: int, W: int) -> dict:
    return {
        "stress_interior": 2 * W + 6 * W + (16 if nm > 0 else 8) + 3 * nm * 8,
        "stress_pml": 2 * W + 6 * W + 8 + 4 * 8,
        "vel_interior": 3 * W + 4 * W + 4,
        "vel_pml": 3 * W + 4 * W + 4 + 4 * 8,
    }


def measured_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().splitlines() if r.strip()]
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                if int(r[0]) != self.device:
                    continue
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(n)
            except Exception:
                continue
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def _omp_threads(n: int) -> int:
    """Set the OpenMP thread count of the libgomp the in-tree libraries link against; returns what is in force."""
    import ctypes

    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return n


def cpu_port_throughput(nm: int, sample=(384, 384, 384), steps: int = 24, warmup: int = 1) -> dict:
    """The reference's CPU path (oracle port, OpenMP over all host cores) on a bounded sample of the workload."""
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import Oracle   # the one place outside tests/ that may execute oracle/: the CPU baseline

    # all the host cores this process may use -- torchrun exports OMP_NUM_THREADS=1 to its workers, which must not throttle
    # the CPU arm (rank 0 runs it alone); SWPC_BENCH_CPU_THREADS overrides
    if os.environ.get("SWPC_BENCH_CPU_SAMPLE"):   # tests shrink the sample; the reported `sample` string says what was run
        sample = tuple(int(v) for v in os.environ["SWPC_BENCH_CPU_SAMPLE"].split(","))
    cores = int(os.environ.get("SWPC_BENCH_CPU_THREADS", "0")) or len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    with tempfile.TemporaryDirectory() as td:
        nx, ny, nz = sample
        inf = write_workload(Path(td), nx, ny, nz, steps + warmup, 1, 1)
        o = Oracle(inf, base_dir=td, nm=nm)
        cores = _omp_threads(cores)   # libgomp may have read the environment before this function ran
        for it in range(1, warmup + 1):
            o.step(it)
        t0 = time.perf_counter()
        for it in range(warmup + 1, warmup + steps + 1):
            o.step(it)
        dt = time.perf_counter() - t0
        o.close()
    return {"value": nx * ny * nz * steps / dt, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": f"{nx}x{ny}x{nz} sub-grid of the same layered NM={nm} PML workload, {steps} timed steps ({dt:.1f} s), "
                      "C/OpenMP restatement of the reference loops (oracle/), float64 fields",
            "seconds": dt, "ms_per_step": dt / steps * 1e3}


# ---------------------------------------------------------------------------------------------------------------------
# secondary lines (N = 1): each a separate run on the same GPU, CUDA events on the library's launch stream
def time_3d_case(inf: Path, wdir: Path, *, nm: int, fdt, K: int, Wm: int, device: int) -> dict:
    """W warm-up + K timed steps of one swpc_3d case (state resident in HBM), the per-sweep stopwatches, and the roofline
    fractions on algorithmic bytes."""
    from openswpc_b200.swpc3d import Swpc3d

    W = np.dtype(fdt).itemsize
    run = Swpc3d(inf, base_dir=wdir, nm=nm, myid=0, field_dtype=fdt)
    try:
        run.attach_device(device)
        out = {"grid": [run["nx"], run["ny"], run["nz"]], "nm": nm, "field_type": "f64" if W == 8 else "f32", "steps": K, "warmup": Wm}
        cells = run["nx"] * run["ny"] * run["nz"]
        run.run(1, Wm)
        run.device_call("swpc3d_sync")
        run.set_option("kernel_timing", 1)
        l0 = run.info("launches")
        run.timer_start()
        run.device_call("swpc3d_run", Wm + 1, Wm + K)
        ms = run.timer_stop()
        run.device_call("swpc3d_sync")
        ms_s, ms_v = run.info("ms_stress"), run.info("ms_vel")
        run.set_option("kernel_timing", 0)
        interior, pml = cell_counts(run)
        bpc = bytes_per_cell(nm, W)
        sb = interior * bpc["stress_interior"] + pml * bpc["stress_pml"]
        vb = interior * bpc["vel_interior"] + pml * bpc["vel_pml"]
        peak, _ = measured_peak()
        out.update({"value": cells * K / (ms / 1e3), "unit": "cell-updates/s", "ms_per_step": ms / K, "gpu_launches": int(run.info("launches") - l0),
                    "ms_stress": ms_s, "ms_vel": ms_v, "cells_interior": interior, "cells_absorber": pml,
                    "bytes_per_cell_domain_weighted": (sb + vb) / cells,
                    "roofline": {"bound": "hbm", "achieved": (sb + vb) / (ms / K / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": (sb + vb) / (ms / K / 1e3) / 1e9 / peak, "what": "algorithmic bytes of both sweeps / step time",
                                 "stress_frac": sb / (ms_s / 1e3) / 1e9 / peak if ms_s > 0 else None,
                                 "vel_frac": vb / (ms_v / 1e3) / 1e9 / peak if ms_v > 0 else None}})
        return out
    finally:
        run.close()


def time_psv_case(inf: Path, wdir: Path, *, nm: int, K: int, Wm: int, device: int) -> dict:
    from openswpc_b200 import _lib
    from openswpc_b200.swpc_psv import SwpcPsv

    run = SwpcPsv(inf, base_dir=wdir, nm=nm)
    try:
        run.attach_device(device)
        lib, h = run.lib, run.handle

        def call(fn, *a):
            _lib.check_psv(getattr(lib, fn)(h, *a))

        def info(key):
            v = C.c_double()
            call("swpcpsv_get_info", key.encode(), C.byref(v))
            return v.value

        run.run(1, Wm)
        call("swpcpsv_sync")
        call("swpcpsv_set_option", b"kernel_timing", 1)
        l0 = info("launches")
        call("swpcpsv_timer_start")
        call("swpcpsv_run", Wm + 1, Wm + K)
        msv = C.c_float()
        call("swpcpsv_timer_stop", C.byref(msv))
        call("swpcpsv_sync")
        ms = msv.value
        ms_s, ms_v = info("ms_stress"), info("ms_vel")
        ci, ca = info("cells_interior"), info("cells_absorber")
        bpc = psv_bytes_per_cell(nm, 8)
        sb = ci * bpc["stress_interior"] + ca * bpc["stress_pml"]
        vb = ci * bpc["vel_interior"] + ca * bpc["vel_pml"]
        peak, _ = measured_peak()
        cells = run["nx"] * run["nz"]
        return {"grid": [run["nx"], run["nz"]], "nm": nm, "field_type": "f64", "steps": K, "warmup": Wm, "value": cells * K / (ms / 1e3),
                "unit": "cell-updates/s", "ms_per_step": ms / K, "gpu_launches": int(info("launches") - l0), "ms_stress": ms_s, "ms_vel": ms_v,
                "cells_interior": int(ci), "cells_absorber": int(ca), "bytes_per_cell_domain_weighted": (sb + vb) / cells,
                "roofline": {"bound": "hbm", "achieved": (sb + vb) / (ms / K / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": (sb + vb) / (ms / K / 1e3) / 1e9 / peak, "what": "algorithmic bytes of both sweeps / step time",
                             "stress_frac": sb / (ms_s / 1e3) / 1e9 / peak if ms_s > 0 else None,
                             "vel_frac": vb / (ms_v / 1e3) / 1e9 / peak if ms_v > 0 else None}}
    finally:
        run.close()


def secondary_lines(td: Path, K: int, Wm: int, device: int, which: set) -> dict:
    out = {}

    def guard(name, fn):
        if name not in which:
            return
        t0 = time.perf_counter()
        try:
            out[name] = fn()
            out[name]["wall_s"] = time.perf_counter() - t0
        except Exception as e:   # a secondary line never voids the headline
            out[name] = {"error": f"{type(e).__name__}: {e}"[:400]}

    guard("psv", lambda: dict(time_psv_case(write_psv_workload(td / "psv", 16384, 8192, Wm + K + 2), td / "psv", nm=3, K=K, Wm=Wm, device=device),
                              workload="BASELINE configs[1]: swpc_psv 2-D P-SV viscoelastic NM=3, ADE-CFS PML na=20, layered model, 16384x8192, f64 fields"))
    guard("elastic", lambda: dict(time_3d_case(write_workload(td / "el", 512, 512, 512, Wm + K + 2, 1, 1, benchmark=True), td / "el", nm=0, fdt=np.float64,
                                               K=K, Wm=Wm, device=device),
                                  workload="BASELINE configs[2]: swpc_3d elastic NM=0, benchmark_mode homogeneous half-space (m_medium.f90:55-74), "
                                           "point source, PML na=20, 512x512x512, f64 fields"))
    guard("f32", lambda: dict(time_3d_case(write_workload(td / "f32", 1024, 1024, 512, Wm + K + 2, 1, 1), td / "f32", nm=3, fdt=np.float32,
                                           K=K, Wm=Wm, device=device),
                              workload="BASELINE configs[3] with float32 fields (the reference's MP=SP build, m_global.f90:30): NM=3, PML, lhm, 1024x1024x512"))
    return out


def weak_base_line(td: Path, K: int, Wm: int, device: int) -> dict:
    t0 = time.perf_counter()
    try:
        r = time_3d_case(write_workload(td / "wb", 512, 1024, 1024, Wm + K + 2, 1, 1, hetero=True), td / "wb", nm=3, fdt=np.float64, K=K, Wm=Wm, device=device)
        r["workload"] = ("the per-GPU workload of the N>1 runs on ONE GPU: lhm_rmed heterogeneous crust, 512x1024x1024, 1x1, NM=3, PML, f64 -- "
                         "weak-scaling efficiency at N GPUs = value(N) / (N * this value)")
        r["wall_s"] = time.perf_counter() - t0
        return r
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:400]}


# ---------------------------------------------------------------------------------------------------------------------
# N > 1: the NCCL exchange against the emulated one, bit for bit, on the boxes the driver runs
def nccl_parity(td: Path, rank: int, world: int, local: int, npx: int, npy: int, nt: int = 30) -> dict | None:
    """A small decomposed heterogeneous case (lhm_rmed, 40 x 36 columns per rank, nz = 48, NM=3, PML na=6; boundary-first overlap on,
    as in the timed run) stepped over NCCL on the N GPUs, and the same decomposition stepped as emulated ranks on rank 0's GPU
    with swpc3d_comm_local (the exchange the single-GPU parity tests tie to the oracle).  Rank 0 compares every rank's nine
    fields over the whole memory box (owned cells and halo planes), the station traces and the progress amplitudes."""
    import torch.distributed as dist

    from openswpc_b200 import _lib
    from openswpc_b200.distributed import allreduce_minmax, attach_nccl
    from openswpc_b200.swpc3d import Swpc3d

    bx, by, nz = 40, 36, 48
    nx, ny = bx * npx, by * npy
    d = td / f"parity{rank}"
    inf = write_workload(d, nx, ny, nz, nt, npx, npy, dx=0.5, dt=0.02, na=6, hetero=True)
    # the source of the bench workload sits at the centre (on rank seams for even layouts); stations on an 8 x 8 grid
    run = Swpc3d(inf, base_dir=d, nm=3, myid=rank)
    allreduce_minmax(run)
    run.attach_device(local)
    attach_nccl(run)
    vm = run.run(1, nt)
    run.write_sac(d / "out")
    mine = {"fields": run.download_fields(), "wav": run.wav() if run["nst"] else None, "vm": vm, "vmin": run["vmin"], "vmax": run["vmax"]}
    run.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    res = None
    if rank == 0:
        lib = _lib.load()
        emu = []
        for q in range(world):
            e = Swpc3d(inf, base_dir=d, nm=3, myid=q)
            e.set_minmax(mine["vmin"], mine["vmax"])
            e.attach_device(local)
            emu.append(e)
        hs = (C.c_void_p * world)(*[e.handle for e in emu])
        for it in range(1, nt + 1):
            for e in emu:
                e.device_call("swpc3d_wav_store", it)
                e.device_call("swpc3d_update_stress")
                e.device_call("swpc3d_stressglut", it)
            _lib.check(lib.swpc3d_comm_local(hs, world, 0))
            for e in emu:
                e.device_call("swpc3d_update_vel")
                e.device_call("swpc3d_bodyforce", it)
            _lib.check(lib.swpc3d_comm_local(hs, world, 1))
        nf = nw = nst = 0
        bad = []
        amp = 0.0
        for q, e in enumerate(emu):
            ref = e.download_fields()
            for n, a in ref.items():
                amp = max(amp, float(np.abs(a).max()))
                if np.array_equal(a, gathered[q]["fields"][n]):
                    nf += 1
                else:
                    bad.append(f"rank {q} field {n}")
            if e["nst"]:
                e.write_sac(d / f"emu{q}")
                nst += e["nst"]
                if np.array_equal(e.wav(), gathered[q]["wav"]):
                    nw += e["nst"] * 3
                else:
                    bad.append(f"rank {q} traces")
            e.close()
        vm_same = all(np.array_equal(g["vm"], gathered[0]["vm"]) for g in gathered)
        res = {"ok": not bad and vm_same and amp > 0, "ranks": world, "decomposition": [npx, npy], "steps": nt, "grid": [nx, ny, nz],
               "fields_bit_exact": nf, "fields_total": 9 * world, "traces_bit_exact": nw, "traces_total": 3 * nst,
               "progress_lines_identical_on_all_ranks": vm_same, "max_abs_field": amp, "mismatches": bad[:8],
               "what": "NCCL send/recv run on N GPUs (boundary-first overlap) vs the same ranks emulated on one GPU with swpc3d_comm_local; "
                       "whole memory boxes incl. halo planes, station traces, progress amplitudes"}
    box = [res]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", default="", help="nx,ny,nz per GPU (development only; default = the BASELINE workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary / weak_base / e2e_snap runs (development, profiling)")
    ap.add_argument("--secondary", default="", help="comma separated subset of psv,elastic,f32,weak_base (development)")
    ap.add_argument("--hetero", type=int, default=-1, help="1: lhm_rmed model, 0: layered lhm (default: lhm at N=1, lhm_rmed at N>1)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--opts", default="", help="key=value,... library options (development only, e.g. overlap=0)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    n_gpus = max(a.gpus, world)
    nm = 3
    from openswpc_b200.distributed import layout_for

    npx, npy = layout_for(world)
    if a.grid:
        bx, by, bz = map(int, a.grid.split(","))
    elif world == 1:
        bx, by, bz = 1024, 1024, 512
    else:
        bx, by, bz = 512, 1024, 1024
    nx, ny, nz = bx * npx, by * npy, bz
    hetero = (world > 1) if a.hetero < 0 else bool(a.hetero)
    model = ("synthetic heterogeneous crust (lhm_rmed: 8 layers x Gaussian random media, eps = 3 %)" if hetero
             else "synthetic layered model (lhm, 8 layers)")
    workload = (f"swpc_3d viscoelastic NM=3 GZB + ADE-CFS PML na=20, {model}, "
                f"{nx}x{ny}x{nz} global, {npx}x{npy} x-y decomposition, {bx}x{by}x{bz} per GPU")
    config = {"workload": workload, "grid": [nx, ny, nz], "per_gpu_grid": [bx, by, bz], "decomposition": [npx, npy], "nm": nm,
              "abc_type": "pml", "na": 20, "field_type": a.dtype, "other_arrays": "f32", "dt": 0.025, "dx": 0.5,
              "l2": "inputs_larger_than_l2 (>= 88 GB of state per GPU is streamed every step)"}

    # ------------------------------------------------------------------ reference arm: CPU port on host cores
    if a.impl == "reference":
        if rank != 0:
            return
        res = cpu_port_throughput(nm, steps=max(1, a.steps), warmup=max(1, a.warmup))
        config = dict(config, reference_sample="384x384x384 sub-grid of the workload (a bounded sample: the rate metric is size-independent on "
                                               "the CPU; the host does not change with N, so at N>1 this is still ONE host)")
        line = {"metric": "cell_updates_per_s", "value": res["value"], "unit": "cell-updates/s", "n_gpus": n_gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "reference Fortran cannot be built in this image (no gfortran/MPI); this is the C/OpenMP port of its loops"}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the swpc3d_b200 path has no CPU fallback")
    from openswpc_b200 import _lib
    from openswpc_b200.distributed import allreduce_minmax, attach_nccl, init_process_group
    from openswpc_b200.swpc3d import Swpc3d

    # host-side setup is OpenMP code: share the box's cores between the ranks instead of oversubscribing them
    # (torchrun exports OMP_NUM_THREADS=1 to its workers: set the count on the runtime itself)
    _lib.load()
    _omp_threads(int(os.environ.get("SWPC_BENCH_SETUP_THREADS", "0")) or max(1, len(os.sched_getaffinity(0)) // world))
    init_process_group("nccl" if world > 1 else None)
    import torch.distributed as dist

    fdt = np.float64 if a.dtype == "f64" else np.float32
    W = np.dtype(fdt).itemsize
    K, Wm = a.steps, max(a.warmup, 3)
    nt = Wm + 3 * K + 12
    td = tempfile.TemporaryDirectory()

    parity = None
    if world > 1:
        parity = nccl_parity(Path(td.name), rank, world, local, npx, npy)
        if not parity or not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": "cell_updates_per_s", "value": None, "n_gpus": world, "parity": parity,
                                  "error": "NCCL run differs from the emulated decomposition"}), flush=True)
            raise SystemExit(3)

    wdir = Path(td.name) / f"rank{rank}"
    inf = write_workload(wdir, nx, ny, nz, nt, npx, npy, hetero=hetero, extra=SNAP_BLOCK)   # (snapshot files only exist once snap_open is called)
    t_setup = time.perf_counter()
    run = Swpc3d(inf, base_dir=wdir, nm=nm, myid=rank, field_dtype=fdt)
    allreduce_minmax(run)
    t_host = time.perf_counter() - t_setup
    t_up = time.perf_counter()
    run.attach_device(local)
    attach_nccl(run)
    run.device_call("swpc3d_sync")
    t_upload = time.perf_counter() - t_up
    for kv in filter(None, a.opts.split(",")):
        run.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    def barrier():
        run.device_call("swpc3d_sync")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # warm-up (also through the public API, so that every code path is warm -- including the first NCCL max-reduce of
    # report__progress, whose lazy connection set-up would otherwise land in the timed e2e region)
    run.run(1, Wm)
    vtmp = (C.c_float * 3)()
    run.device_call("swpc3d_vmax_global", vtmp)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- timed region 1: K steps, state resident in HBM, CUDA events on the library's launch stream
    run.set_option("kernel_timing", 1)
    l0 = run.info("launches")
    barrier()
    run.timer_start()
    run.device_call("swpc3d_run", Wm + 1, Wm + K)
    ms = run.timer_stop()
    barrier()
    launches = run.info("launches") - l0
    ms_stress, ms_vel = run.info("ms_stress"), run.info("ms_vel")
    ms_halo, n_halo, halo_bytes = run.info("ms_halo"), run.info("n_halo"), run.info("halo_bytes")
    run.set_option("kernel_timing", 0)
    # the halo phase on its own (nothing else on the GPU): 10 velocity exchanges, idempotent on an up-to-date halo
    ms_halo_alone = ms_halo_alone_nccl = 0.0
    p2p_on = 0.0
    if world > 1:
        p2p_on = run.info("p2p_ok")

        def exchange_alone():
            barrier()
            for _ in range(3):   # untimed: the first exchange after a host barrier waits out the ranks' launch skew
                run.device_call("swpc3d_comm_vel")
            run.set_option("kernel_timing", 1)
            for _ in range(10):
                run.device_call("swpc3d_comm_vel")
            barrier()
            v = run.info("ms_halo")
            run.set_option("kernel_timing", 0)
            return v

        ms_halo_alone = exchange_alone()
        if p2p_on:   # the library path (pack + ncclSend/ncclRecv + unpack) beside it, for the record
            run.set_option("p2p", 0)
            exchange_alone()
            ms_halo_alone_nccl = exchange_alone()
            run.set_option("p2p", 1)

    # ---- timed region 2 (e2e): the call a user makes -- Swpc3d.run() + waveform read-back, host wall clock.
    # Every step the host evaluates the moment-rate values and copies them to the device; every ntdec_r steps the
    # max amplitudes come back; at the end the station traces are read to host buffers and written as SAC.
    barrier()
    t0 = time.perf_counter()
    vm = run.run(Wm + K + 1, Wm + 2 * K)
    nfiles = run.write_sac(wdir / "out")
    run.device_call("swpc3d_sync")
    t_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 3 (e2e_snap): the same public call with the snapshot files of the reference's example open (example/input.inf:55-83:
    # xz / ob sections x ps / v / u, every 5 steps, decimation 2, netCDF): slice kernels every step, and at output steps the reduction onto
    # the I/O ranks, the copies to the host and the file records -- all of it beside the sweeps (asynchronous snapshot path)
    t_snap = None
    if not a.no_secondary:
        run.snap_open(Path(td.name) / "snap")      # the I/O ranks of m_snap.f90:163-191 write
        run.run(Wm + 2 * K + 1, Wm + 2 * K + 10)   # warm-up: two output steps (first use of the snapshot communicator, pinned buffers, files)
        barrier()
        t0 = time.perf_counter()
        run.run(Wm + 2 * K + 11, Wm + 3 * K + 10)
        run.device_call("swpc3d_sync")
        t_snap = time.perf_counter() - t0
        barrier()
        t1 = time.perf_counter()
        run.snap_close()
        t_snap_close = time.perf_counter() - t1

    nsrc, nst, ntw = run["nsrc"], run["nst"], run["ntw"]
    h2d = 4.0 * nsrc
    d2h = (12.0 * len(vm) + 4.0 * 3 * nst * ntw) / K
    interior, pml = cell_counts(run)
    # with the boundary-first overlap the sweep stopwatches bracket the CORE region on the launch stream (the boundary slabs
    # run on the exchange stream): count the cells those brackets cover
    overlapped = world > 1 and "overlap=0" not in a.opts
    c_interior, c_pml = cell_counts(run, core_region(run, world)) if overlapped else (interior, pml)
    exposed_rank = ms / K - ms_stress - ms_vel   # this rank's step time outside its two sweep brackets (not clamped)
    stats = torch.tensor([ms, t_e2e * 1e3, launches, h2d, d2h, float(interior), float(pml), ms_stress, ms_vel, ms_halo,
                          halo_bytes / max(n_halo, 1.0), ms_halo_alone, exposed_rank, ms_halo_alone_nccl,
                          (t_snap or 0.0) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = stats
    mx, sm = mx.cpu().numpy(), sm.cpu().numpy()
    run.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    cells = nx * ny * nz
    ms_max, e2e_ms_max = float(mx[0]), float(mx[1])
    value = cells * K / (ms_max / 1e3)
    e2e_value = cells * K / (e2e_ms_max / 1e3)
    bpc = bytes_per_cell(nm, W)
    # roofline of the dominant kernel (fused stress sweep) on rank 0: algorithmic bytes of ONE launch / its mean duration
    stress_bytes = c_interior * bpc["stress_interior"] + c_pml * bpc["stress_pml"]
    vel_bytes = c_interior * bpc["vel_interior"] + c_pml * bpc["vel_pml"]
    step_bytes = (interior * (bpc["stress_interior"] + bpc["vel_interior"]) + pml * (bpc["stress_pml"] + bpc["vel_pml"]))
    peak, peak_src = measured_peak()
    achieved = stress_bytes / (ms_stress / 1e3) / 1e9 if ms_stress > 0 else None
    traffic, traffic_source = None, None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            key = f"{bx}x{by}x{bz}_{a.dtype}_nm{nm}"
            traffic = tj.get(key, {}).get("stress_dram_bytes_per_launch")
            if traffic is not None:
                traffic_source = ("ncu dram__bytes_read.sum + dram__bytes_write.sum of the sweep's launches, NOT measured in this run: "
                                  + str(tj[key].get("source", "profiles/traffic.json")))
        except Exception:
            traffic = None
    line = {
        "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": config,
        "roofline": {"bound": "hbm", "kernel": "fused stress sweep: stress_tma<F,NM=3> (interior tiles) + absorber-shell kernels", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_source,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": stress_bytes, "ms_per_launch": ms_stress,
                     "bytes_per_cell": bpc, "cells_interior": c_interior, "cells_absorber": c_pml,
                     "cells_note": ("cells of the core region the stopwatch brackets cover (boundary slabs run on the exchange stream)"
                                    if overlapped else "all owned cells of rank 0")},
        "roofline_step": {"achieved": step_bytes / (ms_max / K / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": step_bytes / (ms_max / K / 1e3) / 1e9 / peak, "ms_vel_per_launch": ms_vel,
                          "vel_achieved": vel_bytes / (ms_vel / 1e3) / 1e9 if ms_vel > 0 else None,
                          "note": "per-GPU algorithmic bytes of both sweeps (all owned cells of rank 0) / step time of the slowest rank"},
        "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": float(sm[3]), "d2h_bytes_per_step": float(sm[4]),
                "ms_per_step": e2e_ms_max / K, "sac_files": nfiles,
                "what": "Swpc3d.run() (host-evaluated source terms H2D every step, max-amplitude D2H every ntdec_r steps) + station "
                        "traces D2H + SAC write; fields stay device-resident as in the reference's `!$acc enter data` design",
                "one_time_upload_s": t_upload, "host_setup_s": t_host},
        "halo": None if world == 1 else {
            "what": ("one exchange of one field family (2 per step) = push kernels that store the face planes straight into the neighbours' receive "
                     "buffers over NVLink (CUDA IPC peer memory, release/acquire flags) + wait + pull kernels" if p2p_on else
                     "one exchange of one field family (2 per step) = pack kernels + ncclSend/ncclRecv with up to 4 neighbours + unpack kernels") +
                    "; CUDA events on the exchange stream; it runs beside the core sweep (boundary-first overlap)",
            "transport": "peer-to-peer stores (halo_push / halo_wait / halo_pull)" if p2p_on else "NCCL send/recv",
            "ms_per_exchange_alone_nccl_path_max": float(mx[13]) if p2p_on else None,
            "ms_per_exchange_overlapped_max": float(mx[9]), "ms_per_exchange_alone_max": float(mx[11]),
            "bytes_sent_per_exchange_max": float(mx[10]),
            "achieved_GBs_per_direction": float(mx[10]) / (float(mx[11]) / 1e3) / 1e9 if mx[11] > 0 else None,
            "nvlink_peak_GBs_per_direction": 900.0,
            "nvlink_frac": float(mx[10]) / (float(mx[11]) / 1e3) / 1e9 / 900.0 if mx[11] > 0 else None,
            "nvlink_frac_note": "bytes sent by the busiest rank / time of the exchange run alone (every kernel of the exchange included) / 900 GB/s",
            "outside_sweeps_ms_per_step_max": float(mx[12]),
            "outside_sweeps_note": "max over ranks of (step - core stress sweep - core velocity sweep) on that rank, not clamped: source, "
                                   "station and join overhead plus any exchange tail the core sweep did not cover; the exposed cost of "
                                   "the exchange proper is ms_per_step here minus weak_base.ms_per_step of the N=1 line (same per-GPU workload)",
            "overlap": "boundary-first: exchange stream overlaps the core sweeps"},
        "gpu_launches": int(sm[2]),
        "progress_lines": [[float(x) for x in r] for r in vm[-2:]],
    }
    if t_snap is not None:
        line["e2e_snap"] = {"value": cells * K / (float(mx[14]) / 1e3), "unit": "cell-updates/s", "ms_per_step": float(mx[14]) / K, "close_s": t_snap_close,
                            "what": "Swpc3d.run() with the snapshot files of example/input.inf:55-83 open (xz / ob sections x ps / v / u, ntdec_s = 5, decimation 2, "
                                    "netCDF): slice kernels, reduction onto the I/O ranks, D2H and file records included, on the headline run's own state; "
                                    "compare ms_per_step with e2e.ms_per_step"}
    if parity is not None:
        line["parity"] = parity
    line["clocks"] = clocks
    if world == 1 and not a.no_secondary and not a.grid:
        which = set(filter(None, a.secondary.split(","))) or {"psv", "elastic", "f32", "weak_base"}
        Ks, Ws = min(K, 20), 3
        sampler2 = ClockSampler(local)
        sampler2.start()
        line["secondary"] = secondary_lines(Path(td.name), Ks, Ws, local, which)
        if "weak_base" in which:
            line["weak_base"] = weak_base_line(Path(td.name), Ks, Ws, local)
        line["secondary"]["clocks"] = sampler2.stop()
    if world == 1 and not a.no_cpu_baseline:
        try:
            line["cpu_baseline"] = {k: v for k, v in cpu_port_throughput(nm).items() if k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:   # the baseline is informative; never let it void the GPU measurement
            line["cpu_baseline"] = {"value": None, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    td.cleanup()


if __name__ == "__main__":
    main()
