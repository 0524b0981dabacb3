#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native swpc_3d time-stepping path.

Metric (BASELINE.json): 3-D viscoelastic cell-updates/s (+ % of the HBM roofline).
  N = 1 : BASELINE configs[3] -- swpc_3d, NM=3 GZB, ADE-CFS PML (na=20), synthetic layered model (the 8-layer table
          of the reference's example/lhm.dat), 1024 x 1024 x 512 on one B200, float64 fields (reference default MP=DP).
  N > 1 : BASELINE configs[4] shape -- weak scaling, 512 x 1024 x 1024 cells per GPU, x-y decomposition 2x1 / 4x1 / 4x2,
          NCCL send/recv halo exchange overlapped with the core sweeps; synthetic heterogeneous crust = the layered table
          with Gaussian random media per layer (vmodel lhm_rmed).
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


def write_workload(d: Path, nx: int, ny: int, nz: int, nt: int, npx: int, npy: int, dx=0.5, dt=0.025, na=20, hetero=False) -> Path:
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
"""
    p = d / "input.inf"
    p.write_text(inf)
    return p


def cell_counts(run) -> tuple[int, int]:
    """(interior cells, absorber cells) of this rank (m_global.f90:334-376)."""
    nz, na = run["nz"], run["na"]
    nxk = max(0, run["iend_k"] - run["ibeg_k"] + 1)
    nyk = max(0, run["jend_k"] - run["jbeg_k"] + 1)
    nzk = max(0, run["kend_k"] - run["kbeg_k"] + 1)
    interior = nxk * nyk * nzk
    total = run["nxp"] * run["nyp"] * nz
    return interior, total - interior


def bytes_per_cell(nm: int, W: int) -> dict:
    """Algorithmic (compulsory) HBM bytes per cell per sweep, SURVEY 8d / DESIGN.md."""
    return {
        "stress_interior": 3 * W + 12 * W + (16 if nm > 0 else 8) + 6 * nm * 8,
        "stress_pml": 3 * W + 12 * W + 8 + 9 * 8,
        "vel_interior": 6 * W + 6 * W + 4,
        "vel_pml": 6 * W + 6 * W + 4 + 9 * 8,
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", default="", help="nx,ny,nz per GPU (development only; default = the BASELINE workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    hetero = world > 1
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
    nt = Wm + 2 * K + 2
    td = tempfile.TemporaryDirectory()
    wdir = Path(td.name) / f"rank{rank}"
    inf = write_workload(wdir, nx, ny, nz, nt, npx, npy, hetero=hetero)
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
    ms_halo_alone = 0.0
    if world > 1:
        run.set_option("kernel_timing", 1)
        barrier()
        for _ in range(10):
            run.device_call("swpc3d_comm_vel")
        barrier()
        ms_halo_alone = run.info("ms_halo")
        run.set_option("kernel_timing", 0)

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

    nsrc, nst, ntw = run["nsrc"], run["nst"], run["ntw"]
    h2d = 4.0 * nsrc
    d2h = (12.0 * len(vm) + 4.0 * 3 * nst * ntw) / K
    interior, pml = cell_counts(run)
    stats = torch.tensor([ms, t_e2e * 1e3, launches, h2d, d2h, float(interior), float(pml), ms_stress, ms_vel, ms_halo,
                          halo_bytes / max(n_halo, 1.0), ms_halo_alone], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = stats
    mx, sm = mx.cpu().numpy(), sm.cpu().numpy()
    if rank != 0:
        run.close()
        return

    cells = nx * ny * nz
    ms_max, e2e_ms_max = float(mx[0]), float(mx[1])
    value = cells * K / (ms_max / 1e3)
    e2e_value = cells * K / (e2e_ms_max / 1e3)
    bpc = bytes_per_cell(nm, W)
    # roofline of the dominant kernel (fused stress sweep) on rank 0: algorithmic bytes of ONE launch / its mean duration
    stress_bytes = interior * bpc["stress_interior"] + pml * bpc["stress_pml"]
    vel_bytes = interior * bpc["vel_interior"] + pml * bpc["vel_pml"]
    peak, peak_src = measured_peak()
    achieved = stress_bytes / (ms_stress / 1e3) / 1e9 if ms_stress > 0 else None
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            key = f"{bx}x{by}x{bz}_{a.dtype}_nm{nm}"
            traffic = tj.get(key, {}).get("stress_dram_bytes_per_launch")
        except Exception:
            traffic = None
    step_bytes = stress_bytes + vel_bytes
    line = {
        "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": config,
        "roofline": {"bound": "hbm", "kernel": "fused stress sweep: stress_tma<F,NM=3> (interior tiles) + sweep_direct<F,3,1> (absorber shell)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": stress_bytes, "ms_per_launch": ms_stress,
                     "bytes_per_cell": bpc, "cells_interior": interior, "cells_absorber": pml},
        "roofline_step": {"achieved": step_bytes / (ms_max / K / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": step_bytes / (ms_max / K / 1e3) / 1e9 / peak, "ms_vel_per_launch": ms_vel,
                          "vel_achieved": vel_bytes / (ms_vel / 1e3) / 1e9 if ms_vel > 0 else None,
                          "note": "per-GPU algorithmic bytes of both sweeps / step time of the slowest rank"},
        "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": float(sm[3]), "d2h_bytes_per_step": float(sm[4]),
                "ms_per_step": e2e_ms_max / K, "sac_files": nfiles,
                "what": "Swpc3d.run() (host-evaluated source terms H2D every step, max-amplitude D2H every ntdec_r steps) + station "
                        "traces D2H + SAC write; fields stay device-resident as in the reference's `!$acc enter data` design",
                "one_time_upload_s": t_upload, "host_setup_s": t_host},
        "halo": None if world == 1 else {
            "what": "one exchange = pack kernels + ncclSend/ncclRecv with up to 4 neighbours + unpack kernels of one field family "
                    "(2 exchanges per step), CUDA events on the exchange stream; it runs beside the core sweep (boundary-first overlap)",
            "ms_per_exchange_overlapped_max": float(mx[9]), "ms_per_exchange_alone_max": float(mx[11]),
            "bytes_sent_per_exchange_max": float(mx[10]),
            "achieved_GBs_per_direction": float(mx[10]) / (float(mx[11]) / 1e3) / 1e9 if mx[11] > 0 else None,
            "nvlink_peak_GBs_per_direction": 900.0,
            "nvlink_frac": float(mx[10]) / (float(mx[11]) / 1e3) / 1e9 / 900.0 if mx[11] > 0 else None,
            "nvlink_frac_note": "bytes sent by the busiest rank / time of the exchange run alone (pack + NCCL + unpack kernels included) / 900 GB/s",
            "exposed_ms_per_step": max(0.0, ms_max / K - float(mx[7]) - float(mx[8])),
            "overlap": "boundary-first: exchange stream overlaps the core sweeps"},
        "gpu_launches": int(sm[2]),
        "clocks": clocks,
        "progress_lines": [[float(x) for x in r] for r in vm[-2:]],
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            line["cpu_baseline"] = {k: v for k, v in cpu_port_throughput(nm).items() if k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:   # the baseline is informative; never let it void the GPU measurement
            line["cpu_baseline"] = {"value": None, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    run.close()
    td.cleanup()


if __name__ == "__main__":
    main()
