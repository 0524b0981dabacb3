"""Kernel-level timing probe (development tool): synthetic layered medium built with numpy, PML shell, one source.
Times the fused stress and velocity sweeps separately with CUDA events on the launch stream.

  python scripts/perf_probe.py --nx 512 --ny 512 --nz 256 --nm 3 --dtype f64 --tk 128 --ti 2 --jlen 32
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openswpc_b200.device import DeviceRank, RankGeometry  # noqa: E402


def damping(x, H, xb, xe, na, fcut, dt):
    R0 = 10.0 ** (-(np.log10(na) - 1) / np.log10(2.0) - 3.0)
    d0 = -(1.0 / (2.0 * H)) * 2 * 6.0 * np.log(R0)
    xx = np.where(x <= xb + H, (xb + H) - x, np.where(x >= xe - H, x - (xe - H), 0.0))
    q = np.abs(xx / H)
    d, a, b = d0 * q, np.pi * fcut * (1 - q), 1 + 6.0 * q * q
    den = 1 + dt / 2 * (a + d / b)
    return np.stack([((1 + dt / 2 * a) / b) / den, (-1 / b) / den, (1 - dt / 2 * (a + d / b)) / den, (d / b) / den], axis=1).astype(np.float32)


def make_rank(nx, ny, nz, nm, dtype, abc="pml", na=20, dx=0.5, dt=0.025, device=0):
    geom = RankGeometry(nx=nx, ny=ny, nz=nz, nproc_x=1, nproc_y=1, myid=0, ibeg=1, iend=nx, jbeg=1, jend=ny,
                        ibeg_k=na + 1 if abc == "pml" else 1, iend_k=nx - na if abc == "pml" else nx,
                        jbeg_k=na + 1 if abc == "pml" else 1, jend_k=ny - na if abc == "pml" else ny, kbeg_k=1,
                        kend_k=nz - na if abc == "pml" else nz, na=na)
    ts = np.array([7.9577475, 0.79577476, 0.07957747], dtype=np.float32)[:nm]
    dev = DeviceRank(geom, dx=dx, dy=dx, dz=dx, dt=dt, nm=nm, abc_type=abc, ts=ts, field_dtype=dtype, device=device)
    nym, nxm, nzm = geom.shape3
    k = np.arange(-2, nz + 4)
    z = -10.0 + (k - 0.5) * dx
    rho1 = np.where(z < 0, 0.001, 2.3 + 0.02 * np.clip(z, 0, 50)).astype(np.float32)
    vs1 = np.where(z < 0, 0.0, 3.1 + 0.03 * np.clip(z, 0, 50)).astype(np.float32)
    vp1 = np.where(z < 0, 0.0, 5.5 + 0.05 * np.clip(z, 0, 50)).astype(np.float32)
    mu1 = rho1 * vs1 * vs1
    lam1 = rho1 * (vp1 * vp1 - 2 * vs1 * vs1)
    tp1 = np.where(z < 0, 0.3, 0.008).astype(np.float32)

    def bc(a):
        return np.broadcast_to(a.astype(np.float32), (nym, nxm, nzm))

    kfs = int(np.sum(z < 0) - 3)  # last air cell (1-based k)
    m2 = np.full((nym, nxm), kfs, dtype=np.int32)
    dev.upload_medium(bc(rho1), bc(lam1), bc(mu1), bc(tp1), bc(2 * tp1), m2, m2, np.maximum(m2 - 2, 1), np.minimum(m2 + 2, nz),
                      np.maximum(m2 - 2, 1), np.minimum(m2 + 2, nz))
    if abc == "pml":
        H = na * dx
        xc = -nx * dx / 2 + (np.arange(1, nx + 1) - 0.5) * dx
        yc = -ny * dx / 2 + (np.arange(1, ny + 1) - 0.5) * dx
        zc = -10.0 + (np.arange(1, nz + 1) - 0.5) * dx
        f = lambda c, b, e: (damping(c, H, b, e, na, 0.25, dt), damping(c + dx / 2, H, b, e, na, 0.25, dt))
        gx, gy, gz = f(xc, -nx * dx / 2, nx * dx / 2), f(yc, -ny * dx / 2, ny * dx / 2), f(zc, -10.0, -10.0 + nz * dx)
        dev.setup_pml(gx[0], gx[1], gy[0], gy[1], gz[0], gz[1])
    else:
        one = lambda n: np.ones(n, dtype=np.float32)
        dev.setup_cerjan(one(nxm), one(nxm), one(nym), one(nym), one(nzm), one(nzm))
    dev.set_sources(np.array([[nx // 2, ny // 2, kfs + 20]]), np.array([1.0]), np.array([[0.58, 0.58, 0.58, 0.1, 0.2, 0.3]]),
                    np.array([[0.1, 4.0]]), stftype="kupper")
    return dev, geom


def bytes_per_cell(nm, W, pml_frac):
    interior = (3 * W + 12 * W + (16 if nm > 0 else 8) + 6 * nm * 8) + (6 * W + 6 * W + 4)
    pml = (3 * W + 12 * W + 8 + 72) + (12 * W + 4 + 72)
    return (1 - pml_frac) * interior + pml_frac * pml


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--ny", type=int, default=512)
    ap.add_argument("--nz", type=int, default=256)
    ap.add_argument("--nm", type=int, default=3)
    ap.add_argument("--na", type=int, default=20)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--abc", default="pml")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--configs", default="128,2,32", help="semicolon separated tk,ti,jlen[,pf] tuples")
    ap.add_argument("--opts", default="", help="comma separated key=value library options applied before every config")
    a = ap.parse_args()
    dtype = np.float64 if a.dtype == "f64" else np.float32
    dev, geom = make_rank(a.nx, a.ny, a.nz, a.nm, dtype, abc=a.abc, na=a.na)
    ncell = a.nx * a.ny * a.nz
    W = np.dtype(dtype).itemsize
    pml_frac = 1 - ((a.nx - 2 * a.na) * (a.ny - 2 * a.na) * (a.nz - a.na)) / ncell if a.abc == "pml" else 0.0
    bpc = bytes_per_cell(a.nm, W, pml_frac)
    for kv in filter(None, a.opts.split(",")):
        dev.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    for cfg in a.configs.split(";"):
        vals = list(map(int, cfg.split(",")))
        tk, ti, jlen, pf = (vals + [0])[:4]
        if len(vals) > 4:
            dev.set_option("tma_jl", vals[4])
        if len(vals) > 5:
            dev.set_option("tma", vals[5])
        dev.set_option("tk", tk)
        dev.set_option("ti", ti)
        dev.set_option("jlen", jlen)
        dev.set_option("pf", pf)
        for it in range(1, 4):
            dev.step(it)
        dev.sync()
        dev.timer_start()
        for _ in range(a.steps):
            dev.update_stress()
        ms_s = dev.timer_stop() / a.steps
        dev.timer_start()
        for _ in range(a.steps):
            dev.update_vel()
        ms_v = dev.timer_stop() / a.steps
        dev.timer_start()
        for it in range(4, 4 + a.steps):
            dev.step(it)
        ms_t = dev.timer_stop() / a.steps
        print(json.dumps({"grid": [a.nx, a.ny, a.nz], "nm": a.nm, "dtype": a.dtype, "abc": a.abc, "tk": tk, "ti": ti, "jlen": jlen, "pf": pf, "opts": a.opts,
                          "ms_stress": round(ms_s, 3), "ms_vel": round(ms_v, 3), "ms_step": round(ms_t, 3),
                          "gcells_s": round(ncell / ms_t / 1e6, 3), "bytes_per_cell": round(bpc, 1),
                          "GBs": round(ncell * bpc / ms_t / 1e6, 1), "vmax": [float(x) for x in dev.vmax()]}), flush=True)


if __name__ == "__main__":
    main()
