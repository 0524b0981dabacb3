"""Development probe (not product, not a test): times the swpc_psv sweeps at a given grid with a synthetic layered medium
built directly in numpy; PML profiles come from the oracle's damping_profile helper."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import oracle_lib  # noqa: E402
from openswpc_b200.psv_device import PsvGeometry, PsvRank  # noqa: E402


def main():
    nx, nz = (int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16384,8192".split(",")))
    nm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    opts = dict(kv.split("=") for kv in sys.argv[3:])
    na, dx, dt = 20, 0.25, 0.0125
    geom = PsvGeometry(nx=nx, nz=nz, nproc_x=1, myid=0, ibeg=1, iend=nx, ibeg_k=na + 1, iend_k=nx - na, kend_k=nz - na, na=na)
    ts = np.array([3.18, 0.318, 0.0318], dtype=np.float32)[:nm]
    d = PsvRank(geom, dx=dx, dz=dx, dt=dt, nm=nm, abc_type="pml", ts=ts)
    nxm, nzm = geom.shape2
    k = np.arange(-2, nz + 4)
    air = k <= 40
    vs = np.where(air, 0.0, 2.5 + 1.5 * (k / nz))
    vp = vs * 1.73
    rho1 = np.where(air, 0.001, 2.4 + 0.4 * (k / nz)).astype(np.float32)
    col = lambda v: np.ascontiguousarray(np.broadcast_to(v.astype(np.float32), (nxm, nzm)))
    kfs = np.full(nxm, 40, dtype=np.int32)
    z0 = np.zeros(nxm, dtype=np.int32)
    d.upload_medium(col(rho1), col(rho1 * (vp * vp - 2 * vs * vs)), col(rho1 * vs * vs), col(np.full(k.shape, 0.005)), col(np.full(k.shape, 0.01)),
                    kfs, kfs, kfs + 2, z0, kfs - 2, kfs + 2)
    lib = oracle_lib.lib("dp")
    def prof(n, beg, half):
        out = np.zeros((n, 4), dtype=np.float32)
        for i in range(n):
            x = beg + (i + 0.5) * dx + (dx / 2 if half else 0)
            lib.ora_damping_profile(C.c_float(x), C.c_float(na * dx), C.c_float(beg), C.c_float(beg + n * dx), na, C.c_float(1.0), C.c_float(dt),
                                    out[i].ctypes.data_as(C.POINTER(C.c_float)))
        return out
    d.setup_pml(prof(nx, -nx / 2 * dx, False), prof(nx, -nx / 2 * dx, True), prof(nz, -10.0, False), prof(nz, -10.0, True))
    d.set_sources([nx // 2], [nz // 3], [1.0], [0.7], [0.5], [0.2], [[0.1, 2.0]])
    for k_, v in opts.items():
        d.set_option(k_, int(v))
    d.set_option("kernel_timing", 1)
    d.run(1, 5)
    d.sync()
    d.set_option("kernel_timing", 1)
    n = 30
    d.timer_start()
    d.run(6, 5 + n)
    ms = d.timer_stop() / n
    ci, ca = d.info("cells_interior"), d.info("cells_absorber")
    W = 8
    b_s = ci * (2 * W + 6 * W + 16 + 3 * nm * 8) + ca * (2 * W + 6 * W + 8 + 4 * 8)
    b_v = ci * (3 * W + 4 * W + 4) + ca * (3 * W + 4 * W + 4 + 4 * 8)
    ms_s, ms_v = d.info("ms_stress"), d.info("ms_vel")
    print(f"grid {nx}x{nz} nm={nm} {opts}: {ms:.3f} ms/step = {nx * nz / ms / 1e6:.2f} Gcell/s; stress {ms_s:.3f} ms {b_s / ms_s / 1e6:.0f} GB/s; "
          f"vel {ms_v:.3f} ms {b_v / ms_v / 1e6:.0f} GB/s; step {(b_s + b_v) / ms / 1e6:.0f} GB/s; vmax {d.vmax()}")


if __name__ == "__main__":
    main()
