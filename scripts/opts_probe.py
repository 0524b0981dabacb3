"""Development probe (not product, not a test): ONE device-resident bench workload, many library option sets -- per set a
few warm-up steps, then K timed steps with the per-sweep stopwatches.

  python scripts/opts_probe.py [--grid 1024,1024,512] [--nm 3] [--dtype f64] [--steps 8] [--hetero 0] "pml_tma=0" "pml_tma=1,pml_promo_bottom=2" ...
"""
import argparse
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from openswpc_b200.swpc3d import Swpc3d  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="1024,1024,512")
    ap.add_argument("--nm", type=int, default=3)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--hetero", type=int, default=0)
    ap.add_argument("--benchmark", type=int, default=0)
    ap.add_argument("sets", nargs="*", default=[""])
    a = ap.parse_args()
    nx, ny, nz = map(int, a.grid.split(","))
    fdt = np.float64 if a.dtype == "f64" else np.float32
    W = np.dtype(fdt).itemsize
    K = a.steps
    with tempfile.TemporaryDirectory() as td:
        inf = bench.write_workload(Path(td), nx, ny, nz, 10 + (K + 3) * len(a.sets), 1, 1, hetero=bool(a.hetero), benchmark=bool(a.benchmark))
        run = Swpc3d(inf, base_dir=td, nm=a.nm, field_dtype=fdt)
        run.attach_device(0)
        interior, pml = bench.cell_counts(run)
        bpc = bench.bytes_per_cell(a.nm, W)
        sb = interior * bpc["stress_interior"] + pml * bpc["stress_pml"]
        vb = interior * bpc["vel_interior"] + pml * bpc["vel_pml"]
        peak, _ = bench.measured_peak()
        it = 1
        run.device_call("swpc3d_run", it, it + 2)
        it += 3
        for s in a.sets:
            for kv in filter(None, s.split(",")):
                run.set_option(kv.split("=")[0], int(kv.split("=")[1]))
            run.device_call("swpc3d_run", it, it + 2)
            it += 3
            run.device_call("swpc3d_sync")
            run.set_option("kernel_timing", 1)
            run.timer_start()
            run.device_call("swpc3d_run", it, it + K - 1)
            ms = run.timer_stop() / K
            it += K
            ms_s, ms_v = run.info("ms_stress"), run.info("ms_vel")
            run.set_option("kernel_timing", 0)
            print(json.dumps({"opts": s, "ms_step": round(ms, 3), "ms_stress": round(ms_s, 3), "ms_vel": round(ms_v, 3),
                              "gcell_s": round(nx * ny * nz / ms / 1e6, 3), "frac_step": round((sb + vb) / ms / 1e6 / peak, 4),
                              "frac_stress": round(sb / ms_s / 1e6 / peak, 4), "frac_vel": round(vb / ms_v / 1e6 / peak, 4),
                              "pml_items": [run.info("pml_items_walls"), run.info("pml_items_bottom"), run.info("pml_direct_boxes")]}), flush=True)
        run.close()


if __name__ == "__main__":
    main()
