"""Generate tests/golden/example_oracle.npz : the oracle's run of the reference's example/input.inf
(384^3, nt=1000, 2x2 emulated ranks, NM=3, PML) -- the 20 max-amplitude triplets printed by
report__progress (compare example/example.out:17-36) and the 3-station velocity traces.

Needs /root/reference (build container only); takes ~25 min on 8 cores.  Usage:
    python tests/golden/make_example_golden.py [nt]
"""
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle_lib import Oracle  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
o = Oracle("/root/reference/example/input.inf", base_dir="/root/reference", nm=3, nt=nt)
t0 = time.time()
vm = o.run(1, nt)
wall = time.time() - t0
names, ijk, wav = [], [], []
for q in range(o.nranks):
    s_ijk, s_nm = o.stations(q)
    w = o.wav(q)
    for n, nm in enumerate(s_nm):
        names.append(nm)
        ijk.append(s_ijk[n])
        wav.append(w[n])
hdr = {k: o.cfg(k) for k in ["vmin", "vmax", "fmax", "c", "r", "M0"]}
np.savez_compressed(HERE / "example_oracle.npz", vmax_lines=vm, station_names=np.array(names), station_ijk=np.array(ijk),
                    wav=np.array(wav), nt=nt, wall_s=wall, **hdr)
print("wall %.1f s, %.3g cell-updates/s" % (wall, 384**3 * nt / wall))
for i, v in enumerate(vm):
    print("it=%07d ( %9.2E %9.2E %9.2E )" % ((i + 1) * 50, v[0], v[1], v[2]))
