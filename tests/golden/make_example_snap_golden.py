"""Generate tests/golden/example_snap_oracle.npz: the oracle's run of the reference's example/input.inf AS SHIPPED -- including
its snapshot block (example/input.inf:55-83: snp_format netcdf, xz / ob sections x ps / v / u, ntdec_s = 5, decimation 2) -- on
ONE rank (the GPU test runs 1x1; in this land model the fields do not depend on the decomposition, see test_gpu_example).
The six products are 590 MB of float32, so the fixture holds digests: SHA-256 of every product's whole record array and of its
running maxima, plus per-record max |value| of every variable (exact float32, to localise a mismatch).

Needs /root/reference (build container only); ~25 min on 8 cores.  Usage:  python tests/golden/make_example_snap_golden.py [nt]
"""
import hashlib
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as OL  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
o = Oracle("/root/reference/example/input.inf", base_dir="/root/reference", nm=3, nt=nt, nproc_x=1, nproc_y=1)
t0 = time.time()
vm = o.run(1, nt)
out = {"nt": nt, "vmax_lines": vm, "wall_s": time.time() - t0}
for q in range(15):
    OL._snap_bind(o)
    if o.lib.ora_snap_nrec(o.h, q) <= 0:
        continue
    sec, typ = divmod(q, 3)
    name = f"{OL.SNAP_SECTIONS[sec]}_{OL.SNAP_TYPES[typ]}"
    recs, its = OL.snap_records(o, q)
    out[f"{name}_sha256"] = hashlib.sha256(np.ascontiguousarray(recs).tobytes()).hexdigest()
    out[f"{name}_shape"] = np.array(recs.shape)
    out[f"{name}_its"] = np.array(its)
    out[f"{name}_recmax"] = np.abs(recs).max(axis=(2, 3)).astype(np.float32)      # (nrec, nvar)
    mx = OL.snap_max(o, q)
    if mx is not None and sec in (3, 4) and typ != 0:
        out[f"{name}_max_sha256"] = hashlib.sha256(np.ascontiguousarray(mx).tobytes()).hexdigest()
    for m in range(3):
        out[f"{name}_med{m}_sha256"] = hashlib.sha256(np.ascontiguousarray(OL.snap_medium(o, q, m)).tobytes()).hexdigest()
np.savez_compressed(HERE / "example_snap_oracle.npz", **out)
print("wall %.1f s; products:" % out["wall_s"], sorted(k[:-7] for k in out if k.endswith("_sha256") and "_m" not in k))
