import sys, tempfile, pathlib
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import write_case
from openswpc_b200.swpc3d import Swpc3d
from oracle_lib import Oracle
from test_planewave import pw_extra, FIELDS
d = pathlib.Path(tempfile.mkdtemp())
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2
inf = write_case(d, nt=nt, vmodel="lhm_land", extra=pw_extra("p"))
o = Oracle(inf, base_dir=d, nm=3)
o.run(1, nt)
run = Swpc3d(inf, base_dir=d, nm=3)
run.attach_device(0)
run.run(1, nt)
got = run.download_fields()
nz = run["nz"]
for f in FIELDS:
    a, b = got[f][:, :, 3:3 + nz], o.field(0, f)[:, :, 3:3 + nz]
    bad = np.argwhere(a != b)
    print(f, len(bad), "j:", np.unique(bad[:, 0])[:12], "i:", np.unique(bad[:, 1])[:12], "k:", np.unique(bad[:, 2])[:8])
    if len(bad):
        j, i, k = bad[0]
        print("   first", j, i, k, a[j, i, k], b[j, i, k])
