"""compute-sanitizer target: a small PML NM=3 case with 2x2 emulated ranks (pack/unpack), TMA + direct kernels."""
import sys, tempfile
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from pathlib import Path
import numpy as np
from helpers import write_case, device_from_oracle
from oracle_lib import Oracle
from openswpc_b200.device import comm_local
d = Path(tempfile.mkdtemp())
npx, npy = (2, 2) if len(sys.argv) < 2 else map(int, sys.argv[1].split('x'))
inf = write_case(d, nt=6, nx=96, ny=88, nz=76, na=8, nproc_x=npx, nproc_y=npy)
o = Oracle(inf, base_dir=d, nm=3)
devs = [device_from_oracle(o, q, device=0) for q in range(o.nranks)]
for it in range(1, 7):
    o.step(it)
    for x in devs:
        x.wav_store(it); x.update_stress(); x.stressglut(it)
    if len(devs) > 1: comm_local(devs, 'stress')
    for x in devs:
        x.update_vel(); x.bodyforce(it)
    if len(devs) > 1: comm_local(devs, 'vel')
ok = True
for q, x in enumerate(devs):
    got = x.download_fields(); r = o.rank(q)
    sl = (slice(3, 3 + r['nyp']), slice(3, 3 + r['nxp']), slice(3, 3 + 76))
    for n in got:
        ok &= bool(np.array_equal(got[n][sl], o.field(q, n)[sl]))
    print('rank', q, 'tma_ok', x.info('tma_ok'))
print('bit-exact:', ok)
