import sys, tempfile
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from pathlib import Path
import numpy as np
from helpers import write_case, device_from_oracle
from oracle_lib import Oracle
def run(mp, nx, ny, nz, na, nt=12):
    d = Path(tempfile.mkdtemp())
    inf = write_case(d, nt=nt, nx=nx, ny=ny, nz=nz, na=na, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    o = Oracle(inf, base_dir=d, nm=3, mp=mp)
    fd = np.float64 if mp=='dp' else np.float32
    try:
        dev = device_from_oracle(o, 0, field_dtype=fd, device=0)
        o.run(1, nt); dev.run(1, nt); dev.sync()
    except Exception as e:
        print(mp, nx,ny,nz, 'ERROR', e); return
    got = dev.download_fields(); r = o.rank(0)
    sl = (slice(3,3+r['nyp']), slice(3,3+r['nxp']), slice(3,3+nz))
    bad = {n: float(np.abs(got[n][sl]-o.field(0,n)[sl]).max()) for n in got if not np.array_equal(got[n][sl], o.field(0,n)[sl])}
    print(mp, nx,ny,nz, 'tma_ok', dev.info('tma_ok'), 'max|Vz|', np.abs(o.field(0,'Vz')).max(), 'mismatch', bad)
import sys
cases = {'a': ('dp', 48,40,44,6), 'b': ('sp', 48,40,44,6), 'c': ('dp', 120,104,150,10), 'd': ('sp', 120,104,150,10)}
for c in sys.argv[1:]:
    run(*cases[c])
