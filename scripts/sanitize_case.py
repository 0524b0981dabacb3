"""compute-sanitizer targets (small cases, results checked against the oracle):
  python scripts/sanitize_case.py 2x2     PML NM=3, 2x2 emulated ranks: TMA stress + ring velocity + shell boxes, pack/unpack, sources, stations
  python scripts/sanitize_case.py split   one rank through swpc3d_step with the boundary-first split forced (second stream, phased sources)
  python scripts/sanitize_case.py green   Green's-function mode through the host driver (green_store / green_source kernels)
  python scripts/sanitize_case.py psv     swpc_psv: 2 emulated ranks, PML NM=3
  python scripts/sanitize_case.py elastic NM=0, one rank: the two-blocks-per-SM instantiation of stress_tma (2-stage ring)
  python scripts/sanitize_case.py sp      float32 fields, one rank: vel_ring2, 4-stage stress_tma, pml_tma<float>
  python scripts/sanitize_case.py bottom  nz = 64, na = 20, one rank, option bottom_tma = 1 (whole-line bottom tiles) + persistent stress_tma_p
Round 2: the 2x2 / split / elastic cases now run pml_tma (ticket-scheduled shell) and, for split, the tiled boundary slabs.
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from helpers import device_from_oracle, write_case  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "2x2"
d = Path(tempfile.mkdtemp())
FIELDS = ("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")


def check(o, devs, nz):
    ok = True
    for q, x in enumerate(devs):
        got, r = x.download_fields(), o.rank(q)
        sl = (slice(3, 3 + r['nyp']), slice(3, 3 + r['nxp']), slice(3, 3 + nz))
        for n in FIELDS:
            ok &= bool(np.array_equal(got[n][sl], o.field(q, n)[sl]))
    return ok


if mode == "psv":
    import test_gpu_psv as P

    o, devs = P._pair(d, 12, nx=100, nproc_x=2)
    P._step_all(o, devs, 12)
    P._compare(o, devs)
    print('bit-exact: True')
elif mode == "green":
    import test_green as G
    from openswpc_b200.swpc3d import Swpc3d

    inf = G._case(d, nt=12, cmp="x", bforce=True)
    o = Oracle(inf, base_dir=d, nm=3)
    o.run(1, 12)
    run = Swpc3d(inf, base_dir=d, nm=3)
    run.attach_device(0)
    run.run(1, 12)
    run.write_green(d / "out")
    print('bit-exact:', bool(np.array_equal(run.array("green_gf"), o.green(0)["gf"])))
elif mode == "elastic":
    inf = write_case(d, nt=6, nx=96, ny=88, nz=76, na=8, vmodel="lhm_land")
    o = Oracle(inf, base_dir=d, nm=0)
    x = device_from_oracle(o, 0, device=0)
    o.run(1, 6)
    x.run(1, 6)
    x.sync()
    print('tma_ok', x.info('tma_ok'), 'bit-exact:', check(o, [x], 76))
elif mode == "sp":
    inf = write_case(d, nt=6, nx=96, ny=88, nz=76, na=8)
    o = Oracle(inf, base_dir=d, nm=3, mp="sp")
    x = device_from_oracle(o, 0, field_dtype=np.float32, device=0)
    o.run(1, 6)
    x.run(1, 6)
    x.sync()
    print('pml items', x.info('pml_items_walls'), x.info('pml_items_bottom'), 'bit-exact:', check(o, [x], 76))
elif mode == "many":   # more work items than SMs: every persistent block walks several items (mailbox / ring reuse across items)
    inf = write_case(d, nt=4, nx=232, ny=216, nz=76, na=8)
    o = Oracle(inf, base_dir=d, nm=3)
    x = device_from_oracle(o, 0, device=0)
    if len(sys.argv) > 2:
        x.set_option(sys.argv[2], int(sys.argv[3]))
    o.run(1, 4)
    x.run(1, 4)
    x.sync()
    print('pml items', x.info('pml_items_walls'), x.info('pml_items_bottom'), 'bit-exact:', check(o, [x], 76))
elif mode == "bottom":
    inf = write_case(d, nt=6, nx=96, ny=88, nz=64, na=20)
    o = Oracle(inf, base_dir=d, nm=3)
    x = device_from_oracle(o, 0, device=0)
    x.set_option("bottom_tma", 1)
    x.set_option("tma_persist", 1)
    o.run(1, 6)
    x.run(1, 6)
    x.sync()
    print('bottom items', x.info('bottom_items'), x.info('bottom_items_vel'), 'bit-exact:', check(o, [x], 64))
elif mode == "split":
    inf = write_case(d, nt=6, nx=96, ny=88, nz=76, na=8, sources=["-23.3 -21.3 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8",
                                                                  "0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    o = Oracle(inf, base_dir=d, nm=3)
    x = device_from_oracle(o, 0, device=0)
    x.set_option("split_test", 1)
    o.run(1, 6)
    x.run(1, 6)
    x.sync()
    print('tma_ok', x.info('tma_ok'), 'bit-exact:', check(o, [x], 76))
else:
    from openswpc_b200.device import comm_local

    npx, npy = map(int, mode.split('x'))
    inf = write_case(d, nt=6, nx=96, ny=88, nz=76, na=8, nproc_x=npx, nproc_y=npy)
    o = Oracle(inf, base_dir=d, nm=3)
    devs = [device_from_oracle(o, q, device=0) for q in range(o.nranks)]
    for it in range(1, 7):
        o.step(it)
        for x in devs:
            x.wav_store(it); x.update_stress(); x.stressglut(it)
        if len(devs) > 1:
            comm_local(devs, 'stress')
        for x in devs:
            x.update_vel(); x.bodyforce(it)
        if len(devs) > 1:
            comm_local(devs, 'vel')
    print('tma_ok', [x.info('tma_ok') for x in devs], 'bit-exact:', check(o, devs, 76))
