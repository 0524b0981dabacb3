"""Worker of tests/test_multi_gpu.py: one rank of a torchrun job.  Runs the product's driver on this rank's GPU with
NCCL halo exchange and checks fields, traces and progress amplitudes against the oracle's emulated decomposition."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from helpers import write_case  # noqa: E402
import oracle_lib as OL  # noqa: E402
from snap_helpers import check_file, snap_extra  # noqa: E402
from openswpc_b200.distributed import allreduce_minmax, attach_nccl, init_process_group  # noqa: E402
from openswpc_b200.swpc3d import Swpc3d  # noqa: E402
from oracle_lib import Oracle  # noqa: E402


def main():
    work = Path(sys.argv[1])
    npx, npy, nt = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    p2p = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    rank, world, local = init_process_group("nccl")
    d = work / f"r{rank}"
    inf = write_case(d, nt=nt, nproc_x=npx, nproc_y=npy, nx=56, ny=48, ntdec_r=5,
                     sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"], title="mg", extra=snap_extra(2, 3, 2, 4))
    run = Swpc3d(inf, base_dir=d, nm=3, myid=rank)
    allreduce_minmax(run)
    run.attach_device(local)
    attach_nccl(run)
    run.set_option("p2p", p2p)                     # 1: halo planes pushed straight into the neighbour's buffer over NVLink; 0: NCCL send/recv
    p2p_on = run.info("p2p_ok")
    run.snap_open(work / "snap")                   # every rank takes part; the I/O ranks of m_snap.f90:163-191 write
    vm = run.run(1, nt)
    run.snap_close()
    run.write_sac(d / "out")
    o = Oracle(inf, base_dir=d, nm=3)
    assert np.float32(run["vmin"]) == np.float32(o.cfg("vmin")) and np.float32(run["vmax"]) == np.float32(o.cfg("vmax"))
    vm_ref = o.run(1, nt)
    np.testing.assert_array_equal(vm, vm_ref)
    got = run.download_fields()
    r = o.rank(rank)
    # owned cells and the halo planes filled by the exchange (i, j margins of width 2 next to a neighbour)
    sl = (slice(3, 3 + r["nyp"]), slice(3, 3 + r["nxp"]), slice(3, 3 + o.cfg("nz")))
    for n, a in got.items():
        ref = o.field(rank, n)
        assert np.array_equal(a[sl], ref[sl]), (rank, n, float(np.abs(a[sl] - ref[sl]).max()))
        assert np.array_equal(a[:, :, 3:3 + o.cfg("nz")], ref[:, :, 3:3 + o.cfg("nz")]), (rank, n, "halo")
    if run["nst"]:
        np.testing.assert_array_equal(run.wav(), o.wav(rank))
    import torch.distributed as dist
    dist.barrier()
    for q in range(rank, 15, world):                # snapshot files, shared directory, checks spread over the ranks
        sec, typ = divmod(q, 3)
        check_file(work / "snap" / f"mg.3d.{OL.SNAP_SECTIONS[sec]}.{OL.SNAP_TYPES[typ]}.nc", o, q, "mg", run["dt"], 4)
    print(f"rank {rank}/{world} ok: nst={run['nst']} nsrc={run['nsrc']} p2p={int(p2p_on)}", flush=True)


if __name__ == "__main__":
    main()
