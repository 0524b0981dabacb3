"""Worker of tests/test_multi_gpu.py::test_a_silent_neighbour_is_detected: two ranks; rank 1 leaves the time loop early (as a rank
does that runs into the divergence abort of m_report.f90:144-151) and stays alive; rank 0 must come back from the time loop
with an error within the communication time-out instead of waiting for ever."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from helpers import write_case  # noqa: E402
from openswpc_b200.distributed import allreduce_minmax, attach_nccl, init_process_group  # noqa: E402
from openswpc_b200.swpc3d import Swpc3d, Swpc3dHostError  # noqa: E402


def main():
    work, p2p = Path(sys.argv[1]), int(sys.argv[2])
    rank, world, local = init_process_group("nccl")
    d = work / f"r{rank}"
    inf = write_case(d, nt=200, nproc_x=2, nproc_y=1, nx=56, ny=48, ntdec_r=5, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"], title="fail")
    run = Swpc3d(inf, base_dir=d, nm=3, myid=rank)
    allreduce_minmax(run)
    run.attach_device(local)
    attach_nccl(run)
    run.set_option("p2p", p2p)
    run.set_option("comm_timeout_s", 4)
    if rank == 1:
        run.run(1, 12)                       # ... and then nothing: no exchange, no reduction
        print("rank 1 left the loop", flush=True)
        time.sleep(20)
        return
    t0 = time.time()
    try:
        run.run(1, 200)
    except Swpc3dHostError as e:
        dt = time.time() - t0
        print(f"rank 0 aborted cleanly after {dt:.1f} s: {e}", flush=True)
        assert dt < 30, dt
        assert "aborted" in str(e) or "timed out" in str(e), str(e)
        return
    raise SystemExit("rank 0 finished 200 steps although its neighbour stopped after 12")


if __name__ == "__main__":
    main()
