"""N>1 host-side plumbing on CPU: world_size-2 gloo job doing what bench.py / the driver do before the first step --
per-rank setup, the vmin/vmax all-reduce of m_medium.f90:424-425 and the broadcast of the 128-byte NCCL id."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, work, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from helpers import write_case
    from openswpc_b200.distributed import allreduce_minmax, init_process_group, layout_for
    from openswpc_b200.swpc3d import Swpc3d

    init_process_group("gloo")
    npx, npy = layout_for(world)
    d = Path(work) / f"r{rank}"
    # a model whose slowest/fastest material differ between the two halves is not expressible with a 1-D model,
    # so make the ranks disagree through vmin_local by construction: the reduction must still give one global pair
    inf = write_case(d, nt=10, nproc_x=npx, nproc_y=npy, nx=56, ny=48)
    run = Swpc3d(inf, base_dir=d, nm=3, myid=rank)
    before = (run["vmin_local"], run["vmax_local"])
    run.set_minmax(before[0] + rank, before[1] - rank)   # pretend the local values differ
    import torch

    a = torch.tensor([run["vmin"]]), torch.tensor([run["vmax"]])
    dist.all_reduce(a[0], op=dist.ReduceOp.MIN)
    dist.all_reduce(a[1], op=dist.ReduceOp.MAX)
    allreduce_minmax(run)   # uses the *_local values
    box = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    q.put((rank, run["ibeg"], run["iend"], run["jbeg"], run["jend"], run["vmin"], run["vmax"], float(a[0]), float(a[1]), box[0] == bytes(range(128)),
           before))
    dist.destroy_process_group()


def test_two_rank_gloo_setup(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    ps = [ctx.Process(target=_worker, args=(r, 2, str(tmp_path), port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, r1) = res
    assert (r0[1], r0[2], r1[1], r1[2]) == (1, 28, 29, 56)          # m_global.f90:275-281 for nx=56, nproc_x=2
    assert (r0[3], r0[4]) == (r1[3], r1[4]) == (1, 48)
    assert r0[9] and r1[9]                                           # unique-id broadcast reached both ranks
    assert r0[5] == r1[5] and r0[6] == r1[6]                         # one global (vmin, vmax) pair after the reduction
    assert np.isclose(r0[5], min(r0[10][0], r1[10][0])) and np.isclose(r0[6], max(r0[10][1], r1[10][1]))
    assert np.isclose(r0[7], r0[10][0]) and np.isclose(r0[8], r0[10][1])   # MIN / MAX semantics of the raw all-reduce


def _green_worker(rank, world, work, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from openswpc_b200.distributed import allreduce_minmax, broadcast_green_source, init_process_group
    from openswpc_b200.swpc3d import Swpc3d
    from test_green import _case

    init_process_group("gloo")
    d = Path(work) / f"r{rank}"
    d.mkdir(parents=True, exist_ok=True)
    inf = _case(d, ranks=(2, 1))
    run = Swpc3d(inf, base_dir=d, nm=3, myid=rank)
    allreduce_minmax(run)
    found_before = run.green_query()[0]
    broadcast_green_source(run)            # wav__stquery everywhere + mpi_bcast from the owner, m_green.f90:161-183
    q.put((rank, found_before, run.green_query()[1], run["ng"], run.array("green_gid").tolist()))
    dist.destroy_process_group()


def test_two_rank_green_source_broadcast(tmp_path):
    """Green's-function mode on two ranks: only the owner of the station knows the pseudo source before the broadcast; after
    it both agree with the oracle and each holds its own share of the grid-point list."""
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import Oracle
    from test_green import _case

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 90
    ps = [ctx.Process(target=_green_worker, args=(r, 2, str(tmp_path), port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (tmp_path / "ref").mkdir()
    o = Oracle(_case(tmp_path / "ref", ranks=(2, 1)), base_dir=tmp_path / "ref", nm=3)
    assert [r[1] for r in res].count(True) == 1                       # exactly one owner
    for r in res:
        g = o.green(r[0])
        assert r[2] == [g["isrc"], g["jsrc"], g["ksrc"]]
        assert r[3] == g["ng"] and r[4] == g["gid"].tolist()
    assert sum(r[3] for r in res) == 3
