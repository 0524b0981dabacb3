"""One process per GPU (torchrun): torch.distributed is only the plumbing -- it carries the 128-byte NCCL unique id
and the two setup-time reductions of the reference (vmin/vmax, m_medium.f90:424-425); the halo exchange itself
is NCCL send/recv issued by the CUDA library on its own stream (swpc3d_comm_stress / swpc3d_comm_vel)."""
from __future__ import annotations

import ctypes as C
import os

from . import _lib


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend: str | None = None):
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_minmax(run) -> None:
    """mpi_allreduce(vmin, MIN) / (vmax, MAX) of velocity_minmax (m_medium.f90:424-425)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    a = torch.tensor([run["vmin_local"]], dtype=torch.float32, device=dev)
    b = torch.tensor([run["vmax_local"]], dtype=torch.float32, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MIN)
    dist.all_reduce(b, op=dist.ReduceOp.MAX)
    run.set_minmax(float(a.item()), float(b.item()))


def broadcast_green_source(run) -> None:
    """Green's-function mode: wav__stquery on every rank + mpi_bcast from the owner (m_green.f90:161-183)."""
    import torch.distributed as dist

    if not run["green_mode"]:
        return
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    mine = run.green_query()
    allq = [None] * dist.get_world_size()
    dist.all_gather_object(allq, mine)
    owners = [q for q in allq if q[0]]
    if not owners:
        raise RuntimeError("green_stnm is not inside the model (assert, m_green.f90:165)")
    _, ijk, xyz, ll = owners[-1]   # mpi_allreduce(MAX) of the owner's rank (:168-174)
    run.green_set_source(ijk, xyz, ll)


def attach_nccl(run) -> None:
    """Create the library's own NCCL communicator: rank 0 makes the unique id, torch.distributed broadcasts it."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    lib = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        _lib.check(lib.swpc3d_nccl_unique_id(buf))
        box[0] = buf.raw
    dist.broadcast_object_list(box, src=0)
    _lib.check(lib.swpc3d_comm_init(run.handle, box[0], world, rank))


def attach_nccl_psv(run) -> None:
    """The same for a `SwpcPsv` run (swpcpsv_nccl_unique_id / swpcpsv_comm_init; 1-D decomposition along x)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    lib = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        _lib.check_psv(lib.swpcpsv_nccl_unique_id(buf))
        box[0] = buf.raw
    dist.broadcast_object_list(box, src=0)
    _lib.check_psv(lib.swpcpsv_comm_init(run.handle, box[0], world, rank))


def layout_for(world: int) -> tuple[int, int]:
    """x-y decomposition used by the benchmark: 1x1, 2x1, 4x1, 4x2."""
    return {1: (1, 1), 2: (2, 1), 4: (4, 1), 8: (4, 2)}.get(world, (world, 1))
