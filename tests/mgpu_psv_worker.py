"""Worker of tests/test_multi_gpu.py::test_psv_nccl: one rank of a torchrun job running swpc_psv through the product's host
driver with the NCCL column exchange and the snapshot reduce, checked against the oracle's emulated decomposition."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from psv_oracle import PsvOracle, psv_case_text, write_psv_files  # noqa: E402
from openswpc_b200.distributed import allreduce_minmax, attach_nccl_psv, init_process_group  # noqa: E402
from openswpc_b200.swpc_psv import SwpcPsv  # noqa: E402


def main():
    work, nt = Path(sys.argv[1]), int(sys.argv[2])
    rank, world, local = init_process_group("nccl")
    d = work / f"r{rank}"
    d.mkdir(parents=True, exist_ok=True)
    write_psv_files(d, sources=["10.2 0.0 4.2 0.05 0.6 1e15 0.7 0.0 -0.3 0.0 0.5 0.0", "-8.4 0.0 7.0 0.2 0.8 5e14 0.1 0.0 0.9 0.0 -0.4 0.0"])
    inf = d / "input.inf"
    inf.write_text(psv_case_text(nt=nt, nx=100, nproc_x=world, products="v,u",
                                 extra=" snp_format = 'netcdf'\n xz_ps%sw = .true.\n xz_v%sw = .true.\n xz_u%sw = .true.\n idec = 2\n kdec = 2\n ntdec_s = 5"))
    run = SwpcPsv(inf, base_dir=d, nm=3, myid=rank)
    allreduce_minmax(run)
    run.attach_device(local)
    attach_nccl_psv(run)
    run.snap_open(work / "snap")
    vm = run.run(1, nt)
    run.snap_close()
    run.write_wav(d / "out")
    o = PsvOracle(inf, base_dir=d, nm=3)
    vm_ref = o.run(1, nt)
    np.testing.assert_array_equal(vm, vm_ref)
    got = run.download_fields()
    r = o.rank(rank)
    nxo, nz = r["iend"] - r["ibeg"] + 1, o.cfg("nz")
    for n, a in got.items():
        ref = o.field(rank, n)
        assert np.array_equal(a[3:3 + nxo, 3:3 + nz], ref[3:3 + nxo, 3:3 + nz]), (rank, n)
        if n in ("Sxx", "Sxz", "Vx", "Vz"):
            assert np.array_equal(a[1:5 + nxo, 3:3 + nz], ref[1:5 + nxo, 3:3 + nz]), (rank, n, "halo")
    if run["nst"]:
        np.testing.assert_array_equal(run.wav(0), o.wav(rank, 0).reshape(run.wav(0).shape))
    import torch.distributed as dist

    dist.barrier()
    from scipy.io import netcdf_file

    tags, names = ("ps", "v", "u"), (("divergence", "rotation"), ("Vx", "Vz"), ("Ux", "Uz"))
    for p in range(rank, 3, world):
        recs, its = o.snap_records(p)
        with netcdf_file(str(work / "snap" / f"psvtest.psv.xz.{tags[p]}.nc"), "r", mmap=False) as f:
            for m, name in enumerate(["rho", "lambda", "mu"]):
                np.testing.assert_array_equal(f.variables[name][:], o.snap_medium(m), err_msg=name)
            assert f.variables["t"].shape[0] == len(its) > 1
            for v, name in enumerate(names[p]):
                np.testing.assert_array_equal(f.variables[name][:], recs[:, v], err_msg=name)
        assert np.abs(recs).max() > 0
    print(f"rank {rank}/{world} psv ok: nst={run['nst']} nsrc={run['nsrc']}", flush=True)


if __name__ == "__main__":
    main()
