"""The reference's example configuration (384^3, NM=3, PML, nt=1000) on the GPU against the committed oracle fixture
and the reference's own example.out lines."""
from pathlib import Path

import numpy as np
import pytest

from example_case import EXAMPLE_OUT_LINES, es92, write_example
from helpers import rel_l2
from openswpc_b200.swpc3d import Swpc3d

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "example_oracle.npz"


def test_example_run_full(tmp_path):
    d = np.load(GOLD)
    inf = write_example(tmp_path, nt=1000, nproc_x=1, nproc_y=1)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    vm = run.run(1, 1000)
    # (1) the reference's known answers, example.out:17-36
    for got, ref in zip(vm, EXAMPLE_OUT_LINES):
        assert [es92(float(v)) for v in got] == ref.split(), (got, ref)
    # (2) station seismograms vs the oracle's full run (2x2 emulated ranks): north-star bar 1e-5 relative L2
    run.write_sac(tmp_path / "out")
    names = run.station_names()
    w = run.wav()
    for n, name in enumerate(d["station_names"]):
        i = names.index(str(name))
        for c in range(3):
            e = rel_l2(w[i, c], d["wav"][n, c])
            assert e <= 1e-5, (name, c, e)
    np.testing.assert_array_equal(vm, d["vmax_lines"])
    assert (tmp_path / "out" / "wav" / "swpc.3d.st01.Vz.sac").stat().st_size == 632 + 4 * 200


def test_example_as_shipped_with_snapshots(tmp_path):
    """example/input.inf as the reference ships it, snapshot block included (:55-83: netCDF, xz and ob sections x ps / v / u,
    every 5 steps, decimation 2): the six files of the full 384^3, 1000-step run against the oracle's records (digests in
    tests/golden/example_snap_oracle.npz -- the records themselves are 590 MB), written by the asynchronous snapshot path
    (device-side fetch on its own stream, writer thread) while the time loop goes on."""
    import hashlib

    from scipy.io import netcdf_file

    g = np.load(Path(__file__).resolve().parent / "golden" / "example_snap_oracle.npz")
    nt = int(g["nt"])
    inf = write_example(tmp_path, nt=nt, nproc_x=1, nproc_y=1, snapshots=True)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    run.snap_open(tmp_path / "snap")
    vm = run.run(1, nt)
    run.snap_close()
    np.testing.assert_array_equal(vm, g["vmax_lines"])
    vnames = {"ps": ["div", "rot_x", "rot_y", "rot_z"], "v": ["Vx", "Vy", "Vz"], "u": ["Ux", "Uy", "Uz"]}
    seen = 0
    for sec in ("xz", "ob"):
        for typ in ("ps", "v", "u"):
            name = f"{sec}_{typ}"
            with netcdf_file(str(tmp_path / "snap" / f"swpc.3d.{sec}.{typ}.nc"), "r", mmap=False) as f:
                recs = np.stack([f.variables[v][:] for v in vnames[typ]], axis=1).astype(np.float32)     # (nrec, nvar, n2, n1)
                assert list(recs.shape) == list(g[f"{name}_shape"]) and recs.shape[0] == (nt - 1) // 5 + 1
                np.testing.assert_array_equal(f.variables["t"][:], (g[f"{name}_its"].astype(np.float32) * np.float32(0.02)).astype(np.float32))
                # per-record maxima first (they say WHERE a mismatch is), then the digest of everything
                np.testing.assert_array_equal(np.abs(recs).max(axis=(2, 3)), g[f"{name}_recmax"], err_msg=name)
                assert hashlib.sha256(np.ascontiguousarray(recs).tobytes()).hexdigest() == str(g[f"{name}_sha256"]), name
                for m, mn in enumerate(["rho", "lambda", "mu"]):
                    assert hashlib.sha256(np.ascontiguousarray(f.variables[mn][:].astype(np.float32)).tobytes()).hexdigest() == str(g[f"{name}_med{m}_sha256"]), (name, mn)
                if sec == "ob" and typ != "ps":
                    mx = np.stack([f.variables[v][:] for v in ("max-V", "max-H", "max-A")]).astype(np.float32)
                    assert hashlib.sha256(np.ascontiguousarray(mx).tobytes()).hexdigest() == str(g[f"{name}_max_sha256"]), name
                assert np.abs(recs).max() > 0
            seen += 1
    assert seen == 6
