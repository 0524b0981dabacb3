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
