"""Pin the oracle to the reference's own known answers (example/example.out -- the reference ships no tests):
header values and the 20 max-amplitude lines of the 384^3 / nt=1000 / 2x2-rank example run."""
from pathlib import Path

import numpy as np
import pytest

from example_case import EXAMPLE_OUT_LINES, es92, write_example
from oracle_lib import Oracle

GOLD = Path(__file__).resolve().parent / "golden" / "example_oracle.npz"


def test_fixture_reproduces_all_20_example_out_lines():
    """tests/golden/example_oracle.npz is the oracle's full run (make_example_golden.py, ~37 min on 8 cores); every line
    of example.out:17-36 is reproduced in all printed digits."""
    d = np.load(GOLD)
    assert int(d["nt"]) == 1000 and d["vmax_lines"].shape == (20, 3)
    for got, ref in zip(d["vmax_lines"], EXAMPLE_OUT_LINES):
        assert [es92(float(v)) for v in got] == ref.split(), (got, ref)
    # example.out:9-13
    assert f"{float(d['c']):.3f}" == "0.645" and f"{float(d['r']):.3f}" == "12.488"
    assert f"{float(d['vmin']):.3f}" == "3.122" and f"{float(d['vmax']):.3f}" == "7.977" and f"{float(d['fmax']):.3f}" == "0.500"
    assert list(d["station_names"]) == ["st01", "st02", "st03"]
    np.testing.assert_array_equal(d["station_ijk"], [[192, 192, 21], [172, 182, 21], [212, 202, 21]])


def test_example_header_and_indexing_live(tmp_path):
    """setup chain on the example's parameters: c, r, vmin, vmax, fmax (example.out:9-13) and the hand-derivable
    indices of SURVEY 8c-3."""
    inf = write_example(tmp_path, nt=10)
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    assert f"{o.cfg('c'):.3f}" == "0.645" and f"{o.cfg('r'):.3f}" == "12.488"
    assert f"{o.cfg('vmin'):.3f}" == "3.122" and f"{o.cfg('vmax'):.3f}" == "7.977" and f"{o.cfg('fmax'):.3f}" == "0.500"
    assert o.nranks == 4 and o.cfg("ntw") == 2
    assert [o.rank(q)[k] for q in range(4) for k in ("ibeg", "iend", "jbeg", "jend")] == [1, 192, 1, 192, 193, 384, 1, 192, 1, 192, 193, 384, 193, 384, 193, 384]
    for q in range(4):   # the source sits in every rank's sleeve (m_source.f90:209-211)
        assert o.sources(q)[0].tolist() == [[192, 192, 24]]
    assert o.stations(0) [0].tolist() == [[192, 192, 21], [172, 182, 21]] and o.stations(3)[0].tolist() == [[212, 202, 21]]
    assert o.imap(0, "kfs")[10, 10] == 20 and o.imap(0, "kob")[10, 10] == 20


@pytest.mark.slow
def test_first_four_lines_on_a_sub_box(tmp_path):
    """A 112x112x96 box around the source (nothing reaches its absorber before t ~ 4 s) reproduces
    example.out:17-20 in every printed digit; ~10^6 cells x 200 steps."""
    inf = write_example(tmp_path, n=112, nz=96, nt=200, nproc_x=1, nproc_y=1)
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    vm = o.run(1, 200)
    for got, ref in zip(vm, EXAMPLE_OUT_LINES[:4]):
        assert [es92(float(v)) for v in got] == ref.split(), (got, ref)
