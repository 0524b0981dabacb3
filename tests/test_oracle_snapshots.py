"""Oracle snapshot products (oracle/ora_snap.c): self-consistency on the CPU -- the decomposed run assembles the same
global slices as the single-rank run, the v product is the velocity field at the slice, u integrates v."""
import numpy as np

import oracle_lib as OL
from helpers import write_case
from oracle_lib import Oracle
from snap_helpers import snap_extra


def test_snapshots_independent_of_decomposition(tmp_path):
    nt = 13
    a = Oracle(write_case(tmp_path / "a", nt=nt, extra=snap_extra(2, 3, 2, 4)), base_dir=tmp_path / "a", nm=3)
    b = Oracle(write_case(tmp_path / "b", nt=nt, nproc_x=2, nproc_y=2, extra=snap_extra(2, 3, 2, 4)), base_dir=tmp_path / "b", nm=3)
    a.run(1, nt)
    b.run(1, nt)
    assert OL.snap_info(a) == OL.snap_info(b)
    for q in range(15):
        ra, ia = OL.snap_records(a, q)
        rb, ib = OL.snap_records(b, q)
        assert ia == ib == [1, 5, 9, 13]
        np.testing.assert_array_equal(ra, rb)
        for m in range(6 if q // 3 in (0, 3, 4) else 3):
            np.testing.assert_array_equal(OL.snap_medium(a, q, m), OL.snap_medium(b, q, m))
        if q // 3 in (3, 4) and q % 3:
            np.testing.assert_array_equal(OL.snap_max(a, q), OL.snap_max(b, q))


def test_v_slice_is_the_field(tmp_path):
    nt = 9
    o = Oracle(write_case(tmp_path, nt=nt, extra=snap_extra(2, 2, 2, 4)), base_dir=tmp_path, nm=3)
    for it in range(1, nt + 1):
        o.step(it)
        if it == 8:                     # fields entering step 9 = what snap__write(9) samples
            vx = o.field(0, "Vx")
    info = OL.snap_info(o)
    recs, its = OL.snap_records(o, 0 * 3 + 1)        # xy.v
    assert its == [1, 5, 9]
    sc = np.float32(o.cfg("UC")) * np.float32(o.cfg("M0"))
    k0 = info["k0_xy"]
    ii = np.arange(1, info["nxs"] + 1) * 2 - 1
    jj = np.arange(1, info["nys"] + 1) * 2 - 1
    ref = (vx[np.ix_(jj + 2, ii + 2, [k0 + 2])][:, :, 0] * np.float64(sc)).astype(np.float32)
    np.testing.assert_allclose(recs[2, 0], ref, rtol=1e-6, atol=0)
    assert np.abs(ref).max() > 0
