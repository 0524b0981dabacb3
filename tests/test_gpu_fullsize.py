"""BASELINE.json's full size (1024 x 1024 x 512, NM = 3, PML, float64 fields; 95 GB of state, byte offsets beyond 2^36) on the
GPU, checked through properties that do not need an oracle run of that size:

* sub-volume equivalence: the scheme has a finite numerical domain of dependence (a cell reads 2 cells away per half step:
  at most 4 cells per time step), so until the disturbance of a source has travelled to the absorber and back, the
  stations next to it record exactly what they record in a small model cut out around the source -- and the small model is
  within the oracle's reach.  The traces must agree bit for bit although every index, stride and box of the large run differs;
* implementation independence: the boundary-first split (`split_test`) and the plain `sweep_direct` kernels (TMA and register
  ring switched off) must reproduce the default path bit for bit at this size;
* silence: stations outside the domain of dependence record exact zeros.
"""
import numpy as np
import pytest

from helpers import write_case
from openswpc_b200.swpc3d import Swpc3d
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

NT = 30
BIG = dict(nx=1024, ny=1024, nz=512)
SMALL = dict(nx=176, ny=176, nz=176)
DX = 0.5
# source position in the small model (half-cell fractions are exact in float32) and the whole-cell shift into the large one
XS, YS, ZS = 0.25, 0.75, 12.1
SHIFT_X, SHIFT_Y = 190.0, 200.0


def _shift_cells(big, small, shift, n_key):
    # i = (x - xbeg) / dx with xbeg = -n dx / 2 (helpers.write_case): cells between the two models' indices of one point
    return int(round((shift + (big[n_key] - small[n_key]) * DX / 2) / DX))


def _case(d, dims, sx, sy, far=False):
    src = [f"{XS + sx} {YS + sy} {ZS} 0.0 0.4 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"]
    st = []
    n = 0
    for ox, oy, z, mode in ((0.0, 0.0, 0.0, "obb"), (1.5, -2.0, 0.0, "fsb"), (-3.0, 1.0, 6.0, "dep"), (2.5, 2.5, 14.0, "dep"), (-1.0, -3.0, 0.0, "oba"),
                            (3.0, 0.5, 20.0, "dep"), (-2.5, -2.5, 9.5, "dep"), (0.5, 3.0, 0.0, "obb")):
        n += 1
        st.append(f"{XS + sx + ox} {YS + sy + oy} {z} n{n:02d} {mode}")
    if far:   # more than 4 NT cells away from the source in the large model: never reached
        for ox, oy, z in ((-90.0, 0.0, 5.0), (0.0, -120.0, 0.0), (-400.0, -400.0, 30.0)):
            n += 1
            st.append(f"{XS + sx + ox} {YS + sy + oy} {z} f{n:02d} dep")
    return write_case(d, nt=NT, na=20, dt=0.02, ntdec_w=1, ntdec_r=5, sources=src, stations=st, **dims)


def _free_bytes():
    import torch

    return torch.cuda.mem_get_info(0)[0]


def test_full_size_sub_volume_equivalence(tmp_path):
    if _free_bytes() < 110e9:
        pytest.skip("needs a 180 GB device")
    inf_s = _case(tmp_path / "small", SMALL, 0.0, 0.0)
    o = Oracle(inf_s, base_dir=tmp_path / "small", nm=3)
    vm_ref = o.run(1, NT)
    w_ref = o.wav(0)
    assert np.abs(w_ref).max() > 0
    inf_b = _case(tmp_path / "big", BIG, SHIFT_X, SHIFT_Y, far=True)
    di, dj = _shift_cells(BIG, SMALL, SHIFT_X, "nx"), _shift_cells(BIG, SMALL, SHIFT_Y, "ny")
    results = {}
    for name, opts in (("default", {}), ("split", {"split_test": 1}), ("direct", {"tma": 0, "vel_ring": 0})):
        run = Swpc3d(inf_b, base_dir=tmp_path / "big", nm=3)
        # the same cells, shifted by whole cells (the test is void otherwise)
        np.testing.assert_array_equal(run.array("src_ijk") - o.sources(0)[0], [[di, dj, 0]])
        np.testing.assert_array_equal(run.array("st_ijk")[:8] - o.stations(0)[0], np.tile([di, dj, 0], (8, 1)))
        assert run.array("src_ijk")[0, 0] > 850 and run.array("src_ijk")[0, 1] > 850           # the far corner: the largest offsets
        run.attach_device(0)
        for k, v in opts.items():
            run.set_option(k, v)
        vm = run.run(1, NT)
        if name == "default":
            assert run.info("tma_ok") == 1.0                  # (the tensor maps are built at the first launch)
        run.write_sac(tmp_path / ("out_" + name))            # fetches the traces from the device
        results[name] = (vm, run.wav())
        run.close()
    vm, w = results["default"]
    # reports at it = 5, 10, 15: the disturbance has not reached the small model's absorber yet (67 cells = 16.75 steps)
    np.testing.assert_array_equal(vm[:3], vm_ref[:3])
    # ... and needs 32 steps to come back to the stations, all of which are within 6 cells of the source
    assert w.shape == (11, 3, NT) and w_ref.shape == (8, 3, NT)
    np.testing.assert_array_equal(w[:8], w_ref)
    assert not w[8:].any()
    for name in ("split", "direct"):
        np.testing.assert_array_equal(results[name][0], vm, err_msg=name)
        np.testing.assert_array_equal(results[name][1], w, err_msg=name)


def test_psv_full_size_sub_volume_equivalence(tmp_path):
    """The same property for swpc_psv at BASELINE configs[1] (16384 x 8192, NM = 3, PML): the stations next to a source in the
    far corner of the large section record, bit for bit, what the oracle records in a 384 x 320 section cut out around it."""
    from openswpc_b200.swpc_psv import SwpcPsv
    from psv_oracle import PsvOracle, psv_case_text, write_psv_files

    nt = 60
    lhm = "# depth rho vp vs qp qs\n0.0 2.3 5.5 3.14 600 300\n3.0 2.4 6.0 3.55 600 300\n16.0 2.8 6.7 3.83 600 300\n"
    offs = ((0.0, 0.0, "obb"), (2.5, 3.0, "dep"), (-3.0, 9.0, "dep"), (1.0, 0.0, "fsb"), (-2.0, 14.5, "dep"), (3.0, 22.0, "dep"))

    def case(d, nx, nz, shift, far):
        d.mkdir(parents=True)
        st = [f"{0.25 + shift + ox} 0.0 {z} n{n:02d} {mode}" for n, (ox, z, mode) in enumerate(offs)]
        if far:   # beyond 4 nt cells from the source: never reached
            st += [f"{0.25 + shift - 200.0} 0.0 5.0 f01 dep", f"{0.25 + shift - 3000.0} 0.0 900.0 f02 dep"]
        write_psv_files(d, sources=[f"{0.25 + shift} 0.0 11.6 0.0 0.4 1e15 0.7 0.0 -0.3 0.0 0.5 0.0"], stations=st)
        (d / "lhm.dat").write_text(lhm)
        inf = d / "input.inf"
        inf.write_text(psv_case_text(nx=nx, nz=nz, nt=nt, na=20, vmodel="lhm", ntdec_w=1, products="v,stress",
                                     extra=f" fn_lhm = 'lhm.dat'\n xbeg = {-nx * 0.5 / 2}\n"))
        return inf

    inf_s = case(tmp_path / "small", 384, 320, 0.0, False)
    o = PsvOracle(inf_s, base_dir=tmp_path / "small", nm=3)
    vm_ref = o.run(1, nt)
    shift = 3500.0
    inf_b = case(tmp_path / "big", 16384, 8192, shift, True)
    di = int(round((shift + (16384 - 384) * 0.5 / 2) / 0.5))
    run = SwpcPsv(inf_b, base_dir=tmp_path / "big", nm=3)
    ns = len(offs)
    src_b, src_s = np.asarray(run["src_ik"]).reshape(-1, 2), np.asarray(o.sources(0)[0])
    np.testing.assert_array_equal(src_b - src_s, [[di, 0]])
    np.testing.assert_array_equal(np.asarray(run["st_ik"]).reshape(-1, 2)[:ns] - np.asarray(o.stations(0)[0]), np.tile([di, 0], (ns, 1)))
    assert src_b[0, 0] > 15000
    run.attach_device(0)
    vm = run.run(1, nt)
    run.write_wav(tmp_path / "out")
    # 172 cells to the small section's absorber = 43 steps: reports at it = 10 .. 40 are covered; the traces for 80 steps
    np.testing.assert_array_equal(vm[:4], vm_ref[:4])
    for prod in (0, 2):
        w, w_ref = run.wav(prod), o.wav(0, prod)
        assert np.abs(w_ref).max() > 0
        np.testing.assert_array_equal(w[:ns], w_ref, err_msg=f"product {prod}")
        assert not w[ns:].any()
    run.close()
