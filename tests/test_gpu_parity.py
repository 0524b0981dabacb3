"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

The library is built with -fmad=false, so that every operation is the same plain-IEEE operation, in the same
kind and order, as in the oracle: fields are required to agree BIT-EXACTLY wherever a single source is applied
per cell (multi-source cells use atomics whose order is free), and station traces within 1e-5 relative L2
(BASELINE.json north_star) -- in practice they are identical.
"""
import numpy as np
import pytest

from helpers import device_from_oracle, rel_l2, write_case
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

FIELDS = ("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")


def _run_pair(tmp_path, nt, *, nm=3, mp="dp", nranks=(1, 1), options=None, **case):
    inf = write_case(tmp_path, nt=nt, nproc_x=nranks[0], nproc_y=nranks[1], **case)
    o = Oracle(inf, base_dir=tmp_path, nm=nm, mp=mp)
    fd = np.float64 if mp == "dp" else np.float32
    devs = [device_from_oracle(o, q, field_dtype=fd, device=0) for q in range(o.nranks)]
    for d in devs:
        for key, val in (options or {}).items():
            d.set_option(key, val)
    from openswpc_b200.device import comm_local

    for it in range(1, nt + 1):
        o.step(it)
        for d in devs:
            d.wav_store(it)
            d.update_stress()
            d.stressglut(it)
        if len(devs) > 1:
            comm_local(devs, "stress")
        for d in devs:
            d.update_vel()
            d.bodyforce(it)
        if len(devs) > 1:
            comm_local(devs, "vel")
    return o, devs


def _compare(o, devs, exact=True, tol=0.0):
    worst = 0.0
    for q, d in enumerate(devs):
        got = d.download_fields()
        r = o.rank(q)
        # owned cells + the halo planes the exchange fills
        for n in FIELDS:
            ref = o.field(q, n)
            a = got[n].astype(np.float64)
            sl = (slice(3, 3 + r["nyp"]), slice(3, 3 + r["nxp"]), slice(3, 3 + o.cfg("nz")))
            if exact:
                assert np.array_equal(a[sl], ref[sl]), f"rank {q} field {n}: max abs diff {np.abs(a[sl] - ref[sl]).max():.3e}"
            else:
                e = rel_l2(a[sl], ref[sl])
                worst = max(worst, e)
                assert e <= tol, f"rank {q} field {n}: rel L2 {e:.3e} > {tol}"
        if d.nst:
            w, wr = d.get_wav(), o.wav(q)
            if exact:
                assert np.array_equal(w, wr)
            else:
                assert rel_l2(w, wr) <= tol
    return worst


def test_pml_nm3_single_rank_bit_exact(tmp_path):
    o, devs = _run_pair(tmp_path, 40, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    assert np.abs(o.field(0, "Vz")).max() > 0
    _compare(o, devs, exact=True)
    np.testing.assert_array_equal(devs[0].vmax() * np.float32(o.cfg("UC")) * np.float32(o.cfg("M0")), o.vmax())


def test_pml_nm0_elastic_bit_exact(tmp_path):
    o, devs = _run_pair(tmp_path, 30, nm=0, vmodel="lhm_land", sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    _compare(o, devs, exact=True)


def test_cerjan_uni_bit_exact(tmp_path):
    o, devs = _run_pair(tmp_path, 30, abc_type="cerjan", vmodel="uni", sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    _compare(o, devs, exact=True)


def test_two_sources_tolerance(tmp_path):
    # two sources, different cells -> still exact; kept as a tolerance test to document the bar
    o, devs = _run_pair(tmp_path, 40)
    worst = _compare(o, devs, exact=False, tol=1e-12)
    assert worst <= 1e-12


@pytest.mark.parametrize("layout", [(2, 1), (2, 2), (3, 2)])
def test_decomposed_matches_oracle_bit_exact(tmp_path, layout):
    o, devs = _run_pair(tmp_path, 30, nranks=layout, nx=50, ny=44, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    _compare(o, devs, exact=True)


def test_bodyforce_mode_bit_exact(tmp_path):
    o, devs = _run_pair(tmp_path, 30, bf_mode=True, sources=["0.3 -0.2 4.1 0.05 0.6 1e12 2e12 -3e12"])
    _compare(o, devs, exact=True)


def test_single_precision_fields(tmp_path):
    # MP=SP build of the reference (m_global.f90:30) against the float32-field CUDA instantiation
    o, devs = _run_pair(tmp_path, 30, mp="sp", sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    _compare(o, devs, exact=True)


@pytest.mark.parametrize("variant", [{}, {"ring_pair": 0}, {"vel_ring": 0}, {"tma": 0}, {"pml_tma": 0}, {"ring_jlen": 7}])
@pytest.mark.parametrize("nz,abc", [(44, "pml"), (45, "pml"), (41, "cerjan")])
def test_single_precision_kernel_variants(tmp_path, variant, nz, abc):
    # float32 fields: vel_ring2 (two cells per thread: odd and even row counts of the interior box, kend_k = 38 / 39 / 41) and
    # every other kernel of the two sweeps against the MP=SP oracle
    o, devs = _run_pair(tmp_path, 24, mp="sp", nz=nz, abc_type=abc, nranks=(2, 1), nx=70, ny=44, options=variant,
                        sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    _compare(o, devs, exact=True)


def test_many_sources_asynchronous_upload(tmp_path):
    # 40 sources in distinct cells: their host-evaluated moment rates reach the device through the pinned ring (17..256 sources;
    # up to 16 ride in the kernel parameters), 70 steps through swpc3d_run so that the 32-slot ring wraps twice
    rng = np.random.default_rng(7)
    src, seen = [], set()
    while len(src) < 40:
        x, y, z = rng.uniform(-8, 8), rng.uniform(-6, 6), rng.uniform(2.5, 12.0)
        key = (int(x / 0.5 + 100), int(y / 0.5 + 100), int(z / 0.5))
        if any((key[0] + a, key[1] + b, key[2] + c) in seen for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)):
            continue                                  # the 4-node shear stencils must not overlap (atomics: free summation order)
        seen.add(key)
        src.append(f"{x:.3f} {y:.3f} {z:.3f} {rng.uniform(0, 0.3):.3f} {rng.uniform(0.3, 0.8):.3f} 1e15 " + " ".join(f"{v:.3f}" for v in rng.uniform(-1, 1, 6)))
    inf = write_case(tmp_path, nt=70, sources=src)
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    d = device_from_oracle(o, 0, device=0)
    o.run(1, 70)
    d.run(1, 70)
    d.sync()
    _compare(o, [d], exact=True)


def test_step_entry_point_and_run(tmp_path):
    inf = write_case(tmp_path, nt=24, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    d = device_from_oracle(o, 0, device=0)
    o.run(1, 24)
    d.run(1, 24)
    d.sync()
    _compare(o, [d], exact=True)


@pytest.mark.parametrize("variant", [{"tma": 0, "vel_ring": 0}, {"pml_tma": 0}, {"tma_persist": 1}, {"tma_persist": 1, "tma_pl": 5}, {"bottom_tma": 1}, {"pml_jl": 4, "pml_jl_bottom": 5}, {"side_streams": 0}, {"tma": 1, "vel_ring": 0}, {"tma": 2, "vel_ring": 0}, {"vel_ring": 1, "ring_jlen": 5, "ring_pf": 0}, {"vel_ring": 1, "ring_pf": 2}, {"flat_bottom": 0},
                                     {"tma_shift": 1}, {"tma_shift": 0}])
@pytest.mark.parametrize("abc", ["pml", "cerjan"])
def test_kernel_variants_bit_exact(tmp_path, variant, abc):
    # every kernel variant of the two sweeps (direct, TMA-staged stress, TMA-staged stress + velocity, register-ring velocity) shares one arithmetic body
    o, devs = _run_pair(tmp_path, 24, nranks=(2, 1), nx=70, ny=44, abc_type=abc, options=variant,
                        sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    _compare(o, devs, exact=True)


@pytest.mark.parametrize("case", [
    dict(nx=48, ny=40, nz=44, na=6),                                   # bottom box starts 2 rows above the absorber (k = 39 is 3 mod 4)
    dict(nx=52, ny=47, nz=70, na=9),                                   # three k-tiles of the walls, the last one partial; 9-column walls = 2 tiles
    dict(nx=90, ny=60, nz=44, na=20),                                  # the bench's absorber: 20 columns = two 10-wide tiles, 20 rows from k = 25
    dict(nx=90, ny=60, nz=45, na=20, no_bottom=True),                  # 20 rows from k = 26 do not fit the 20-row box (it would start at 25)
    dict(nx=44, ny=40, nz=37, na=3),                                   # walls thinner than the pipeline is deep: those regions stay with sweep_direct
    dict(nx=60, ny=52, nz=64, na=10, nranks=(2, 2)),                   # ranks with walls on two sides only
    dict(nx=48, ny=40, nz=44, na=6, mp="sp"),                          # float32 fields
    dict(nx=48, ny=40, nz=44, na=6, nm=0, vmodel="lhm_land"),          # elastic
])
def test_tma_staged_absorber_shell_bit_exact(tmp_path, case):
    """pml_tma (persistent, TMA-staged PML cells) against the oracle and against sweep_direct's PML path, every region kind:
    full / partial k-tiles and column tiles, bottom boxes that start above the absorber, regions too thin for the pipeline."""
    case = dict(case)
    nranks, mp, nm, no_bottom = case.pop("nranks", (1, 1)), case.pop("mp", "dp"), case.pop("nm", 3), case.pop("no_bottom", False)
    src = ["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"]
    o, devs = _run_pair(tmp_path / "a", 36, nranks=nranks, mp=mp, nm=nm, sources=src, **case)
    _compare(o, devs, exact=True)
    thin = case["na"] < 4
    for d in devs:   # the staged kernels did run (or, for the thin walls, did not)
        walls, bottom = d.info("pml_items_walls"), d.info("pml_items_bottom")
        assert (bottom > 0) != no_bottom
        assert walls > 0
        assert d.info("pml_direct_boxes") >= (2 if thin else 0)   # j slabs of 3 planes are shorter than the pipeline: sweep_direct
        assert d.info("pml_items_walls_vel") == walls
    o2, devs2 = _run_pair(tmp_path / "b", 36, nranks=nranks, mp=mp, nm=nm, sources=src, options={"pml_tma": 0}, **case)
    _compare(o2, devs2, exact=True)
    assert all(d.info("pml_items_walls") == 0 and d.info("pml_items_bottom") == 0 for d in devs2)
    # the absorber has been reached: the ADE variables are in use
    r = o.rank(0)
    assert np.abs(o.field(0, "Vz")[3:3 + r["nyp"], 3:3 + case["na"] + 1, 3:-3]).max() > 0 or nranks != (1, 1)


@pytest.mark.parametrize("case", [
    dict(nx=48, ny=44, nz=64, na=20),                                  # the bench's shape: 12 interior rows + 20 PML rows per tile
    dict(nx=52, ny=47, nz=96, na=10),                                  # 22 interior rows, partial tile columns
    dict(nx=48, ny=40, nz=64, na=6),                                   # 26 interior rows: every consumer warp has a role
    dict(nx=48, ny=44, nz=64, na=7),                                   # na not a multiple of 4: R / aux boxes start off the role boundary
    dict(nx=64, ny=56, nz=64, na=10, nranks=(2, 2)),                   # decomposed: core and slab regions each get their plan
    dict(nx=48, ny=44, nz=64, na=20, mp="sp"),                         # float32 fields (vel_ring2 above the tiles)
    dict(nx=48, ny=44, nz=64, na=20, nm=0, vmodel="lhm_land"),         # elastic: no memory variables, no R box
])
def test_whole_line_bottom_tiles_bit_exact(tmp_path, case):
    """bottom_tma (option, off by default: measured no faster): the rows k = nz-31 .. nz as whole 32-row tiles whose interior cells
    and PML cells are updated by different warps of one block; stress_tma / vel_ring stop at k = nz-32.  Against the oracle, and
    against the same run with the tiles off."""
    case = dict(case)
    nranks, mp, nm = case.pop("nranks", (1, 1)), case.pop("mp", "dp"), case.pop("nm", 3)
    zdeep = -3.0 + (case["nz"] - case["na"] - 4) * 0.5      # the second source sits four cells above the absorber
    src = ["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8", f"-2.1 1.7 {zdeep} 0.10 0.5 4e14 -0.2 0.9 0.1 -0.5 0.3 0.6"]
    case.setdefault("zbeg", -3.0)
    o, devs = _run_pair(tmp_path / "a", 40, nranks=nranks, mp=mp, nm=nm, sources=src, options={"bottom_tma": 1}, **case)
    _compare(o, devs, exact=True)
    for d in devs:
        assert d.info("bottom_items") > 0 and d.info("bottom_items_vel") > 0 and d.info("pml_items_bottom") == 0
    o2, devs2 = _run_pair(tmp_path / "b", 40, nranks=nranks, mp=mp, nm=nm, sources=src, **case)    # the default: tiles off
    _compare(o2, devs2, exact=True)
    assert all(d.info("bottom_items") == 0 for d in devs2)
    assert np.abs(o.field(0, "Szz")[3:-3, 3:-3, -8:-3]).max() > 0 or nranks != (1, 1)   # the bottom rows have seen the wave


@pytest.mark.parametrize("opts", [{}, {"bottom_tma": 1}, {"pml_tma": 0}])
def test_unpacked_ade_columns_bit_exact(tmp_path, monkeypatch, opts):
    """SWPC3D_AUX_PACK=0: the round-1 layout of the ADE arrays (every column of the bottom rows on its own 128-byte line) that the
    packed default is measured against; na = 7 puts the columns off the 16-byte grid.  A source just above the absorber."""
    monkeypatch.setenv("SWPC3D_AUX_PACK", "0")
    case = dict(nx=48, ny=44, nz=64, na=7, zbeg=-3.0)
    src = ["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8", f"-2.1 1.7 {-3.0 + (64 - 7 - 4) * 0.5} 0.10 0.5 4e14 -0.2 0.9 0.1 -0.5 0.3 0.6"]
    o, devs = _run_pair(tmp_path / "a", 40, sources=src, options=opts, **case)
    _compare(o, devs, exact=True)
    assert np.abs(o.field(0, "Szz")[3:-3, 3:-3, -8:-3]).max() > 0


@pytest.mark.parametrize("opts", [{}, {"slab_tiled": 0}, {"slab_x": 2}, {"slab_x": 5}])
@pytest.mark.parametrize("abc,bf", [("pml", False), ("cerjan", False), ("pml", True)])
def test_boundary_first_split_bit_exact(tmp_path, abc, bf, opts):
    # swpc3d_step's boundary-first schedule (boundary slabs -> exchange stream | core sweep -> join) with the sweeps and
    # the source terms split as if all four faces had neighbours; sources sit on / next to the slab-core seam
    if bf:
        src = ["-11.3 -9.3 4.1 0.05 0.6 1e12 2e12 -3e12", "-10.7 0.2 5.1 0.05 0.6 1e12 2e12 -3e12", "11.4 8.9 4.6 0.0 0.5 -1e12 1e12 2e12"]
    else:
        src = ["-11.3 -9.3 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8", "-10.7 0.2 5.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8",
               "11.4 8.9 4.6 0.0 0.5 2e15 -0.7 0.3 0.5 -0.4 0.6 0.8", "0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"]
    inf = write_case(tmp_path, nt=24, abc_type=abc, bf_mode=bf, sources=src)
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    d = device_from_oracle(o, 0, device=0)
    d.set_option("split_test", 1)
    for key, val in opts.items():   # x slabs one tile column wide swept by the tiled kernels (default), 2 / 5 columns, or sweep_direct
        d.set_option(key, val)
    o.run(1, 24)
    d.run(1, 24)
    d.sync()
    assert np.abs(o.field(0, "Vz")).max() > 0
    _compare(o, [d], exact=True)


@pytest.mark.parametrize("abc", ["pml", "cerjan"])
def test_ranks_entirely_inside_the_absorber(tmp_path, abc):
    # 4 x 1 ranks of 10 columns with na = 10: the first and the last rank own absorber cells only (empty interior kernel box,
    # m_global.f90:355-376), the middle ones interior columns only
    o, devs = _run_pair(tmp_path, 30, nranks=(4, 1), nx=40, ny=36, nz=40, na=10, abc_type=abc,
                        sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    r0 = o.rank(0)
    if abc == "pml":
        assert r0["iend_k"] < r0["ibeg_k"]
    assert all(np.abs(o.field(q, "Vz")).max() > 0 for q in range(4))
    _compare(o, devs, exact=True)


def test_odd_sizes_and_padding(tmp_path):
    # sizes that are multiples of nothing, and a source one cell from the PML
    o, devs = _run_pair(tmp_path, 30, nranks=(1, 1), nx=37, ny=29, nz=35, na=5, sources=["-5.9 3.7 3.3 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    _compare(o, devs, exact=True)


def test_create_gives_memory_back_when_it_fails_half_way():
    """A rank that does not fit: the field block (72 B/cell) still fits, the memory variables do not.  swpc3d_create must report
    the CUDA error and release what it had taken, so that a smaller rank can be created afterwards."""
    import torch

    from openswpc_b200._lib import Swpc3dError
    from openswpc_b200.device import DeviceRank, RankGeometry

    free0, total = torch.cuda.mem_get_info(0)
    ncell_target = int(free0 * 0.62 / 72)                      # fields fit (62 % of what is free), fields + memory variables do not
    nz = 1000
    n = int((ncell_target / (nz + 56)) ** 0.5)
    geom = RankGeometry(nx=n, ny=n, nz=nz, nproc_x=1, nproc_y=1, myid=0, ibeg=1, iend=n, jbeg=1, jend=n, ibeg_k=21, iend_k=n - 20, jbeg_k=21,
                        jend_k=n - 20, kbeg_k=1, kend_k=nz - 20, na=20)
    ts = np.array([7.9577475, 0.79577476, 0.07957747], dtype=np.float32)
    with pytest.raises(Swpc3dError, match="out of memory"):
        DeviceRank(geom, dx=0.5, dy=0.5, dz=0.5, dt=0.02, nm=3, abc_type="pml", ts=ts, device=0)
    free1, _ = torch.cuda.mem_get_info(0)
    assert free1 >= free0 - (256 << 20), (free0, free1)
    small = RankGeometry(nx=64, ny=64, nz=64, nproc_x=1, nproc_y=1, myid=0, ibeg=1, iend=64, jbeg=1, jend=64, ibeg_k=11, iend_k=54, jbeg_k=11,
                         jend_k=54, kbeg_k=1, kend_k=54, na=10)
    d = DeviceRank(small, dx=0.5, dy=0.5, dz=0.5, dt=0.02, nm=3, abc_type="pml", ts=ts, device=0)
    d.close()


@pytest.mark.parametrize("model", ["grd", "grd_flat_land", "lhm_rmed"])
@pytest.mark.parametrize("nranks", [(1, 1), (2, 2)])
def test_laterally_heterogeneous_models_bit_exact(tmp_path, model, nranks):
    """Topography / bathymetry that changes from column to column (vmodel_grd: kfs, kob and the 2nd-order bands kfs_top ..
    kob_bot differ per column, land and sea columns side by side) and random-media perturbations of every cell (lhm_rmed):
    the sweeps, the PML shell and the free-surface / ocean-bottom band logic on media the 1-D models of the other tests
    cannot produce.  Sources sit under land and under sea; a long enough run for the waves to reach the surface."""
    from helpers import write_grd, write_rmed

    nt = 60
    if model.startswith("grd"):
        lon = 139.40 + 0.01 * np.arange(72)
        lat = 35.50 + 0.01 * np.arange(46)
        LO, LA = np.meshgrid(lon, lat)
        write_grd(tmp_path / "g1.grd", lon, lat, 700.0 * np.sin((LO - 139.76) * 60.0) * np.cos((LA - 35.72) * 55.0) + 100.0)   # +-0.7 km of relief
        write_grd(tmp_path / "g2.grd", lon, lat, 3200.0 + 900.0 * np.cos((LO - 139.7) * 25.0) + 400.0 * np.sin((LA - 35.7) * 30.0))
        write_grd(tmp_path / "g3.grd", lon, lat, 9500.0 + 1500.0 * np.sin((LO - 139.8) * 12.0 + (LA - 35.7) * 9.0))
        (tmp_path / "grd.lst").write_text("'g1.grd' 2.1 3.0 1.6 100 50 0\n'g2.grd'  2.5 5.0 2.9 300 150 0\n g3.grd  2.9 6.8 3.9 500 250 1\n")
        vm = "vmodel_type = 'grd'\n fn_grdlst = 'grd.lst'\n dir_grd = '.'\n" + (" is_ocean = .false.\n" if model == "grd_flat_land" else "")
    else:
        rng = np.random.default_rng(3)
        write_rmed(tmp_path / "r1.nc", (0.05 * rng.standard_normal((20, 12, 16))).astype(np.float32))
        write_rmed(tmp_path / "r2.nc", (0.08 * rng.standard_normal((30, 25, 35))).astype(np.float32))
        (tmp_path / "layers_rmed.dat").write_text("# depth rho vp vs Qp Qs rmed\n0.0 2.3 5.5 3.14 600 300 r1.nc\n3.0 2.4 6.0 3.55 400 200 r2.nc\n"
                                                  "9.0 2.8 6.7 3.83 600 300 r1.nc\n15.0 3.2 7.8 4.46 600 300 r2.nc\n")
        vm = "vmodel_type = 'lhm_rmed'\n fn_lhm_rmed = 'layers_rmed.dat'\n dir_rmed = '.'\n"
    o, devs = _run_pair(tmp_path, nt, nranks=nranks, nx=52, ny=44, nz=48, zbeg=-2.0, vmodel="raw:" + vm, dt=0.01,
                        sources=["0.3 -0.2 2.1 0.02 0.3 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8", "-4.1 3.7 1.6 0.05 0.25 4e14 -0.2 0.9 0.1 -0.5 0.3 0.6"])
    _compare(o, devs, exact=True)
    if model == "grd":
        kfs, kob = o.imap(0, "kfs")[4:-4, 4:-4], o.imap(0, "kob")[4:-4, 4:-4]
        assert kob.max() - kob.min() >= 2 and (kob > kfs).any() and (kob == kfs).any()      # relief; sea and land columns
    assert np.abs(o.field(0, "Vz")[3:-3, 3:-3, 3:10]).max() > 0                              # the waves have reached the surface
