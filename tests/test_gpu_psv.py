"""GPU parity of the swpc_psv path: the CUDA kernels (through the C ABI of include/swpcpsv_b200.h) against the CPU oracle
(oracle/psv.c) on the same inputs.  The library is built with -fmad=false and keeps the reference's kinds and evaluation
order, so fields, memory variables and station traces are required to be BIT-IDENTICAL (the north-star bar is 1e-5 relative
L2 for the seismograms)."""
import numpy as np
import pytest

from psv_oracle import FIELDS, MAPS, MEDIUM, PsvOracle, psv_case_text, psv_device_from_oracle, write_psv_files

pytestmark = pytest.mark.gpu


def _pair(tmp_path, nt, *, nm=3, sp=False, nproc_x=1, sources=None, stations=None, hetero=False, **kw):
    write_psv_files(tmp_path, sources=sources, stations=stations)
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=nt, **kw))
    o = PsvOracle(inf, base_dir=tmp_path, nm=nm, nproc_x=nproc_x, sp=sp)
    if hetero:
        _make_heterogeneous(o)
    devs = [psv_device_from_oracle(o, q, device=0) for q in range(o.nranks)]
    return o, devs


def _make_heterogeneous(o, seed=20251017):
    """Replace the layered medium by a random one with topography, an ocean basin and velocity perturbations (the same global
    model cut per rank), then re-run surface_detection: exercises kfs != kob, the 2nd-order bands and the harmonic mu."""
    nx, nz = o.cfg("nx"), o.cfg("nz")
    rng = np.random.default_rng(seed)
    i = np.arange(-2, nx + 4)                      # global columns incl. margins
    k = np.arange(-2, nz + 4)
    ksurf = 9 + np.round(2.5 * np.sin(2 * np.pi * i / 37.0) + 1.5 * np.cos(2 * np.pi * i / 11.0)).astype(int)   # free surface (air above)
    ksea = np.where((i > 0.35 * nx) & (i < 0.6 * nx), ksurf + 4, ksurf)                                          # sea floor
    xi = rng.normal(0, 0.04, size=(i.size, k.size))
    vs = 2.9 * (1 + xi) * (1 + 0.004 * k[None, :])
    vp = 5.0 * (1 + xi) * (1 + 0.004 * k[None, :])
    rho = 2.6 * (1 + 0.8 * xi)
    air = k[None, :] <= ksurf[:, None]
    sea = (~air) & (k[None, :] <= ksea[:, None])
    rho = np.where(air, 0.001, np.where(sea, 1.0, rho)).astype(np.float32)
    mu = np.where(air | sea, 0.0, rho * vs * vs).astype(np.float32)
    lam = np.where(air, 0.0, np.where(sea, 1.0 * 1.5 * 1.5, rho * (vp * vp - 2 * vs * vs))).astype(np.float32)
    taup = np.where(air, 0.05, 0.004 * (1 + rng.random(size=rho.shape))).astype(np.float32)
    taus = np.where(air, 0.05, 0.008 * (1 + rng.random(size=rho.shape))).astype(np.float32)
    for q in range(o.nranks):
        r = o.rank(q)
        sl = slice(r["ibeg_m"] + 2, r["ibeg_m"] + 2 + r["nxm"])
        for n, a in (("rho", rho), ("lam", lam), ("mu", mu), ("taup", taup), ("taus", taus)):
            o.set_field(q, n, a[sl])
    o.redetect_surface()


def _step_all(o, devs, nt):
    from openswpc_b200.psv_device import step_local

    for it in range(1, nt + 1):
        o.step(it)
        step_local(devs, it)


def _compare(o, devs, products=(0,)):
    nz = o.cfg("nz")
    for q, d in enumerate(devs):
        got = d.download_fields()
        r = o.rank(q)
        nxo = r["iend"] - r["ibeg"] + 1
        # owned cells and the halo columns filled by the exchange (ibeg-2..iend+2), k = 1..nz
        sl = (slice(1, 5 + nxo), slice(3, 3 + nz))
        for n in FIELDS:
            ref, a = o.field(q, n), got[n].astype(np.float64)
            own = (slice(3, 3 + nxo), slice(3, 3 + nz))
            assert np.array_equal(a[own], ref[own]), f"rank {q} {n}: max abs diff {np.abs(a[own] - ref[own]).max():.3e} of {np.abs(ref[own]).max():.3e}"
            if n in ("Sxx", "Sxz", "Vx", "Vz"):
                assert np.array_equal(a[sl], ref[sl]), f"rank {q} {n}: halo columns differ"
        if o.cfg("nm") > 0:
            mv = d.download_memvars()
            for n in ("Rxx", "Rzz", "Rxz"):
                assert np.array_equal(mv[n][3:3 + nxo, 3:3 + nz], o.memvar(q, n)[3:3 + nxo, 3:3 + nz]), f"rank {q} {n}"
        for prod in products if d.nst else ():
            assert np.array_equal(d.get_wav(prod), o.wav(q, prod)), f"rank {q} station product {prod}"


@pytest.mark.parametrize("opts", [{}, {"ilen": 7, "pf": 0}])
def test_pml_nm3_bit_exact(tmp_path, opts):
    o, devs = _pair(tmp_path, 80, products="v,u,stress,strain")
    for key, val in opts.items():
        devs[0].set_option(key, val)
    _step_all(o, devs, 80)
    assert np.abs(o.field(0, "Vz")).max() > 0
    _compare(o, devs, products=(0, 1, 2, 3))
    assert np.array_equal(devs[0].vmax() * np.float32(o.cfg("UC")) * np.float32(o.cfg("M0")), o.vmax())


def test_cerjan_nm0_bit_exact(tmp_path):
    o, devs = _pair(tmp_path, 60, nm=0, abc="cerjan")
    _step_all(o, devs, 60)
    _compare(o, devs)


def test_float32_fields_bit_exact(tmp_path):
    o, devs = _pair(tmp_path, 60, sp=True)
    _step_all(o, devs, 60)
    _compare(o, devs)


@pytest.mark.parametrize("abc", ["pml", "cerjan"])
def test_three_ranks_uneven_bit_exact(tmp_path, abc):
    o, devs = _pair(tmp_path, 100, nx=100, nproc_x=3, abc=abc, sources=["10.2 0.0 4.2 0.05 0.6 1e15 0.7 0.0 -0.3 0.0 0.5 0.0",
                                                                          "-8.4 0.0 7.0 0.2 0.8 5e14 0.1 0.0 0.9 0.0 -0.4 0.0"])
    _step_all(o, devs, 100)
    assert all(np.abs(o.field(q, "Vx")).max() > 0 for q in range(3))
    _compare(o, devs)


@pytest.mark.parametrize("abc", ["pml", "cerjan"])
def test_body_force_in_absorber_order(tmp_path, abc):
    # a force inside the absorber: interior sweep -> force -> absorber sweep must keep the reference's summation order
    o, devs = _pair(tmp_path, 60, bf_mode=True, abc=abc, sources=["-21.3 0.0 4.2 0.05 0.6 3e9 0.0 -4e9", "1.3 0.0 2.2 0.05 0.5 1e9 0.0 2e9"])
    _step_all(o, devs, 60)
    assert np.abs(o.field(0, "Vz")).max() > 0
    _compare(o, devs)


def test_heterogeneous_ocean_two_ranks_bit_exact(tmp_path):
    o, devs = _pair(tmp_path, 120, nx=128, nz=96, nproc_x=2, hetero=True, stftype="herrmann",
                    sources=["0.3 0.0 9.2 0.05 0.6 1e15 0.7 0.0 -0.3 0.0 0.5 0.0"], stations=["-6.1 0.0 0.0 st01 obb", "5.3 0.0 3.0 st02 oba",
                                                                                             "11.2 0.0 0.0 st03 fsb"], products="v,strain")
    kfs, kob = o.map(0, "kfs"), o.map(0, "kob")
    assert (kob[3:-3] != kfs[3:-3]).any()
    _step_all(o, devs, 120)
    _compare(o, devs, products=(0, 3))


def test_run_entry_point_and_launch_count(tmp_path):
    o, devs = _pair(tmp_path, 30)
    d = devs[0]
    n0 = d.info("launches")
    d.run(1, 30)
    d.sync()
    o.run(1, 30)
    _compare(o, devs)
    # per step: stress sweep + glut + 2 halo unpacks (outer sides) + vel sweep + 2 halo unpacks, + wav_store on sampled steps
    assert d.info("launches") - n0 >= 30 * 3


@pytest.mark.parametrize("wav_format", ["sac", "tar_st"])
def test_host_driver_end_to_end_files_identical(tmp_path, wav_format):
    """input.inf -> C++ setup chain -> GPU loop -> waveform files, against the oracle's files byte for byte."""
    from openswpc_b200.swpc_psv import SwpcPsv

    write_psv_files(tmp_path)
    (tmp_path / "lhm.dat").write_text("# depth rho vp vs qp qs\n0.0 2.3 5.5 3.14 600 300\n3.0 2.4 6.0 3.55 600 300\n16.0 2.8 6.7 3.83 600 300\n")
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=100, vmodel="lhm", products="v,u,stress,strain", extra=f" fn_lhm = 'lhm.dat'\n wav_format = '{wav_format}'"))
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    o.set_exedate(1700000000, 540)
    vm_ref = o.run(1, 100)
    run = SwpcPsv(inf, base_dir=tmp_path, nm=3)
    run.set_exedate(1700000000, 540)
    run.attach_device(0)
    vm = run.run(1, 100)
    assert np.array_equal(vm, vm_ref) and vm.max() > 0
    n = run.write_wav(tmp_path / "gpu")
    assert n == 4 * 10
    for prod in range(4):
        assert np.array_equal(run.wav(prod), o.wav(0, prod)), prod
    if wav_format == "sac":
        assert o.write_sac(tmp_path / "ref") == n
        for f in sorted((tmp_path / "ref" / "wav").iterdir()):
            assert (tmp_path / "gpu" / "wav" / f.name).read_bytes() == f.read_bytes(), f.name
    else:
        import tarfile

        with tarfile.open(tmp_path / "gpu" / "wav" / "psvtest.psv.st02.sac.tar") as t:
            names = t.getnames()
            assert names[0] == "psvtest.psv.st02.Vx.sac" and len(names) == 10     # m_wav.f90:428: title.psv.stnm.cmp.sac
            body = t.extractfile(names[1]).read()
        assert np.array_equal(np.frombuffer(body[632:], dtype="<f4"), o.wav(0, 0)[1, 1])


PSV_SNAP_TAGS = ("ps", "v", "u")
PSV_SNAP_VARS = (("divergence", "rotation"), ("Vx", "Vz"), ("Ux", "Uz"))


@pytest.mark.parametrize("dec", [(2, 2, 5), (1, 1, 4), (3, 4, 7)])
def test_snapshot_files_netcdf(tmp_path, dec):
    """m_snap.f90 of swpc_psv: the three xz products written by the product's driver from the device slices, read back with an
    independent netCDF reader and compared record by record, bit for bit, with the oracle."""
    from scipy.io import netcdf_file

    from openswpc_b200.swpc_psv import SwpcPsv

    nt = 43
    write_psv_files(tmp_path)
    inf = tmp_path / "input.inf"
    extra = f" snp_format = 'netcdf'\n xz_ps%sw = .true.\n xz_v%sw = .true.\n xz_u%sw = .true.\n idec = {dec[0]}\n kdec = {dec[1]}\n ntdec_s = {dec[2]}"
    inf.write_text(psv_case_text(nt=nt, extra=extra))
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    o.run(1, nt)
    run = SwpcPsv(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    run.snap_open(tmp_path / "gpu")
    run.run(1, nt)
    run.snap_close()
    info = o.snap_info()
    x, z = o.snap_coords()
    title = o.cfg("title")
    for p in range(3):
        path = tmp_path / "gpu" / f"{title}.psv.xz.{PSV_SNAP_TAGS[p]}.nc"
        assert path.exists(), path.name
        recs, its = o.snap_records(p)
        with netcdf_file(str(path), "r", mmap=False) as f:
            assert f.dimensions == {"x": info["nxs"], "z": info["nzs"], "t": None}
            assert f.generated_by == b"SWPC" and f.codetype == b"SWPC_PSV" and f.hdrver == 6 and f.coordinate == b"xz"
            assert f.datatype == [b"ps", b"v2", b"u2"][p] and f.nsnp == 2 and f.nmed == 3 and f.ns1 == info["nxs"] and f.ns2 == info["nzs"]
            np.testing.assert_array_equal(f.variables["x"][:], x)
            np.testing.assert_array_equal(f.variables["z"][:], z)
            for m, name in enumerate(["rho", "lambda", "mu"]):
                np.testing.assert_array_equal(f.variables[name][:], o.snap_medium(m), err_msg=name)
            assert len(its) == f.variables["t"].shape[0] > 1
            np.testing.assert_array_equal(f.variables["t"][:], np.array([np.float32(it) * np.float32(run["dt"]) for it in its], dtype=np.float32))
            for v, name in enumerate(PSV_SNAP_VARS[p]):
                var = f.variables[name]
                assert var.dimensions == ("t", "z", "x")
                np.testing.assert_array_equal(var[:], recs[:, v], err_msg=f"{path.name}:{name}")
                np.testing.assert_array_equal(var.actual_range, [min(recs[:, v].min(), 0), max(recs[:, v].max(), 0)])
        assert np.abs(recs).max() > 0, path.name


def test_snapshot_native_stream(tmp_path):
    """snp_format = 'native': write_snp_header (m_snap.f90:272-315), the three medium slices, then two arrays per output step."""
    import struct

    from openswpc_b200.swpc_psv import SwpcPsv

    nt = 21
    write_psv_files(tmp_path)
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=nt, extra=" xz_v%sw = .true.\n idec = 2\n kdec = 2\n ntdec_s = 5"))
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    o.run(1, nt)
    run = SwpcPsv(inf, base_dir=tmp_path, nm=3)
    run.set_exedate(1_700_000_000, 540)
    run.attach_device(0)
    run.snap_open(tmp_path / "gpu")
    run.run(1, nt)
    run.snap_close()
    info = o.snap_info()
    x, z = o.snap_coords()
    recs, its = o.snap_records(1)
    raw = (tmp_path / "gpu" / f"{o.cfg('title')}.psv.xz.v.snp").read_bytes()
    title = o.cfg("title").ljust(80).encode()
    hdr = (b"STREAMIO" + b"SWPC_PSV" + struct.pack("<i", 6) + title + struct.pack("<i", 1_700_000_000) + b"xz" + b"v2" +
           struct.pack("<ii", info["nxs"], info["nzs"]) + struct.pack("<ffff", x[0], z[0], x[1] - x[0], z[1] - z[0]) +
           struct.pack("<f", np.float32(run["dt"]) * np.float32(5)) + struct.pack("<iiii", o.cfg("na") // 2, o.cfg("na") // 2, 3, 2) +
           struct.pack("<ffffff", np.float32(139.7604), np.float32(35.7182), 0.0, 0.0, 0.0, 0.0))   # clon clat phi defaults (m_global.f90:140-142), 3 x dum
    assert raw[:len(hdr)] == hdr
    body = np.frombuffer(raw[len(hdr):], dtype=np.float32)
    np2 = info["nxs"] * info["nzs"]
    assert body.size == np2 * (3 + 2 * len(its))
    for m in range(3):
        np.testing.assert_array_equal(body[m * np2:(m + 1) * np2].reshape(info["nzs"], info["nxs"]), o.snap_medium(m))
    np.testing.assert_array_equal(body[3 * np2:].reshape(len(its), 2, info["nzs"], info["nxs"]), recs)


@pytest.mark.parametrize("ps,abc", [("p", "pml"), ("s", "pml"), ("s", "cerjan")])
def test_planewave_mode_bit_exact(tmp_path, ps, abc):
    """pw_mode through the product's driver: initial condition on the host, edge extrapolation on the device ahead of the
    PML updates (m_absorb_p.f90:114-155, :287-326), against the oracle."""
    from openswpc_b200.swpc_psv import SwpcPsv

    nt = 60
    write_psv_files(tmp_path)
    inf = tmp_path / "input.inf"
    inf.write_text(psv_case_text(nt=nt, abc=abc, extra=f" pw_mode = .true.\n pw_ztop = 14.0\n pw_zlen = 6.0\n pw_ps = '{ps}'\n pw_dip = 20.0"))
    o = PsvOracle(inf, base_dir=tmp_path, nm=3)
    vm_ref = o.run(1, nt)
    run = SwpcPsv(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    vm = run.run(1, nt)
    assert np.array_equal(vm, vm_ref) and np.abs(vm_ref).max() > 0
    got = run.download_fields()
    nz, nxo = o.cfg("nz"), o.rank(0)["iend"] - o.rank(0)["ibeg"] + 1
    for n, a in got.items():
        ref = o.field(0, n)
        assert np.array_equal(a[3:3 + nxo, 3:3 + nz], ref[3:3 + nxo, 3:3 + nz]), n
    run.write_wav(tmp_path / "out")
    assert np.array_equal(run.wav(0), o.wav(0, 0)) and np.abs(o.wav(0, 0)).max() > 0
