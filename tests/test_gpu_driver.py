"""End-to-end on the GPU through the product's own driver (input.inf -> host setup -> device -> SAC files), against
the oracle run of the same input: progress-line amplitudes, station traces and SAC files byte for byte."""
import numpy as np
import pytest

from helpers import rel_l2, write_case
from openswpc_b200.swpc3d import Swpc3d
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [dict(), dict(abc_type="cerjan", vmodel="uni"), dict(benchmark=True, nx=64, ny=64, nz=80, na=20)])
def test_driver_sac_matches_oracle(tmp_path, case):
    nt = 60
    inf = write_case(tmp_path, nt=nt, ntdec_r=10, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"], **case)
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.lib.ora_set_exedate(o.h, 1_700_000_000, 540)
    vm_ref = o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.set_exedate(1_700_000_000, 540)
    run.attach_device(0)
    vm = run.run(1, nt)
    np.testing.assert_array_equal(vm, vm_ref)                      # report__progress triplets
    n_ref = o.write_sac(tmp_path / "ref")
    n = run.write_sac(tmp_path / "gpu")
    assert n == n_ref == 3 * run["nst"]
    assert n > 0 or case.get("benchmark")
    for f in sorted((tmp_path / "ref" / "wav").glob("*.sac")):
        g = tmp_path / "gpu" / "wav" / f.name
        assert g.exists(), f.name
        assert g.read_bytes() == f.read_bytes(), f.name
    if run["nst"]:
        w = run.wav()
        assert rel_l2(w, o.wav(0)) <= 1e-5                         # the north-star bar; in fact identical
        np.testing.assert_array_equal(w, o.wav(0))


def test_source_on_the_model_edge(tmp_path):
    """A stress-glut stencil at i = 1, j = ny reaches the outer halo planes; the reference zeroes them again at every exchange
    (unconditional unpack of the never-received buffers, m_global.f90:458-488): the whole memory box must agree."""
    nt = 12
    inf = write_case(tmp_path, nt=nt, sources=["-11.9 9.9 4.1 0.0 0.3 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    assert run.array("src_ijk").tolist() == [[1, 40, 15]]
    run.attach_device(0)
    run.run(1, nt)
    got, nz = run.download_fields(), run["nz"]
    for f, a in got.items():
        np.testing.assert_array_equal(a[:, :, 3:3 + nz], o.field(0, f)[:, :, 3:3 + nz], err_msg=f)
    assert np.abs(got["Sxy"]).max() > 0


def test_sac_header_fields(tmp_path):
    inf = write_case(tmp_path, nt=20, title="hdrtest")
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.set_exedate(86400 * 365, 0)
    run.attach_device(0)
    run.run(1, 20)
    run.write_sac(tmp_path / "o")
    raw = (tmp_path / "o" / "wav" / "hdrtest.3d.st01.Vz.sac").read_bytes()
    f = np.frombuffer(raw[:280], dtype=np.float32)
    i = np.frombuffer(raw[280:420], dtype=np.int32)
    assert len(raw) == 632 + 4 * run["ntw"]
    assert f[0] == np.float32(int(np.float64(np.float32(2 * np.float32(0.02))) * 1e7)) / np.float32(1e7)   # delta, m_sac.f90:339
    assert i[9] == run["ntw"] and i[6] == 6 and i[15] == 1 and i[16] == 7
    assert i[0] == 1971 and i[1] == 1
    assert raw[440:448] == b"st01    " and raw[448:464] == b"hdrtest         " and raw[600:608] == b"Vz      "


def test_all_station_products_sac(tmp_path):
    """sw_wav_u / sw_wav_stress / sw_wav_strain (m_wav.f90:430-615): 18 SAC files per station, byte-identical."""
    nt = 40
    inf = write_case(tmp_path, nt=nt, extra="sw_wav_u = .true.\n sw_wav_stress = .true.\n sw_wav_strain = .true.")
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.lib.ora_set_exedate(o.h, 1_700_000_000, 540)
    o.run(1, nt)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.set_exedate(1_700_000_000, 540)
    run.attach_device(0)
    run.run(1, nt)
    assert o.write_sac(tmp_path / "ref") == run.write_sac(tmp_path / "gpu") == 18 * run["nst"]
    names = sorted(f.name for f in (tmp_path / "ref" / "wav").glob("*.sac"))
    assert any(".Uz." in n for n in names) and any(".Sxy." in n for n in names) and any(".Eyz." in n for n in names)
    for n in names:
        assert (tmp_path / "gpu" / "wav" / n).read_bytes() == (tmp_path / "ref" / "wav" / n).read_bytes(), n
    assert np.abs(run.array("wav_strain")).max() > 0 and np.abs(run.array("wav_u")).max() > 0


@pytest.mark.parametrize("fmt", ["csf", "tar_st", "tar_node"])
def test_waveform_containers(tmp_path, fmt):
    """wav_format = csf / tar_st / tar_node (m_wav.f90:707-761): the containers hold the very SAC records of the sac format."""
    import io
    import struct
    import tarfile

    nt = 24
    kw = dict(nt=nt, title="box", extra="sw_wav_u = .true.\n sw_wav_strain = .true." if fmt != "csf" else "")
    o = Oracle(write_case(tmp_path / "o", **kw), base_dir=tmp_path / "o", nm=3)
    o.lib.ora_set_exedate(o.h, 1_700_000_000, 540)
    o.run(1, nt)
    o.write_sac(tmp_path / "ref")
    run = Swpc3d(write_case(tmp_path / "g", wav_format=fmt, **kw), base_dir=tmp_path / "g", nm=3)
    run.set_exedate(1_700_000_000, 540)
    run.attach_device(0)
    run.run(1, nt)
    nfiles = run.write_sac(tmp_path / "gpu")
    names, nst, ntw = run.station_names(), run["nst"], run["ntw"]
    ref = lambda st, cmp: (tmp_path / "ref" / "wav" / f"box.3d.{st}.{cmp}.sac").read_bytes()
    cmps = ["Vx", "Vy", "Vz"] if fmt == "csf" else ["Vx", "Vy", "Vz", "Ux", "Uy", "Uz", "Exx", "Eyy", "Ezz", "Eyz", "Exz", "Exy"]
    assert nfiles == nst * len(cmps)
    if fmt == "csf":
        raw = (tmp_path / "gpu" / "wav" / "box__00000__.csf").read_bytes()
        assert raw == b"CSFD" + struct.pack("<ii", 3 * nst, ntw) + b"".join(ref(st, c) for st in names for c in cmps)
        return
    files = {"tar_node": ["box.3d.000000.sac.tar"], "tar_st": [f"box.3d.{st}.sac.tar" for st in names]}[fmt]
    assert sorted(p.name for p in (tmp_path / "gpu" / "wav").iterdir()) == sorted(files)
    for fn in files:
        raw = (tmp_path / "gpu" / "wav" / fn).read_bytes()
        sts = names if fmt == "tar_node" else [fn.split(".")[2]]
        want = [(f"box.{st}.3d.{c}.sac", ref(st, c)) for st in sts for c in cmps]
        with tarfile.open(fileobj=io.BytesIO(raw)) as tf:       # an independent reader accepts it (checksums included)
            mem = tf.getmembers()
            assert [m.name for m in mem] == [w[0] for w in want]
            for m, w in zip(mem, want):
                assert tf.extractfile(m).read() == w[1], m.name
                assert m.mtime == 1_700_000_000 and m.mode == 0o644 and m.uname == "root"
        blk = 512 + (632 + 4 * ntw + 511) // 512 * 512 + (512 if (632 + 4 * ntw) % 512 == 0 else 0)
        assert len(raw) == blk * len(want) + 1024 and raw[-1024:] == bytes(1024)
        assert raw[257:263] == b"ustar\0" and raw[263:265] == b"\0\0" and raw[100:108] == b"00000644"   # tar__whdr quirks


def test_station_products_decomposed(tmp_path):
    """2x2 emulated ranks: the strain sampler reads corner halo cells that no exchange fills -- same on both sides."""
    import ctypes as C

    from helpers import device_from_oracle
    from openswpc_b200.device import comm_local

    nt = 30
    inf = write_case(tmp_path, nt=nt, nproc_x=2, nproc_y=2, nx=50, ny=44, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"],
                     stations=["0.4 0.3 0.0 sA obb", "-0.3 -0.4 2.0 sB dep", "0.4 -0.3 0.0 sC fsb", "-0.4 0.4 0.0 sD oba"],
                     extra="sw_wav_u = .true.\n sw_wav_stress = .true.\n sw_wav_strain = .true.")
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    devs = [device_from_oracle(o, q, device=0) for q in range(o.nranks)]
    for d in devs:
        d.set_wav_products(True, True, True, True)
    for it in range(1, nt + 1):
        o.step(it)
        for d in devs:
            d.wav_store(it); d.update_stress(); d.stressglut(it)
        comm_local(devs, "stress")
        for d in devs:
            d.update_vel(); d.bodyforce(it)
        comm_local(devs, "vel")
    o.lib.ora_get_wav_product.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    seen = 0
    for q, d in enumerate(devs):
        if not d.nst:
            continue
        for which in range(4):
            got = d.get_wav_product(which)
            ref = np.zeros_like(got)
            o.lib.ora_get_wav_product(o.h, q, which, ref.ctypes.data_as(C.POINTER(C.c_float)))
            np.testing.assert_array_equal(got, ref, err_msg=f"rank {q} product {which}")
            seen += 1
    assert seen >= 8   # the four stations sit around the 2x2 corner, one per rank


def test_progressive_waveform_output(tmp_path):
    """ntdec_w_prg (m_wav.f90:74, :619-621): the waveform files are (re)written every ntdec_w_prg steps while the run goes on,
    holding the samples taken so far and zeros after them."""
    nt = 20
    inf = write_case(tmp_path, nt=nt, ntdec_w=1, extra="ntdec_w_prg = 7", sources=["0.3 -0.2 4.1 0.0 0.3 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"],
                     stations=["0.5 0.5 3.0 st01 dep", "-1.0 1.5 5.0 st02 dep"])
    inf.write_text(inf.read_text().replace("odir = './out'", f"odir = '{tmp_path}/out'"))   # (odir is relative to the working directory)
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    o.run(1, nt)
    ref = o.wav(0)                                   # (nst, 3, ntw), ntw = nt
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    run.run(1, 12)                                   # the files were written at it = 1 and it = 8; nobody calls write_sac here
    files = sorted((tmp_path / "out" / "wav").glob("*.sac"))
    assert len(files) == 3 * run["nst"] == 6
    f = [p for p in files if "st01" in p.name and p.name.endswith("Vz.sac")][0]
    data = np.frombuffer(f.read_bytes()[632:], dtype="<f4")
    assert data.size == nt
    assert np.array_equal(data[:8], ref[0, 2, :8]) and np.abs(data[:8]).max() > 0
    assert not data[8:].any()


def test_stopwatch_report(tmp_path):
    """stopwatch_mode (main.f90:58, :148-154; m_pwatch.f90:146-195): <odir>/<title>.tim in the reference's table layout, filled
    from the CUDA-event stopwatches of the library; `stopwatch_mode = .false.` writes nothing."""
    import re

    nt = 30
    inf = write_case(tmp_path, nt=nt, title="tim", nx=64, ny=64, nz=64, na=10)
    run = Swpc3d(inf, base_dir=tmp_path, nm=3)
    run.attach_device(0)
    run.run(1, nt)
    run.write_tim(tmp_path / "o")
    lines = (tmp_path / "o" / "tim.tim").read_text().splitlines()
    assert lines[0].startswith("#   CPU     #ID       Procedure Name         Real Time[s]") and lines[1].startswith("# -------+-------+")
    rows = [re.match(r"^   (\d{5})   (\d{5})    (.{22})\s*([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)$", l) for l in lines[2:]]
    assert len(rows) == 4 and all(rows)
    names = [r.group(3).strip() for r in rows]
    assert names == ["kernel__update_stress", "kernel__update_vel", "global__comm", "others"]
    t = [float(r.group(4)) for r in rows]
    assert t[0] > 0 and t[1] > 0 and t[2] == 0.0                     # one rank: no exchange
    assert abs(float(rows[-1].group(5)) - sum(t)) < 2e-3 and abs(float(rows[-1].group(7)) - 100.0) < 0.01
    assert abs(sum(t) - run["loop_seconds"]) < 5e-3
    quiet = write_case(tmp_path / "q", nt=4, title="quiet", extra="stopwatch_mode = .false.")
    r2 = Swpc3d(quiet, base_dir=tmp_path / "q", nm=3)
    r2.attach_device(0)
    r2.run(1, 4)
    r2.write_tim(tmp_path / "o2")
    assert not (tmp_path / "o2" / "quiet.tim").exists()
