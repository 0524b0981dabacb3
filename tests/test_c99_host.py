"""A plain C99 caller of the C ABI (tests/c99_host.c; `gcc -std=c99 -pedantic -Wall -Wextra -Werror`): the headers under
include/ must be valid C as they stand, the structs that cross the boundary must have the layout the ctypes faces (and the
`bind(c)` types of openswpc_b200/fortran/*.f90, which list the same members in the same order) assume, every prototype must
resolve against the shared library -- and, on a GPU box, the call sequence of INTEGRATION.md driven from C, with no Python
near the handle, must reproduce the oracle bit for bit."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build(tmp_path) -> Path:
    from openswpc_b200 import _lib

    _lib.build()
    exe = tmp_path / "c99_host"
    libdir = _lib.LIB_PATH.parent
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c99_host.c"),
                    "-o", str(exe), "-L", str(libdir), "-lswpc3d_b200", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_headers_are_c99_and_struct_layouts_match(tmp_path):
    from openswpc_b200._lib import Grid, PsvGrid

    exe = _build(tmp_path)
    out = subprocess.run([str(exe), "layout"], check=True, capture_output=True, text=True).stdout
    got = dict(re.findall(r"^(\S+(?: \S+)?) (\d+)$", out, flags=re.M))
    assert int(got["entry points"]) >= 50
    assert int(got["sizeof swpc3d_grid"]) == C.sizeof(Grid)
    for name, _ in Grid._fields_:
        assert int(got[f"swpc3d_grid.{name}"]) == getattr(Grid, name).offset, name
    assert int(got["sizeof swpcpsv_grid"]) == C.sizeof(PsvGrid)
    for name in ("nx", "nz", "device", "dx", "dz", "dt"):
        assert int(got[f"swpcpsv_grid.{name}"]) == getattr(PsvGrid, name).offset, name
    # the Fortran binding declares the same members in the same order (bind(c) gives them the C layout)
    f90 = (ROOT / "openswpc_b200" / "fortran" / "m_swpc3d_b200.f90").read_text()
    body = re.search(r"type, bind\(c\)(?:, public)? :: swpc3d_grid(.*?)end type", f90, flags=re.S | re.I).group(1)
    members = [m.strip() for decl in re.findall(r"::\s*([^\n!]+)", body) for m in decl.split(",")]
    assert members == [n for n, _ in Grid._fields_]


@pytest.mark.gpu
def test_c99_host_drives_the_smoke_case_bit_exact(tmp_path):
    from oracle_lib import Oracle

    from helpers import o_source_details, write_case

    nt = 20
    inf = write_case(tmp_path, nt=nt, sources=["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8"])
    o = Oracle(inf, base_dir=tmp_path, nm=3)
    r = o.rank(0)
    from openswpc_b200._lib import Grid

    g = Grid()
    for n in ("nx", "ny", "nz", "nproc_x", "nproc_y", "na"):
        setattr(g, n, int(o.cfg(n)))
    for n in ("ibeg", "iend", "jbeg", "jend", "ibeg_k", "iend_k", "jbeg_k", "jend_k", "kbeg_k", "kend_k"):
        setattr(g, n, int(r[n]))
    g.myid, g.nm, g.abc_type, g.field_bytes, g.device = 0, 3, 1, 8, 0
    g.dx, g.dy, g.dz, g.dt = float(o.cfg("dx")), float(o.cfg("dy")), float(o.cfg("dz")), float(np.float32(o.cfg("dt")))
    ijk, mo = o.sources(0)
    mij, prm = o_source_details(o, 0)
    sijk, _ = o.stations(0)
    nsrc, nst = len(mo), len(sijk)
    blob = [bytes(g), np.asarray(o.ts(), dtype=np.float32).tobytes(),
            np.array([nt, nsrc, nst, o.cfg("ntdec_w"), o.cfg("ntw"), 0, 0, 0], dtype=np.int32).tobytes(),
            np.array([o.cfg("M0"), o.cfg("UC")], dtype=np.float32).tobytes(), b"kupper".ljust(16, b"\0")]
    blob += [np.ascontiguousarray(o.field(0, n), dtype=np.float32).tobytes() for n in ("rho", "lam", "mu", "taup", "taus")]
    blob += [np.ascontiguousarray(o.imap(0, n), dtype=np.int32).tobytes() for n in ("kfs", "kob", "kfs_top", "kfs_bot", "kob_top", "kob_bot", "kbeg_a")]
    blob += [np.ascontiguousarray(o.profile(0, n), dtype=np.float32).tobytes() for n in ("gxc", "gxe", "gyc", "gye", "gzc", "gze")]
    blob += [np.ascontiguousarray(np.asarray(ijk, dtype=np.int32).T).tobytes(), np.asarray(mo, dtype=np.float64).tobytes(),
             np.ascontiguousarray(np.asarray(mij, dtype=np.float64).T).tobytes(), np.asarray(prm, dtype=np.float32).tobytes(),
             np.ascontiguousarray(np.asarray(sijk, dtype=np.int32).T).tobytes()]
    (tmp_path / "in.bin").write_bytes(b"".join(blob))
    exe = _build(tmp_path)
    p = subprocess.run([str(exe), "run", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    o.run(1, nt)
    raw = (tmp_path / "out.bin").read_bytes()
    shape = o.field(0, "Vx").shape
    n3 = int(np.prod(shape))
    fields = np.frombuffer(raw, dtype=np.float64, count=9 * n3, offset=12).reshape(9, *shape)
    sl = (slice(3, 3 + r["nyp"]), slice(3, 3 + r["nxp"]), slice(3, 3 + o.cfg("nz")))
    for q, n in enumerate(("Vx", "Vy", "Vz", "Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy")):
        assert np.array_equal(fields[q][sl], o.field(0, n)[sl]), n
    assert np.abs(fields[2]).max() > 0
    wav = np.frombuffer(raw, dtype=np.float32, offset=12 + 9 * n3 * 8).reshape(nst, 3, o.cfg("ntw"))
    np.testing.assert_array_equal(wav, o.wav(0))
