"""Shared test helpers: synthetic input.inf cases and the oracle -> device bridge.

`device_from_oracle` feeds the CUDA library (through the C ABI) with the setup arrays computed by the oracle, so
that the GPU kernels can be checked in isolation from the product's own host-side setup chain (which is checked
separately against the oracle in tests/test_host_setup.py).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

LHM_OCEAN = """# depth rho vp vs Qp Qs
  1.0   2.3   5.5   3.14   600   300
  3.0   2.4   6.0   3.55   400   200
  9.0   2.8   6.7   3.83   600   300
 15.0   3.2   7.8   4.46   600   300
"""

LHM_LAND = """# depth rho vp vs Qp Qs
  0.0   2.3   5.5   3.14   600   300
  3.0   2.4   6.0   3.55   400   200
  9.0   2.8   6.7   3.83   600   300
 15.0   3.2   7.8   4.46   600   300
"""


def write_case(d: Path, *, nx=48, ny=40, nz=44, nt=40, dx=0.5, dy=0.5, dz=0.5, dt=0.02, na=6, nproc_x=1, nproc_y=1,
               abc_type="pml", vmodel="lhm_ocean", zbeg=-3.0, sources=None, stations=None, stftype="kupper",
               stf_format="xym0ij", sdep_fit="asis", wav_format="sac", bf_mode=False, ntdec_w=2, ntdec_r=10, extra="", benchmark=False, title="case") -> Path:
    d = Path(d)
    d.mkdir(parents=True, exist_ok=True)
    xbeg, ybeg = -nx * dx / 2, -ny * dy / 2
    if sources is None:
        #   x    y    z   tbeg trise  mo   mxx myy mzz myz mxz mxy
        sources = ["0.3 -0.2 4.1 0.05 0.6 1e15 0.7 -0.3 0.5 0.4 -0.6 0.8",
                   "-2.1 1.7 6.3 0.10 0.5 4e14 -0.2 0.9 0.1 -0.5 0.3 0.6"]
    if stations is None:
        stations = ["0.0 0.0 0.0 st01 obb", "-3.2 2.1 2.0 st02 dep", "4.3 -3.9 0.0 st03 fsb", "2.2 4.4 0.0 st04 oba"]
    (d / "source.dat").write_text("# test sources\n" + "\n".join(sources) + "\n")
    (d / "stloc.xy").write_text("# test stations\n" + "\n".join(stations) + "\n")
    (d / "lhm_ocean.dat").write_text(LHM_OCEAN)
    (d / "lhm_land.dat").write_text(LHM_LAND)
    if vmodel == "uni":
        vm = "vmodel_type = 'uni'\n vp0 = 5.0\n vs0 = 2.9\n rho0 = 2.6\n qp0 = 300\n qs0 = 150\n topo0 = 0.4\n"
    elif vmodel.startswith("raw:"):   # the caller supplies the vmodel lines itself
        vm = vmodel[4:]
    else:
        vm = f"vmodel_type = 'lhm'\n fn_lhm = '{vmodel}.dat'\n"
    inf = f"""
 !! synthetic test case
 title = '{title}'
 odir = './out'
 ntdec_r = {ntdec_r}
 benchmark_mode = {'.true.' if benchmark else '.false.'}
 nproc_x = {nproc_x}
 nproc_y = {nproc_y}
 nx = {nx}
 ny = {ny}
 nz = {nz}
 nt = {nt}
 dx = {dx}
 dy = {dy}
 dz = {dz}
 dt = {dt}
 vcut = 0.0
 xbeg = {xbeg}
 ybeg = {ybeg}
 zbeg = {zbeg}
 tbeg = 0.0
 fq_min = 0.05
 fq_max = 5.0
 fq_ref = 1.0
 sw_wav_v = .true.
 ntdec_w = {ntdec_w}
 st_format = 'xy'
 fn_stloc = 'stloc.xy'
 wav_format = '{wav_format}'
 stf_format = '{stf_format}'
 stftype = '{stftype}'
 fn_stf = "source.dat"
 sdep_fit = '{sdep_fit}'
 bf_mode = {'.true.' if bf_mode else '.false.'}
 abc_type = '{abc_type}'
 na = {na}
 munk_profile = .true.
 {vm}
 {extra}
"""
    p = d / "input.inf"
    p.write_text(inf)
    return p


def device_from_oracle(o, q: int, field_dtype=np.float64, device: int = -1):
    """Create a DeviceRank mirroring oracle rank q, with every setup array taken from the oracle."""
    from openswpc_b200.device import DeviceRank, RankGeometry

    r = o.rank(q)
    geom = RankGeometry(nx=o.cfg("nx"), ny=o.cfg("ny"), nz=o.cfg("nz"), nproc_x=o.cfg("nproc_x"), nproc_y=o.cfg("nproc_y"),
                        myid=q, ibeg=r["ibeg"], iend=r["iend"], jbeg=r["jbeg"], jend=r["jend"], ibeg_k=r["ibeg_k"],
                        iend_k=r["iend_k"], jbeg_k=r["jbeg_k"], jend_k=r["jend_k"], kbeg_k=r["kbeg_k"], kend_k=r["kend_k"],
                        na=o.cfg("na"))
    abc = o.cfg("abc_type")
    nm = o.cfg("nm")
    dev = DeviceRank(geom, dx=o.cfg("dx"), dy=o.cfg("dy"), dz=o.cfg("dz"), dt=o.cfg("dt"), nm=nm, abc_type=abc,
                     ts=o.ts() if nm > 0 else None, field_dtype=field_dtype, device=device)
    med = [o.field(q, n).astype(np.float32) for n in ("rho", "lam", "mu", "taup", "taus")]
    maps = [o.imap(q, n) for n in ("kfs", "kob", "kfs_top", "kfs_bot", "kob_top", "kob_bot", "kbeg_a")]
    dev.upload_medium(*med, *maps)
    if abc == "pml":
        dev.setup_pml(*[o.profile(q, n) for n in ("gxc", "gxe", "gyc", "gye", "gzc", "gze")])
    else:
        dev.setup_cerjan(*[o.profile(q, n) for n in ("gx_c", "gx_b", "gy_c", "gy_b", "gz_c", "gz_b")])
    ijk, mo = o.sources(q)
    if len(mo):
        mij, prm = o_source_details(o, q)
        dev.set_sources(ijk, mo, mij, prm, stftype=o.cfg("stftype"), bf_mode=bool(o.cfg("bf_mode")), tbeg=0.0)
    sijk, _ = o.stations(q)
    if len(sijk):
        dev.set_stations(sijk, o.cfg("ntdec_w"), o.cfg("ntw"), o.cfg("M0"), o.cfg("UC"))
    return dev


def o_source_details(o, q):
    """(mij (nsrc,6), srcprm (nsrc,2)) of oracle rank q."""
    import ctypes as C

    n = o.rank(q)["nsrc"]
    mij = np.zeros((max(n, 1), 6), dtype=np.float64)
    prm = np.zeros((max(n, 1), 2), dtype=np.float32)
    o.lib.ora_get_source_details.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    o.lib.ora_get_source_details.restype = C.c_int
    o.lib.ora_get_source_details(o.h, q, mij.ctypes.data_as(C.POINTER(C.c_double)), prm.ctypes.data_as(C.POINTER(C.c_float)))
    return mij[:n], prm[:n]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.sqrt(np.sum((a - b) ** 2))
    n = np.sqrt(np.sum(b ** 2))
    return d / n if n > 0 else d


def write_rmed(path: Path, xi: np.ndarray, dx=0.5) -> None:
    """A random-media volume laid out as tools/gen_rmed3d.f90:91-135 creates it: netCDF classic (NF90_CLOBBER), dimensions
    x, y, z, variables x, y, z and -- 4th, which is how m_rdrmed.f90:93 finds it -- the volume with x fastest.  Written
    byte by byte from the classic-format specification (scipy's writer reorders variables); tests read it back with scipy's
    reader as an independent check.  xi has shape (nz, ny, nx)."""
    import struct

    nz, ny, nx = xi.shape

    def name(s):
        b = s.encode()
        return struct.pack(">I", len(b)) + b + b"\0" * (-len(b) % 4)

    def att_text(k, v):
        b = v.encode()
        return name(k) + struct.pack(">II", 2, len(b)) + b + b"\0" * (-len(b) % 4)

    dims = struct.pack(">II", 0x0A, 3) + b"".join(name(n) + struct.pack(">I", m) for n, m in (("x", nx), ("y", ny), ("z", nz)))
    gatts = struct.pack(">II", 0x0C, 1) + att_text("title", "random media")
    shapes = [("x", [0], nx), ("y", [1], ny), ("z", [2], nz), ("random media", [2, 1, 0], nx * ny * nz)]

    def var_list(begins):
        out = struct.pack(">II", 0x0B, len(shapes))
        for (n, dimids, cnt), beg in zip(shapes, begins):
            out += name(n) + struct.pack(">I", len(dimids)) + b"".join(struct.pack(">I", d) for d in dimids)
            out += struct.pack(">II", 0, 0)                      # no attributes
            out += struct.pack(">III", 5, cnt * 4, beg)          # NC_FLOAT, vsize, begin
        return out

    head = b"CDF\x01" + struct.pack(">I", 0) + dims + gatts
    off = len(head) + len(var_list([0] * 4))
    begins = []
    for _, _, cnt in shapes:
        begins.append(off)
        off += cnt * 4
    with open(path, "wb") as f:
        f.write(head + var_list(begins))
        for n in (nx, ny, nz):
            f.write((np.arange(n) * dx).astype(">f4").tobytes())
        f.write(np.ascontiguousarray(xi).astype(">f4").tobytes())


def write_rmed2d(path: Path, xi: np.ndarray, dx=0.5) -> None:
    """A random-media section as tools/gen_rmed2d.f90:87-123 creates it: netCDF classic, dimensions x, z, variables x, z and --
    3rd, which is how rdrmed__2d (m_rdrmed.f90:38) finds it -- the section with x fastest.  xi has shape (nz, nx)."""
    import struct

    nz, nx = xi.shape

    def name(s):
        b = s.encode()
        return struct.pack(">I", len(b)) + b + b"\0" * (-len(b) % 4)

    def att_text(k, v):
        b = v.encode()
        return name(k) + struct.pack(">II", 2, len(b)) + b + b"\0" * (-len(b) % 4)

    dims = struct.pack(">II", 0x0A, 2) + b"".join(name(n) + struct.pack(">I", m) for n, m in (("x", nx), ("z", nz)))
    gatts = struct.pack(">II", 0x0C, 1) + att_text("title", "random media")
    shapes = [("x", [0], nx), ("z", [1], nz), ("random media", [1, 0], nx * nz)]

    def var_list(begins):
        out = struct.pack(">II", 0x0B, len(shapes))
        for (n, dimids, cnt), beg in zip(shapes, begins):
            out += name(n) + struct.pack(">I", len(dimids)) + b"".join(struct.pack(">I", d) for d in dimids)
            out += struct.pack(">II", 0x0C, 1) + att_text("long_name", n)     # one attribute per variable, as the tool writes
            out += struct.pack(">III", 5, cnt * 4, beg)
        return out

    head = b"CDF\x01" + struct.pack(">I", 0) + dims + gatts
    off = len(head) + len(var_list([0] * 3))
    begins = []
    for _, _, cnt in shapes:
        begins.append(off)
        off += cnt * 4
    with open(path, "wb") as f:
        f.write(head + var_list(begins))
        for n in (nx, nz):
            f.write((np.arange(n) * dx).astype(">f4").tobytes())
        f.write(np.ascontiguousarray(xi).astype(">f4").tobytes())


def write_grd(path: Path, lon: np.ndarray, lat: np.ndarray, z: np.ndarray, zdtype=">f4") -> None:
    """A GMT-style 2-D grid in the netCDF classic container (what `gmt grdconvert in.grd out.grd=cf` gives): dimensions x, y,
    variables x(x), y(y) as doubles and z(y, x).  z has shape (nlat, nlon), depths in metres, positive down."""
    import struct

    nlat, nlon = z.shape

    def name(s):
        b = s.encode()
        return struct.pack(">I", len(b)) + b + b"\0" * (-len(b) % 4)

    ztype = 5 if zdtype == ">f4" else 6
    zsize = 4 if zdtype == ">f4" else 8
    dims = struct.pack(">II", 0x0A, 2) + name("x") + struct.pack(">I", nlon) + name("y") + struct.pack(">I", nlat)
    title = b"synthetic grid"
    gatts = struct.pack(">II", 0x0C, 1) + name("title") + struct.pack(">II", 2, len(title)) + title + b"\0" * (-len(title) % 4)
    shapes = [("x", [0], 6, nlon * 8), ("y", [1], 6, nlat * 8), ("z", [1, 0], ztype, nlon * nlat * zsize)]

    def var_list(begins):
        out = struct.pack(">II", 0x0B, len(shapes))
        for (n, dimids, ty, nbytes), beg in zip(shapes, begins):
            out += name(n) + struct.pack(">I", len(dimids)) + b"".join(struct.pack(">I", d) for d in dimids)
            out += struct.pack(">II", 0, 0) + struct.pack(">III", ty, (nbytes + 3) // 4 * 4, beg)
        return out

    head = b"CDF\x01" + struct.pack(">I", 0) + dims + gatts
    off = len(head) + len(var_list([0] * 3))
    begins = []
    for _, _, _, nbytes in shapes:
        begins.append(off)
        off += (nbytes + 3) // 4 * 4
    with open(path, "wb") as f:
        f.write(head + var_list(begins))
        f.write(np.asarray(lon, dtype=">f8").tobytes())
        f.write(np.asarray(lat, dtype=">f8").tobytes())
        f.write(np.ascontiguousarray(z).astype(zdtype).tobytes())
