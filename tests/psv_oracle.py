"""ctypes binding of the swpc_psv CPU oracle (oracle/psv.c, TEST INFRASTRUCTURE) and helpers that build the device state
of the product from the oracle's setup so that both start from identical inputs."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import oracle_lib

FIELDS = ("Vx", "Vz", "Sxx", "Szz", "Sxz")
MEDIUM = ("rho", "lam", "mu", "taup", "taus")
MAPS = ("kfs", "kob", "kfs_top", "kfs_bot", "kob_top", "kob_bot", "kbeg_a")
_RI = {n: i for i, n in enumerate(("ibeg", "iend", "ibeg_k", "iend_k", "kbeg_k", "kend_k", "nsrc", "nst", "nzm", "nxm", "ibeg_m", "kbeg_m",
                                   "kbeg_min", "nxp"))}
_CI = {n: i for i, n in enumerate(("nx", "nz", "nt", "na", "nm", "nproc_x", "ntw", "ntdec_w", "ntdec_r", "bf_mode", "pw_mode", "sw_v", "sw_u",
                                   "sw_stress", "sw_strain"))}
_CV = {n: i for i, n in enumerate(("vmin", "vmax", "fmax", "fcut", "M0", "UC", "zeta", "d2", "dt", "xbeg", "zbeg", "dx", "dz", "evlo", "evla"))}
_CV.update({"tbeg": 47, "r20x": 48, "r20z": 49})


def _bind(lib):
    if getattr(lib, "_psv_bound", False):
        return lib
    vp, ci, cd, cc, fp = C.c_void_p, C.c_int, C.c_double, C.c_char_p, C.POINTER(C.c_float)
    lib.psv_create.restype = vp
    lib.psv_create.argtypes = [cc, cc, ci, ci, ci]
    lib.psv_create_from_text.restype = vp
    lib.psv_create_from_text.argtypes = [cc, cc, ci, ci, ci]
    lib.psv_destroy.argtypes = [vp]
    lib.psv_last_error.restype = cc
    lib.psv_set_exedate.argtypes = [vp, ci, ci]
    lib.psv_step.argtypes = [vp, ci]
    lib.psv_run.restype = ci
    lib.psv_run.argtypes = [vp, ci, ci, fp, ci]
    lib.psv_vmax.argtypes = [vp, fp]
    lib.psv_nranks.restype = ci
    lib.psv_nranks.argtypes = [vp]
    lib.psv_rank_int.restype = ci
    lib.psv_rank_int.argtypes = [vp, ci, ci]
    lib.psv_cfg_value.restype = cd
    lib.psv_cfg_value.argtypes = [vp, ci]
    lib.psv_cfg_int.restype = ci
    lib.psv_cfg_int.argtypes = [vp, ci]
    lib.psv_cfg_str.restype = cc
    lib.psv_cfg_str.argtypes = [vp, ci]
    for f in ("psv_get_field", "psv_set_field"):
        getattr(lib, f).restype = ci
        getattr(lib, f).argtypes = [vp, ci, cc, C.POINTER(cd)]
    lib.psv_redetect_surface.argtypes = [vp]
    lib.psv_get_memvar.restype = ci
    lib.psv_get_memvar.argtypes = [vp, ci, cc, fp]
    lib.psv_get_map.restype = ci
    lib.psv_get_map.argtypes = [vp, ci, cc, C.POINTER(ci)]
    lib.psv_get_profile.restype = ci
    lib.psv_get_profile.argtypes = [vp, ci, cc, fp]
    lib.psv_get_sources.restype = ci
    lib.psv_get_sources.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(cd)]
    lib.psv_get_stations.restype = ci
    lib.psv_get_stations.argtypes = [vp, ci, C.POINTER(ci), C.c_char_p]
    lib.psv_get_wav.restype = ci
    lib.psv_get_wav.argtypes = [vp, ci, ci, fp]
    lib.psv_write_sac.restype = ci
    lib.psv_write_sac.argtypes = [vp, cc]
    lib.psv_snap_info.argtypes = [vp, C.POINTER(ci)]
    lib.psv_snap_coords.argtypes = [vp, fp, fp]
    lib.psv_snap_nrec.argtypes = [vp, ci]
    lib.psv_snap_nrec.restype = ci
    lib.psv_snap_rec.argtypes = [vp, ci, ci, fp, C.POINTER(ci)]
    lib.psv_snap_medium.argtypes = [vp, ci, fp]
    lib._psv_bound = True
    return lib


class PsvOracle:
    """One swpc_psv run of the CPU oracle: all MPI ranks emulated in this process."""

    def __init__(self, inf, base_dir=".", nm=3, nproc_x=0, nt=0, sp=False, text=None):
        self.sp = sp
        self.lib = _bind(oracle_lib.lib("sp" if sp else "dp"))
        if text is not None:
            self.h = self.lib.psv_create_from_text(text.encode(), str(base_dir).encode(), nm, nproc_x, nt)
        else:
            self.h = self.lib.psv_create(str(inf).encode(), str(base_dir).encode(), nm, nproc_x, nt)
        if not self.h:
            raise RuntimeError("psv oracle: " + self.lib.psv_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.nranks = self.lib.psv_nranks(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.psv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cfg(self, name):
        if name in _CI:
            return self.lib.psv_cfg_int(self.h, _CI[name])
        if name in _CV:
            return self.lib.psv_cfg_value(self.h, _CV[name])
        if name in ("title", "odir", "abc_type", "stftype"):
            return self.lib.psv_cfg_str(self.h, ("title", "odir", "abc_type", "stftype").index(name)).decode()
        if name in ("ts", "c1", "c2", "d1"):
            base = {"ts": 15, "c1": 23, "c2": 31, "d1": 39}[name]
            return np.array([self.lib.psv_cfg_value(self.h, base + m) for m in range(self.cfg("nm"))], dtype=np.float32)
        raise KeyError(name)

    def rank(self, q):
        return {n: self.lib.psv_rank_int(self.h, q, i) for n, i in _RI.items()}

    def shape2(self, q):
        r = self.rank(q)
        return (r["nxm"], r["nzm"])

    def field(self, q, name):
        out = np.zeros(self.shape2(q), dtype=np.float64)
        assert self.lib.psv_get_field(self.h, q, name.encode(), out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        return out

    def set_field(self, q, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.shape2(q)
        assert self.lib.psv_set_field(self.h, q, name.encode(), a.ctypes.data_as(C.POINTER(C.c_double))) == 0

    def redetect_surface(self):
        self.lib.psv_redetect_surface(self.h)

    def memvar(self, q, name):
        out = np.zeros(self.shape2(q) + (self.cfg("nm"),), dtype=np.float32)
        assert self.lib.psv_get_memvar(self.h, q, name.encode(), out.ctypes.data_as(C.POINTER(C.c_float))) == 0
        return out

    def map(self, q, name):
        out = np.zeros(self.shape2(q)[0], dtype=np.int32)
        assert self.lib.psv_get_map(self.h, q, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int))) == 0
        return out

    def profile(self, q, name):
        r = self.rank(q)
        n = {"gxc": 4 * (r["iend"] - r["ibeg"] + 1), "gxe": 4 * (r["iend"] - r["ibeg"] + 1), "gzc": 4 * self.cfg("nz"), "gze": 4 * self.cfg("nz"),
             "gx_c": r["nxm"], "gx_b": r["nxm"], "gz_c": r["nzm"], "gz_b": r["nzm"]}[name]
        out = np.zeros(n, dtype=np.float32)
        assert self.lib.psv_get_profile(self.h, q, name.encode(), out.ctypes.data_as(C.POINTER(C.c_float))) == n
        return out

    def sources(self, q):
        n = self.rank(q)["nsrc"]
        ik = np.zeros((max(n, 1), 2), dtype=np.int32)
        val = np.zeros((max(n, 1), 6), dtype=np.float64)
        self.lib.psv_get_sources(self.h, q, ik.ctypes.data_as(C.POINTER(C.c_int)), val.ctypes.data_as(C.POINTER(C.c_double)))
        return ik[:n], val[:n]

    def stations(self, q):
        n = self.rank(q)["nst"]
        ik = np.zeros((max(n, 1), 2), dtype=np.int32)
        names = C.create_string_buffer(9 * max(n, 1))
        self.lib.psv_get_stations(self.h, q, ik.ctypes.data_as(C.POINTER(C.c_int)), names)
        return ik[:n], [names.raw[9 * i:9 * i + 9].split(b"\0")[0].decode() for i in range(n)]

    def wav(self, q, prod=0):
        n = self.rank(q)["nst"]
        out = np.zeros((n, 2 if prod < 2 else 3, max(self.cfg("ntw"), 0)), dtype=np.float32)
        if n and out.size:
            self.lib.psv_get_wav(self.h, q, prod, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def step(self, it):
        self.lib.psv_step(self.h, it)

    def run(self, it0, it1):
        nrec = max((it1 - it0 + 1) // max(self.cfg("ntdec_r"), 1) + 2, 1)
        vm = np.zeros((nrec, 2), dtype=np.float32)
        n = self.lib.psv_run(self.h, it0, it1, vm.ctypes.data_as(C.POINTER(C.c_float)), nrec)
        return vm[:n]

    def vmax(self):
        out = np.zeros(2, dtype=np.float32)
        self.lib.psv_vmax(self.h, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def gather(self, name):
        """owned cells of every rank -> global (nx, nz) array (k = 1..nz)."""
        nx, nz = self.cfg("nx"), self.cfg("nz")
        out = np.zeros((nx, nz))
        for q in range(self.nranks):
            r = self.rank(q)
            f = self.field(q, name)
            out[r["ibeg"] - 1:r["iend"], :] = f[3:3 + r["iend"] - r["ibeg"] + 1, 3:3 + nz]
        return out

    # ---- snapshots (m_snap.f90)
    def snap_info(self):
        v = (C.c_int * 8)()
        self.lib.psv_snap_info(self.h, v)
        return dict(zip(["idec", "kdec", "ntdec_s", "nxs", "nzs", "sw_ps", "sw_v", "sw_u"], list(v)))

    def snap_coords(self):
        i = self.snap_info()
        x, z = np.zeros(i["nxs"], dtype=np.float32), np.zeros(i["nzs"], dtype=np.float32)
        fp = C.POINTER(C.c_float)
        self.lib.psv_snap_coords(self.h, x.ctypes.data_as(fp), z.ctypes.data_as(fp))
        return x, z

    def snap_records(self, p):
        """(nrec, 2, nzs, nxs) float32 and the list of it0 of product p (0 ps, 1 v, 2 u)"""
        i = self.snap_info()
        n = self.lib.psv_snap_nrec(self.h, p)
        out = np.zeros((n, 2, i["nzs"], i["nxs"]), dtype=np.float32)
        its = []
        for r in range(n):
            it0 = C.c_int()
            self.lib.psv_snap_rec(self.h, p, r, out[r].ctypes.data_as(C.POINTER(C.c_float)), C.byref(it0))
            its.append(it0.value)
        return out, its

    def snap_medium(self, which):
        i = self.snap_info()
        out = np.zeros((i["nzs"], i["nxs"]), dtype=np.float32)
        self.lib.psv_snap_medium(self.h, which, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def write_sac(self, odir):
        return self.lib.psv_write_sac(self.h, str(odir).encode())

    def set_exedate(self, exedate, tz=0):
        self.lib.psv_set_exedate(self.h, exedate, tz)


def psv_case_text(*, nx=96, nz=80, nt=60, dx=0.5, dz=0.5, dt=0.02, na=10, abc="pml", nproc_x=1, vmodel="uni", extra="", fn_stf="source.dat",
                  fn_stloc="stloc.xy", zbeg=-5.0, stftype="kupper", bf_mode=False, stf_format="xym0ij", products="v", ntdec_w=2):
    sw = {p: (".true." if p in products.split(",") else ".false.") for p in ("v", "u", "stress", "strain")}
    return f"""
 title = 'psvtest'
 odir = './out'
 nproc_x = {nproc_x}
 nx = {nx}
 nz = {nz}
 nt = {nt}
 dx = {dx}
 dz = {dz}
 dt = {dt}
 na = {na}
 zbeg = {zbeg}
 abc_type = '{abc}'
 vmodel_type = '{vmodel}'
 vp0 = 5.0
 vs0 = 2.9
 rho0 = 2.6
 qp0 = 200
 qs0 = 100
 topo0 = 0.0
 fq_min = 0.05
 fq_max = 5.0
 fq_ref = 1.0
 fn_stf = '{fn_stf}'
 stftype = '{stftype}'
 stf_format = '{stf_format}'
 bf_mode = {'.true.' if bf_mode else '.false.'}
 fn_stloc = '{fn_stloc}'
 st_format = 'xy'
 ntdec_w = {ntdec_w}
 sw_wav_v = {sw['v']}
 sw_wav_u = {sw['u']}
 sw_wav_stress = {sw['stress']}
 sw_wav_strain = {sw['strain']}
 ntdec_r = 10
{extra}
"""


def write_psv_files(td: Path, sources=None, stations=None):
    td = Path(td)
    src = sources or ["0.3 0.0 4.2 0.05 0.6 1e15 0.7 0.0 -0.3 0.0 0.5 0.0"]
    (td / "source.dat").write_text("# x y z tbeg trise mo mxx myy mzz myz mxz mxy\n" + "\n".join(src) + "\n")
    st = stations or ["-6.1 0.0 0.0 st01 obb", "5.3 0.0 3.0 st02 dep", "11.2 0.0 0.0 st03 fsb", "0.2 0.0 8.0 st04 dep"]
    (td / "stloc.xy").write_text("# x y z name zsw\n" + "\n".join(st) + "\n")


def psv_device_from_oracle(o: PsvOracle, q: int, device: int = 0, field_dtype=None):
    """Create the product's device state for rank q from the oracle's setup arrays (identical inputs on both sides)."""
    from openswpc_b200.psv_device import PsvGeometry, PsvRank

    r = o.rank(q)
    geom = PsvGeometry(nx=o.cfg("nx"), nz=o.cfg("nz"), nproc_x=o.cfg("nproc_x"), myid=q, ibeg=r["ibeg"], iend=r["iend"], ibeg_k=r["ibeg_k"],
                       iend_k=r["iend_k"], kend_k=r["kend_k"], na=o.cfg("na"))
    if field_dtype is None:
        field_dtype = np.float32 if o.sp else np.float64
    abc = o.cfg("abc_type")
    dev = PsvRank(geom, dx=o.cfg("dx"), dz=o.cfg("dz"), dt=o.cfg("dt"), nm=o.cfg("nm"), abc_type=abc, ts=o.cfg("ts"), field_dtype=field_dtype,
                  device=device)
    dev.upload_medium(*[o.field(q, n) for n in MEDIUM], *[o.map(q, n) for n in MAPS])
    if abc == "pml":
        dev.setup_pml(*[o.profile(q, n) for n in ("gxc", "gxe", "gzc", "gze")])
    else:
        dev.setup_cerjan(*[o.profile(q, n) for n in ("gx_c", "gx_b", "gz_c", "gz_b")])
    ik, val = o.sources(q)
    bf = bool(o.cfg("bf_mode"))
    dev.set_sources(ik[:, 0], ik[:, 1], val[:, 0], val[:, 1], val[:, 2], val[:, 3], val[:, 4:6].astype(np.float32), stftype=o.cfg("stftype"),
                    bf_mode=bf, tbeg=o.cfg("tbeg"))
    sik, _ = o.stations(q)
    if len(sik):
        dev.set_stations(sik[:, 0], sik[:, 1], o.cfg("ntdec_w"), o.cfg("ntw"), o.cfg("M0"), o.cfg("UC"), sw_v=o.cfg("sw_v"), sw_u=o.cfg("sw_u"),
                         sw_stress=o.cfg("sw_stress"), sw_strain=o.cfg("sw_strain"))
    return dev
