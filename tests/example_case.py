"""The reference's example run (example/input.inf + lhm.dat + source.dat + stloc.xy) restated as a generated case, so
that tests on the GPU box (where /root/reference does not exist) can run exactly that configuration."""
from pathlib import Path

EXAMPLE_LHM = """# depth rho vp vs Qp Qs  (values of example/lhm.dat)
      0          2.3       5.5      3.14      600     300
      3          2.4       6.0      3.55      600     300
     18          2.8       6.7      3.83      600     300
     33          3.2       7.8      4.46      600     300
    100          3.3       8.0      4.57      600     300
    225          3.4       8.4      4.80      600     300
    325          3.5       8.6      4.91      600     300
    425          3.7       9.3      5.31      600     300
"""

# example/example.out:17-36 -- the reference's only known answers for the time loop (max |Vx|,|Vy|,|Vz| every 50 steps)
EXAMPLE_OUT_LINES = """
 2.94E-05  2.94E-05  1.59E-04
 1.95E-04  1.95E-04  8.87E-04
 1.04E-04  1.04E-04  4.87E-04
 2.62E-05  2.62E-05  4.89E-05
 1.32E-05  1.32E-05  3.59E-05
 1.13E-05  1.13E-05  3.16E-05
 1.26E-05  1.26E-05  2.67E-05
 1.13E-05  1.13E-05  2.33E-05
 9.68E-06  9.68E-06  2.07E-05
 9.22E-06  9.22E-06  1.98E-05
 8.90E-06  8.90E-06  1.90E-05
 8.37E-06  8.37E-06  1.78E-05
 7.78E-06  7.78E-06  1.68E-05
 7.63E-06  7.63E-06  1.59E-05
 7.47E-06  7.47E-06  1.53E-05
 7.01E-06  7.01E-06  1.45E-05
 6.79E-06  6.79E-06  1.35E-05
 7.07E-06  7.07E-06  1.28E-05
 7.38E-06  7.38E-06  1.20E-05
 7.22E-06  7.22E-06  1.14E-05
""".split("\n")[1:-1]


def es92(x: float) -> str:
    """Fortran ES9.2"""
    s = f"{x:.2E}"
    m, e = s.split("E")
    return f"{m}E{int(e):+03d}"


# example/input.inf:55-83, verbatim values
EXAMPLE_SNAP_BLOCK = """
  snp_format       = 'netcdf'
  xy_ps%sw         = .false.
  xz_ps%sw         = .true.
  yz_ps%sw         = .false.
  fs_ps%sw         = .false.
  ob_ps%sw         = .true.
  xy_v%sw          = .false.
  xz_v%sw          = .true.
  yz_v%sw          = .false.
  fs_v%sw          = .false.
  ob_v%sw          = .true.
  xy_u%sw          = .false.
  xz_u%sw          = .true.
  yz_u%sw          = .false.
  fs_u%sw          = .false.
  ob_u%sw          = .true.
  z0_xy            =  7.0
  x0_yz            =  0.0
  y0_xz            =  0.0
  ntdec_s          = 5
  idec             = 2
  jdec             = 2
  kdec             = 2
"""


def write_example(d: Path, *, n=384, nz=384, nt=1000, nproc_x=2, nproc_y=2, snapshots=False) -> Path:
    """example/input.inf (keys that the in-scope path reads), optionally on a smaller box centred on the source;
    `snapshots`: with the snapshot block of example/input.inf:55-83 (six netCDF products every 5 steps)."""
    d = Path(d)
    d.mkdir(parents=True, exist_ok=True)
    (d / "lhm.dat").write_text(EXAMPLE_LHM)
    (d / "source.dat").write_text("   0.0    0.0  2.0   0.1    4.0  1.e15  0.8165  0.8165  0.8165      0.0      0.0      0.0\n")
    (d / "stloc.xy").write_text("     0.0   0.0   0.0       st01   obb\n   -10.0  -5.0   0.0       st02   obb\n    10.0   5.0   0.0       st03   obb\n")
    beg = -0.5 * n / 2
    (d / "input.inf").write_text(f"""
  title            = 'swpc'
  odir             = './out'
  ntdec_r          = 50
  strict_mode      = .false.
  nproc_x          = {nproc_x}
  nproc_y          = {nproc_y}
  nx               = {n}
  ny               = {n}
  nz               = {nz}
  nt               = {nt}
  dx               = 0.5
  dy               = 0.5
  dz               = 0.5
  dt               = 0.02
  vcut             = 1.5
  xbeg             = {beg}
  ybeg             = {beg}
  zbeg             = -10.0
  tbeg             = 0.0
  clon             = 139.7604
  clat             = 35.7182
  phi              = 0.0
  fq_min           = 0.02
  fq_max           = 2.00
  fq_ref           = 1.0
  sw_wav_v         = .true.
  ntdec_w          = 5
  st_format        = 'xy'
  fn_stloc         = './stloc.xy'
  wav_format       = 'sac'
  stf_format       = 'xym0ij'
  stftype          = 'kupper'
  fn_stf           = "./source.dat"
  sdep_fit         = 'asis'
  abc_type         = 'pml'
  na               = 20
  stabilize_pml    = .false.
  vmodel_type      = 'lhm'
  munk_profile     = .true.
  earth_flattening = .false.
  fn_lhm           = 'lhm.dat'
""" + (EXAMPLE_SNAP_BLOCK if snapshots else ""))
    return d / "input.inf"
