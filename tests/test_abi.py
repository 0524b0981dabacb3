"""The drop-in boundary itself, without a GPU: the C-ABI library loads and exports every symbol `include/*.h` declares (and
nothing is bound in Python or Fortran that the headers do not declare), the compute entry points refuse to run without a
CUDA device instead of falling back to anything, and the product never touches the oracle."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

HEADERS = {"swpc3d_b200.h": r"\b(swpc3d_(?!host_)\w+)\s*\(", "swpc3d_host.h": r"\b(swpc3d_host_\w+)\s*\(",
           "swpcpsv_b200.h": r"\b(swpcpsv_(?!host_)\w+)\s*\(", "swpcpsv_host.h": r"\b(swpcpsv_host_\w+)\s*\("}


def _declared(hdr):
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / hdr).read_text(), flags=re.S)   # prototypes only, not the comments
    return set(re.findall(HEADERS[hdr], text))


@pytest.mark.parametrize("hdr", list(HEADERS))
def test_library_exports_every_declared_symbol(hdr):
    from openswpc_b200 import _lib

    lib = _lib.load()
    names = _declared(hdr)
    assert len(names) >= 15, names
    for s in sorted(names):
        assert hasattr(lib, s), f"{hdr} declares {s} but the library does not export it"


def test_python_face_binds_exactly_the_kernel_abi():
    from openswpc_b200 import _lib

    assert set(_lib.SYMBOLS) == _declared("swpc3d_b200.h")
    assert set(_lib.PSV_SYMBOLS) == _declared("swpcpsv_b200.h")
    # the host faces name their entry points in the source: none may be missing from the headers
    for py, hdr in (("swpc3d.py", "swpc3d_host.h"), ("swpc_psv.py", "swpcpsv_host.h")):
        used = set(re.findall(HEADERS[hdr].replace(r"\s*\(", ""), (ROOT / "openswpc_b200" / py).read_text()))
        assert used and used <= _declared(hdr), used - _declared(hdr)


@pytest.mark.parametrize("mod,hdr", [("m_swpc3d_b200.f90", "swpc3d_b200.h"), ("m_swpcpsv_b200.f90", "swpcpsv_b200.h")])
def test_fortran_module_binds_the_whole_kernel_abi(mod, hdr):
    """The iso_c_binding module a Fortran maintainer adds has one `bind(c, name=...)` interface per entry point
    of the header, is `public` about each of them, and binds nothing the header does not declare."""
    src = (ROOT / "openswpc_b200" / "fortran" / mod).read_text()
    bound = set(re.findall(r"bind\(c,\s*name\s*=\s*'(\w+)'\)", src))
    declared = _declared(hdr)
    assert bound == declared, (sorted(declared - bound), sorted(bound - declared))
    public = set(re.findall(r"\b(\w+)\b", " ".join(re.findall(r"^\s*public\s*::(.*)$", src, flags=re.M))))
    assert declared <= public, sorted(declared - public)


def test_compute_entry_points_fail_loudly_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helpers import write_case
    from openswpc_b200._lib import Swpc3dError
    from openswpc_b200.device import DeviceRank, RankGeometry
    from openswpc_b200.swpc3d import Swpc3d, Swpc3dHostError

    geom = RankGeometry(nx=64, ny=64, nz=64, nproc_x=1, nproc_y=1, myid=0, ibeg=1, iend=64, jbeg=1, jend=64, ibeg_k=11, iend_k=54, jbeg_k=11,
                        jend_k=54, kbeg_k=1, kend_k=54, na=10)
    with pytest.raises(Swpc3dError, match="no CUDA device"):
        DeviceRank(geom, dx=0.5, dy=0.5, dz=0.5, dt=0.01, nm=0, abc_type="pml")
    run = Swpc3d(write_case(tmp_path, nt=4), base_dir=tmp_path, nm=3)        # the setup chain is host code and works ...
    with pytest.raises(Swpc3dHostError, match="no CUDA device"):             # ... the time loop exists on the GPU only
        run.attach_device(0)
    with pytest.raises(Swpc3dHostError):
        run.run(1, 4)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under openswpc_b200/ or include/ may import, link, open or name it."""
    offenders = []
    for p in list((ROOT / "openswpc_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".f90"):
            text = p.read_text(errors="ignore")
            if re.search(r"liboracle|oracle/|import oracle|from oracle|ora_\w+\(|psv_oracle|oracle_lib", text):
                offenders.append(str(p.relative_to(ROOT)))
    assert not offenders, offenders
    built = (ROOT / "openswpc_b200" / "_lib.py").read_text()
    assert "oracle" not in built.lower()
