"""Development probe: the reciprocity experiment of tests/test_gpu_reciprocity.py under variations (elastic, shorter run, thicker
absorber, deeper station) to see what limits the agreement."""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import test_gpu_reciprocity as T  # noqa: E402
from helpers import rel_l2  # noqa: E402
from openswpc_b200.swpc3d import Swpc3d  # noqa: E402


def run(label, nm=3, **over):
    names = ["Mxx", "Myz", "fx", "fz"]
    idx = {"Mxx": 0, "Myy": 1, "Mzz": 2, "Myz": 3, "Mxz": 4, "Mxy": 5, "fx": 6, "fy": 7, "fz": 8}
    old = dict(T.COMMON)
    T.COMMON.update(over)
    T.NT = T.COMMON["nt"]
    orig = Swpc3d.__init__

    def init(self, *a, **k):
        k["nm"] = nm
        orig(self, *a, **k)
    Swpc3d.__init__ = init
    try:
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            fwd = {}
            for n in names:
                q = idx[n]
                if q < 6:
                    m = [0.0] * 6; m[q] = 1.0
                    fwd[n], _ = T._forward(td / f"f{n}", mij=m)
                else:
                    f = [0.0] * 3; f[q - 6] = 1.0
                    fwd[n], _ = T._forward(td / f"f{n}", f=f)
            out = []
            for c, cmp in enumerate("xz"):
                gf, _ = T._reciprocal(td / f"r{cmp}", cmp)
                cc = "xyz".index(cmp)
                for n in names:
                    a, b = fwd[n][cc], gf[idx[n]]
                    if np.abs(a).max() > 0:
                        out.append(f"U{cmp}<-{n} {rel_l2(b, a):.4f}")
            print(label, " ".join(out), flush=True)
    finally:
        Swpc3d.__init__ = orig
        T.COMMON.clear(); T.COMMON.update(old)


import json
def run2(label, nm=3, half=True, **over):
    """all 27 pairs, forward velocity advanced by half a sample for the moment-tensor pairs"""
    names = ["Mxx", "Myy", "Mzz", "Myz", "Mxz", "Mxy", "fx", "fy", "fz"]
    old = dict(T.COMMON)
    T.COMMON.update(over)
    T.NT = T.COMMON["nt"]
    orig = Swpc3d.__init__

    def init(self, *a, **k):
        k["nm"] = nm
        orig(self, *a, **k)
    Swpc3d.__init__ = init
    try:
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            fwd = {}
            for q, n in enumerate(names):
                if q < 6:
                    m = [0.0] * 6; m[q] = 1.0
                    fwd[n], _ = T._forward(td / f"f{n}", mij=m)
                else:
                    f = [0.0] * 3; f[q - 6] = 1.0
                    fwd[n], _ = T._forward(td / f"f{n}", f=f)
            out = {}
            for c, cmp in enumerate("xyz"):
                gf, _ = T._reciprocal(td / f"r{cmp}", cmp)
                for q, n in enumerate(names):
                    a, b = fwd[n][c].astype(np.float64), gf[q].astype(np.float64)
                    if q < 6 and half:
                        a = 0.5 * (a[:-1] + a[1:]); b = b[:-1]
                    if np.abs(a).max() > 0:
                        out[f"U{cmp}<-{n}"] = round(float(rel_l2(b, a)), 5)
            print(label, json.dumps(out), "worst", max(out.values()), flush=True)
    finally:
        Swpc3d.__init__ = orig
        T.COMMON.clear(); T.COMMON.update(old)

run2("big256_nt400", nx=256, ny=256, nz=160, nt=400)
run2("big256_nt400_elastic", nm=0, nx=256, ny=256, nz=160, nt=400)
run2("big256_nt400_deepst", nx=256, ny=256, nz=160, nt=400, stations=[f"{T.ST[0]} {T.ST[1]} 6.0 st01 dep"])
