"""Random-media volumes for the `*_rmed` velocity models (vmodel_uni_rmed / lhm_rmed / lgm_rmed).

The reference generates them with its offline tool src/tools/gen_rmed3d.f90 (netCDF classic file: dimensions x, y, z;
variables x, y, z and, 4th, the volume with x fastest -- which is how src/shared/m_rdrmed.f90:73-134 finds it).  This
module writes the same container from a numpy array so that synthetic heterogeneous models can be built without the
Fortran tool or a netCDF library; the spectrum of the field is the caller's business.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np


def write_rmed3d(path, xi: np.ndarray, dx: float = 0.5, dy: float | None = None, dz: float | None = None, title: str = "random media") -> None:
    """xi: (nz, ny, nx) float array of fractional velocity perturbations."""
    nz, ny, nx = xi.shape
    dy = dx if dy is None else dy
    dz = dx if dz is None else dz

    def name(s: str) -> bytes:
        b = s.encode()
        return struct.pack(">I", len(b)) + b + b"\0" * (-len(b) % 4)

    dims = struct.pack(">II", 0x0A, 3) + b"".join(name(n) + struct.pack(">I", m) for n, m in (("x", nx), ("y", ny), ("z", nz)))
    tb = title.encode()
    gatts = struct.pack(">II", 0x0C, 1) + name("title") + struct.pack(">II", 2, len(tb)) + tb + b"\0" * (-len(tb) % 4)
    shapes = [("x", [0], nx), ("y", [1], ny), ("z", [2], nz), ("random media", [2, 1, 0], nx * ny * nz)]

    def var_list(begins) -> bytes:
        out = struct.pack(">II", 0x0B, len(shapes))
        for (n, dimids, cnt), beg in zip(shapes, begins):
            out += name(n) + struct.pack(">I", len(dimids)) + b"".join(struct.pack(">I", d) for d in dimids)
            out += struct.pack(">II", 0, 0) + struct.pack(">II", 5, (cnt * 4) & 0xFFFFFFFF) + struct.pack(">Q", beg)
        return out

    head = b"CDF\x02" + struct.pack(">I", 0) + dims + gatts   # CDF-2: 64-bit offsets
    off = len(head) + len(var_list([0] * 4))
    begins = []
    for _, _, cnt in shapes:
        begins.append(off)
        off += cnt * 4
    with open(Path(path), "wb") as f:
        f.write(head + var_list(begins))
        for n, d in ((nx, dx), (ny, dy), (nz, dz)):
            f.write((np.arange(n) * d).astype(">f4").tobytes())
        f.write(np.ascontiguousarray(xi).astype(">f4").tobytes())


def smoothed_gaussian(shape, sigma_cells: float, epsilon: float, seed: int) -> np.ndarray:
    """A periodic, Gaussian-smoothed white-noise field with standard deviation `epsilon` (shape (nz, ny, nx))."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(shape)
    k = [np.fft.fftfreq(n) for n in shape]
    kk = k[0][:, None, None] ** 2 + k[1][None, :, None] ** 2 + k[2][None, None, :] ** 2
    f = np.fft.ifftn(np.fft.fftn(w) * np.exp(-2.0 * (np.pi * sigma_cells) ** 2 * kk)).real
    return (f * (epsilon / f.std())).astype(np.float32)
