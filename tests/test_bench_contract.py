"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the keys the
driver reads (on rank 0 only), and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_json_line():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must still use every core it is allowed to
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], env={"OMP_NUM_THREADS": "1", "SWPC_BENCH_CPU_SAMPLE": "160,160,160"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_s" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0)) and "160x160x160" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert p.returncode == 0 and not p.stdout.strip()


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_cell_counts_and_core_region_arithmetic():
    """The algorithmic-byte bookkeeping of the roofline: interior / absorber cells of a rank (m_global.f90:334-376) and of the core
    region the sweep stopwatches cover when the exchange is overlapped (every owned column but the slab towards each neighbour)."""
    sys.path.insert(0, str(ROOT))
    import bench

    # rank 1 of a 2 x 1 decomposition of 1024 x 1024 x 1024, na = 20: columns 513..1024, interior kernel box 513..1004 x 21..1004 x 1..1004
    run = {"nz": 1024, "nxp": 512, "nyp": 1024, "ibeg": 513, "jbeg": 1, "ibeg_k": 513, "iend_k": 1004, "jbeg_k": 21, "jend_k": 1004,
           "kbeg_k": 1, "kend_k": 1004, "nproc_x": 2, "nproc_y": 1, "myid": 1}
    interior, pml = bench.cell_counts(run)
    assert interior == 492 * 984 * 1004 and interior + pml == 512 * 1024 * 1024
    li0, li1, lj0, lj1 = bench.core_region(run, 2)
    assert (li0, li1, lj0, lj1) == (8, 511, 0, 1023)                      # one tile column (8) towards the -x neighbour, nothing else
    ci, cp = bench.cell_counts(run, (li0, li1, lj0, lj1))
    assert ci == (492 - 8) * 984 * 1004 and ci + cp == 504 * 1024 * 1024
    assert bench.core_region(run, 1) == (0, 511, 0, 1023)
    b = bench.bytes_per_cell(3, 8)
    assert (b["stress_interior"], b["stress_pml"], b["vel_interior"], b["vel_pml"]) == (280, 200, 100, 172)
    assert bench.psv_bytes_per_cell(3, 8)["stress_interior"] == 152 and bench.bytes_per_cell(0, 8)["stress_interior"] == 128
