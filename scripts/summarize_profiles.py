"""Turn the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/."""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
F = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
TF = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def launch_list(src: Path, dst: Path, header: str, key: str):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid, iu = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(r[iid], {"k": r[ik]})
        d[r[im]] = (float(r[iv].replace(",", "")), r[iu])
    out = []
    for i, d in L.items():
        t, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        out.append((int(i), d["k"].split("(")[0].replace("void ", ""), t[0] * TF[t[1]], rd[0] * F[rd[1]], wr[0] * F[wr[1]]))
    with open(dst, "w") as fo:
        fo.write(header)
        fo.write("id,kernel,ms,dram_read_GB,dram_write_GB\n")
        for o in out:
            fo.write(f"{o[0]},{o[1]},{o[2]:.4f},{o[3] / 1e9:.3f},{o[4] / 1e9:.3f}\n")
    # one steady-state step = everything between two consecutive wav_store / the last stress_tma launches
    idx = [n for n, o in enumerate(out) if o[1].startswith("stress_tma") or ("sweep_direct" in o[1] and ", 1>" in o[1] and not any("stress_tma" in x[1] for x in out))]
    tma = [n for n, o in enumerate(out) if o[1].startswith("stress_tma")]
    if len(tma) >= 2:
        step = out[tma[-2]:tma[-1]]
    else:
        step = out[-4:]
    tot = sum(o[2] for o in step)
    share = collections.OrderedDict()
    for o in step:
        share[o[1]] = share.get(o[1], 0.0) + o[2]
    stress = [o for o in step if o[1].startswith("stress_tma") or ("sweep_direct" in o[1] and ", 1>" in o[1])]
    vel = [o for o in step if "sweep_direct" in o[1] and ", 0>" in o[1] or o[1].startswith("vel_tma") or o[1].startswith("vel_ring")]
    res = {"stress_dram_bytes_per_launch": sum(o[3] + o[4] for o in stress), "vel_dram_bytes_per_launch": sum(o[3] + o[4] for o in vel),
           "stress_ms_under_ncu": sum(o[2] for o in stress), "vel_ms_under_ncu": sum(o[2] for o in vel),
           "step_share": {k: round(v / tot, 4) for k, v in share.items()}, "source": str(dst.relative_to(ROOT)),
           "note": "the fused stress sweep = stress_tma (interior tiles) + 5 sweep_direct launches (absorber shell); serialised under ncu"}
    tj = OUT / "traffic.json"
    cur = json.loads(tj.read_text()) if tj.exists() else {}
    cur[key] = res
    tj.write_text(json.dumps(cur, indent=1))
    print(json.dumps(res, indent=1))


def launch_list_r02(src: Path, dst: Path, header: str, key: str):
    """round 2: the fused stress sweep = stress_tma + pml_tma<., 1, ..>, the velocity sweep = vel_ring / vel_ring2 + pml_tma<., 0, ..>"""
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid, iu = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(r[iid], {"k": r[ik]})
        d[r[im]] = (float(r[iv].replace(",", "")), r[iu])
    out = []
    for i, d in L.items():
        t, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        out.append((int(i), d["k"].split("(")[0].replace("void ", ""), t[0] * TF[t[1]], rd[0] * F[rd[1]], wr[0] * F[wr[1]]))
    with open(dst, "w") as fo:
        fo.write(header)
        fo.write("id,kernel,ms,dram_read_GB,dram_write_GB\n")
        for o in out:
            fo.write(f"{o[0]},{o[1]},{o[2]:.4f},{o[3] / 1e9:.3f},{o[4] / 1e9:.3f}\n")
    tma = [n for n, o in enumerate(out) if o[1].startswith("stress_tma")]
    step = out[tma[-2]:tma[-1]] if len(tma) >= 2 else out
    tot = sum(o[2] for o in step)
    share = collections.OrderedDict()
    for o in step:
        share[o[1]] = share.get(o[1], 0.0) + o[2]
    is_stress = lambda n: n.startswith("stress_tma") or (n.startswith("pml_tma") and ", 1, " in n) or ("sweep_direct" in n and ", 1>" in n)
    is_vel = lambda n: n.startswith("vel_ring") or n.startswith("vel_tma") or (n.startswith("pml_tma") and ", 0, " in n) or ("sweep_direct" in n and ", 0>" in n)
    stress = [o for o in step if is_stress(o[1])]
    vel = [o for o in step if is_vel(o[1])]
    res = {"stress_dram_bytes_per_launch": sum(o[3] + o[4] for o in stress), "vel_dram_bytes_per_launch": sum(o[3] + o[4] for o in vel),
           "stress_ms_under_ncu": sum(o[2] for o in stress), "vel_ms_under_ncu": sum(o[2] for o in vel),
           "step_share": {k: round(v / tot, 4) for k, v in share.items()}, "source": str(dst.relative_to(ROOT)),
           "note": "the fused stress sweep = stress_tma (interior tiles) + pml_tma<F,1,..> launches (absorber shell); serialised under ncu"}
    tj = OUT / "traffic.json"
    cur = json.loads(tj.read_text()) if tj.exists() else {}
    cur[key] = res
    tj.write_text(json.dumps(cur, indent=1))
    print(key, json.dumps(res, indent=1))


def full_summary(rep: Path, dst: Path, header: str):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
            "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    with open(dst, "w") as fo:
        fo.write(header + "\n")
        for r in rows[2:]:
            fo.write(r[hdr.index("Kernel Name")] + "\n")
            for w in want:
                if w in hdr:
                    i = hdr.index(w)
                    fo.write(f"  {w:84s} {r[i]:>22s} {units[i]}\n")
            fo.write("\n")
    print(open(dst).read()[:600])


def launch_list_psv(src: Path, dst: Path, header: str):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid, iu = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(r[iid], {"k": r[ik]})
        d[r[im]] = (float(r[iv].replace(",", "")), r[iu])
    with open(dst, "w") as fo:
        fo.write(header)
        fo.write("id,kernel,ms,dram_read_GB,dram_write_GB\n")
        for i, d in L.items():
            t, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
            fo.write(f"{i},{d['k'].split('(')[0].replace('void ', '')},{t[0] * TF[t[1]]:.4f},{rd[0] * F[rd[1]] / 1e9:.3f},{wr[0] * F[wr[1]] / 1e9:.3f}\n")
    print(open(dst).read()[:900])


if __name__ == "__main__":
    go = ROOT / "gpurun_out"
    if len(sys.argv) >= 5 and sys.argv[1] == "r02":   # r02 <raw ncu csv> <profiles/ name> <traffic.json key> [header line ...]
        launch_list_r02(Path(sys.argv[2]), OUT / sys.argv[3], "".join(f"# {x}\n" for x in sys.argv[5:]), sys.argv[4])
        sys.exit(0)
    if (go / "launches_psv_final.csv").exists():
        launch_list_psv(go / "launches_psv_final.csv", OUT / "r01_launches_psv_16384x8192.csv",
                        "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 python scripts/psv_probe.py 16384,8192 3\n"
                        "# B200, swpc_psv 16384x8192 (nx x nz), NM=3, PML na=20, f64 fields, final code of round 1; psv_sweep<F,NM,1> = stress sweep, "
                        "psv_sweep<F,NM,0> = velocity sweep; per-launch times are cold-cache and serialised\n")
    if (go / "prof_psv_final.ncu-rep").exists():
        full_summary(go / "prof_psv_final.ncu-rep", OUT / "r01_ncu_full_psv_16384x8192.txt",
                     "ncu --set full --clock-control none --import-source on -k regex:psv_sweep -s 12 -c 2 python scripts/psv_probe.py 16384,8192 3\n"
                     "B200 (sm_100a), swpc_psv 16384x8192, NM=3, PML na=20, f64 fields, final code of round 1; one stress and one velocity sweep "
                     "(thread = one k, block = 256 k, marching along i with L2 prefetch of column i+1; velocity operands fetched up front, PsvVelOps)")
    if (go / "launches_r01b.csv").exists():
        launch_list(go / "launches_r01b.csv", OUT / "r01_launches_bench_1024x1024x512_tma.csv",
                    "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 python bench.py --steps 3 --warmup 3 --no-cpu-baseline\n"
                    "# B200, 1024x1024x512 NM=3 PML f64 fields, TMA interior stress kernel + direct shell/velocity kernels; per-launch times are cold-cache and SERIALISED\n"
                    "# (the shell boxes overlap the interior kernel in a real run): compare SHARES\n", "1024x1024x512_f64_nm3")
    if (go / "prof_r01_stress_tma_512x512x256.ncu-rep").exists():
        full_summary(go / "prof_r01_stress_tma_512x512x256.ncu-rep", OUT / "r01_ncu_full_stress_tma_512x512x256.txt",
                     "ncu --set full --clock-control none --import-source on -k regex:stress_tma -s 3 -c 1 python bench.py --grid 512,512,256 --steps 2 --warmup 3\n"
                     "B200 (sm_100a), 512x512x256, NM=3, PML na=20, f64 fields; stress_tma handles the interior tiles (472 x 472 columns, k <= 256 masked to k <= 236): "
                     "tiles 32k x 8i, 32 planes per block, 16 consumer warps + 1 producer warp, 188 KB dynamic shared memory")
