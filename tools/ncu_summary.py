#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of one C3 step into profiles/ (per kernel: time, DRAM bytes, registers,
occupancy, issue rate, top stall reasons) and refresh profiles/ncu_traffic.json (DRAM bytes per launch and pass,
read by bench.py for roofline.traffic).
  python tools/ncu_summary.py gpurun_out/r01c_full.ncu-rep profiles/r01_ncu_full_summary.json"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}


def f(r, name, default=None):
    if name not in col or r[col[name]] in ("", "n/a"):
        return default
    return float(r[col[name]].replace(",", ""))


def pass_of(name, grid):
    g = int(grid.strip("()").split(",")[0])
    if "psf_z_pruned" in name:
        return "psf_z"
    if "x_fwd_kernel<1" in name:
        return "psf_x"
    if "xrow_fwd" in name or "xrowg_fwd" in name or "x_fwd_kernel<0" in name:
        return "x_fwd"
    if "xrow_inv" in name or "xrowg_inv" in name or "x_inv_kernel" in name:
        return "x_inv"
    if "col_tma_kernel<0" in name or "col_tma_kernel<(int)0" in name:
        return "psf_y" if g < 100 else "y_fwd"
    if "col_tma_kernel<1" in name or "col_tma_kernel<(int)1" in name:
        return "y_inv"
    if "col_tma_kernel<2" in name or "col_tma_kernel<(int)2" in name:
        return "z_fused"
    if "col_tma_kernel<3" in name or "col_tma_kernel<(int)3" in name:
        return "z_fused_otf"
    if "col_pipe_kernel<0" in name:
        return "y_fwd"
    if "col_pipe_kernel<1" in name:
        return "y_inv"
    if "col_static_kernel<2" in name or "col_otf" in name:
        return "z_fused"
    if "col_static_kernel<1" in name:
        return "y_inv"
    if "col_static_kernel<0" in name:
        return "psf_y" if g < 1000 else "y_fwd"
    return None


stalls = [c for c in hdr if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
summary, traffic = [], {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    p = pass_of(name, r[col["Grid Size"]])
    if p is None or p in traffic:
        continue
    rd, wr = f(r, "dram__bytes_read.sum", 0.0), f(r, "dram__bytes_write.sum", 0.0)
    unit_r = rows[1][col["dram__bytes_read.sum"]]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_r, 1)
    rd, wr = rd * mult, wr * mult
    us = f(r, "gpu__time_duration.sum", 0.0)
    tu = rows[1][col["gpu__time_duration.sum"]]
    us *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(tu, 1.0)
    top = sorted(((f(r, c, 0.0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stalls),
                 reverse=True)[:3]
    traffic[p] = int(rd + wr)
    summary.append({
        "kernel": p, "name": name[:90], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]], "us": round(us, 1),
        "dram_read_MB": round(rd / 1e6, 1), "dram_write_MB": round(wr / 1e6, 1),
        "dram_GBps": round((rd + wr) / (us * 1e-6) / 1e9, 0) if us else None,
        "regs": f(r, "launch__registers_per_thread"),
        "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_pct_of_nominal": f(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        "smem_bank_conflicts": f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "top_stalls_warps_per_issue": [[n, round(v, 2)] for v, n in top],
    })
json.dump(summary, open(out, "w"), indent=1)
tpath = out.rsplit("/", 1)[0] + "/ncu_traffic.json"
try:
    merged = json.load(open(tpath))     # keep the passes of earlier captures (e.g. the on-the-fly fused pass)
except Exception:
    merged = {}
merged.update(traffic)
json.dump(merged, open(tpath, "w"), indent=1)
print(json.dumps(traffic))
for s in summary:
    print(s["kernel"], s["us"], s["dram_GBps"], s["regs"], s["warps_active_pct"], s["issue_active_pct"], s["top_stalls_warps_per_issue"])
