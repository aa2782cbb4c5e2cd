#!/usr/bin/env python3
"""Summarise an ncu report (`ncu --set full ... -o X`) as a markdown table + the per-kernel DRAM traffic JSON that
bench.py reads (profiles/ncu_traffic.json).   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_rXX [--traffic]
"""
import csv
import io
import json
import subprocess
import sys

COLS = [
    ("time us", "gpu__time_duration.sum"),
    ("DRAM rd MB", "dram__bytes_read.sum"),
    ("DRAM wr MB", "dram__bytes_write.sum"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("XU %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("FP64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("tensor %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor(any) %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warp inst M", "smsp__inst_executed.sum"),
    ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def to_unit(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "inst": 1e-6}
    return v * scale.get(unit, 1.0)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith(STALL) and "not_issued" not in h]
    agg = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("kb::", "")
        a = agg.setdefault(name, {"n": 0, "vals": {}, "stall": {}})
        a["n"] += 1
        for label, key in COLS:
            if key in idx and r[idx[key]] not in ("", "n/a"):
                a["vals"][label] = a["vals"].get(label, 0.0) + to_unit(r[idx[key]], units[idx[key]], label)
        for k in stalls:
            try:
                a["stall"][k[len(STALL):]] = a["stall"].get(k[len(STALL):], 0.0) + float(r[idx[k]])
            except ValueError:
                pass
    lines = ["| kernel | launches | " + " | ".join(l for l, _ in COLS) + " | top stalls |", "|---|---|" + "---|" * (len(COLS) + 1)]
    traffic = {}
    for name, a in agg.items():
        n = a["n"]
        cells = []
        for label, _ in COLS:
            v = a["vals"].get(label)
            cells.append("" if v is None else (f"{v / n:.1f}" if v / n < 1000 else f"{v / n:.0f}"))
        tot = sum(a["stall"].values()) or 1.0
        top = sorted(a["stall"].items(), key=lambda kv: -kv[1])[:3]
        lines.append(f"| {name} | {n} | " + " | ".join(cells) + " | " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top) + " |")
        traffic[name.split("<")[0]] = int((a["vals"].get("DRAM rd MB", 0) + a["vals"].get("DRAM wr MB", 0)) / n * 1e6)
    with open(out + "_summary.md", "w") as f:
        f.write(f"ncu --set full --clock-control none, per-launch averages; source: {rep}\n\n" + "\n".join(lines) + "\n")
    if "--traffic" in sys.argv:
        # poses per launch of the captured command (bench.py scales the figure to its own launches): --poses K, default 30 (bench.py: 150 poses in 5 launches)
        traffic["_poses_per_launch"] = int(sys.argv[sys.argv.index("--poses") + 1]) if "--poses" in sys.argv else 30
        traffic["_source"] = rep
        with open(out.rsplit("/", 1)[0] + "/ncu_traffic.json", "w") as f:
            json.dump(traffic, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
