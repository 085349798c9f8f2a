#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of figures DESIGN.md quotes per kernel launch.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python scripts/ncu_raw_summary.py [--json out.json]
"""
import csv
import json
import re
import sys

KEYS = {
    "time_us": "gpu__time_duration.sum",
    "dram_read_mb": "dram__bytes_read.sum",
    "dram_write_mb": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "smem_dyn_kb": "launch__shared_mem_per_block_dynamic",
    "sm_cycles": "sm__cycles_elapsed.avg",
    "sm_mhz": "sm__cycles_elapsed.avg.per_second",
}


def main():
    rows = list(csv.reader(sys.stdin))
    hdr = next(r for r in rows if "Kernel Name" in r)
    units = rows[rows.index(hdr) + 1]
    out = []
    for r in rows[rows.index(hdr) + 2:]:
        if len(r) != len(hdr):
            continue
        name = re.sub(r"\(.*$", "", re.sub(r"^.*::", "", r[hdr.index("Kernel Name")].split("(")[0]))
        d = {"kernel": name}
        for k, m in KEYS.items():
            if m in hdr:
                v = r[hdr.index(m)].replace(",", "")
                try:
                    v = float(v)
                    u = units[hdr.index(m)]
                    if k == "time_us" and u in ("ms", "msecond"):
                        v *= 1e3
                    if k == "time_us" and u in ("ns", "nsecond"):
                        v /= 1e3
                    if k.endswith("_mb") and u == "Gbyte":
                        v *= 1e3
                    if k.endswith("_mb") and u == "Kbyte":
                        v /= 1e3
                    if k.endswith("_mb") and u == "byte":
                        v /= 1e6
                    if k == "sm_mhz":
                        v = v * {"Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6}.get(u, 1.0)
                    d[k] = round(v, 3)
                except ValueError:
                    pass
        stalls = {}
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h) or \
                re.match(r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
            if m:
                try:
                    stalls[m.group(1)] = round(float(r[i]), 2)
                except ValueError:
                    pass
        d["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out.append(d)
    for d in out:
        print(json.dumps(d))
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
