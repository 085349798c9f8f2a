#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list:
device-time share per kernel and, when the DRAM counters are present, bytes per launch.

    python scripts/ncu_launch_summary.py gpurun_out/launches.csv [--json out.json]
"""
import csv
import hashlib
import io
import os
import json
import re
import sys
from collections import defaultdict


def to_base(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "second": 1e6, "nsecond": 1e-3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * scale.get(unit, 1.0)


def gemm_sources_sha():
    """Hash of the sources the GEMM kernel is compiled from: bench.py only quotes the DRAM traffic of a capture whose
    hash equals the tree's (a changed kernel nulls the number instead of silently keeping a stale one). Run this script
    on a capture BEFORE touching those sources again: the stamp is taken from the tree at summary time."""
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "touhouimageclassification_b200", "csrc")
    h = hashlib.sha256()
    for f in ("gemm_tcgen05.cu", "tic_common.cuh"):   # the device code of the GEMM (tic_internal.cuh holds prototypes only)
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def main():
    path = sys.argv[1]
    text = open(path, errors="replace").read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = defaultdict(lambda: defaultdict(float))
    launches = defaultdict(set)
    for r in rows:
        name = re.sub(r"^void\s+", "", r["Kernel Name"])
        name = re.sub(r"tic::\(anonymous namespace\)::|tic::|<unnamed>::", "", name)
        name = re.sub(r"\(.*$", "", name)
        per[name][r["Metric Name"]] += to_base(r["Metric Value"], r["Metric Unit"])
        launches[name].add(r["ID"])
    total_us = sum(v["gpu__time_duration.sum"] for v in per.values())
    out = {"launches": sum(len(v) for v in launches.values()), "total_ms": total_us / 1e3, "kernels": {}}
    print(f"{out['launches']} launches, {total_us / 1e3:.1f} ms device time (serialised, cold-cache: compare shares)")
    gemm = defaultdict(float)
    for name, m in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(launches[name])
        t = m["gpu__time_duration.sum"]
        rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
        out["kernels"][name] = dict(launches=n, us=t, share=t / total_us, dram_mb_per_launch=(rd + wr) / n / 1e6)
        line = f"{100 * t / total_us:5.1f}%  {t:10.1f} us  {n:4d}  {name}"
        if rd + wr > 0:
            line += f"   DRAM {(rd + wr) / n / 1e6:8.1f} MB/launch, {(rd + wr) / t / 1e6:5.2f} TB/s"
        print(line)
        if name.startswith("gemm_bf16_tcgen05_kernel"):
            gemm["n"] += n; gemm["us"] += t; gemm["bytes"] += rd + wr
    if gemm["n"]:
        out["gemm"] = dict(launches=int(gemm["n"]), share=gemm["us"] / total_us, us_per_launch=gemm["us"] / gemm["n"],
                           dram_bytes_per_launch=gemm["bytes"] / gemm["n"])
        print(f"\nall GEMM launches: {100 * gemm['us'] / total_us:.1f}% of the device time, "
              f"{gemm['bytes'] / gemm['n'] / 1e6:.1f} MB of DRAM traffic per launch")
    out["gemm_sources_sha"] = gemm_sources_sha()
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
