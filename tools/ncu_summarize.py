#!/usr/bin/env python
"""Turns ncu outputs into the small text/JSON summaries committed under profiles/.

  python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/launches_rNN.json
      per-kernel count / total / share of an `ncu --metrics gpu__time_duration.sum --csv` launch list
  python tools/ncu_summarize.py rep gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/ncu_rNN.md
      key metrics of `ncu --set full` captures (needs the ncu CLI; works without a GPU)
"""
from __future__ import annotations

import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__cycles_active.avg", "smsp__inst_executed.sum",
]


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    return re.sub(r"\(.*$", "", name).strip()


def launches(path: str, out: str):
    rows = [ln for ln in open(path, errors="replace") if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg: "OrderedDict[str, list]" = OrderedDict()
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    ks = sorted(agg.items(), key=lambda kv: -kv[1][1])
    res = {"source": path, "launches": n, "total_us": total,
           "note": "ncu per-launch durations are cold-cache and serialised: compare shares, not absolutes",
           "kernels": [{"kernel": k, "launches": c, "total_us": round(t, 1), "share": round(t / total, 4),
                        "avg_us": round(t / c, 2)} for k, (c, t) in ks]}
    json.dump(res, open(out, "w"), indent=1)
    for k in res["kernels"][:25]:
        print(f"{k['share']*100:6.2f} %  {k['launches']:6d} x {k['avg_us']:9.2f} us  {k['kernel']}")
    print(f"{n} launches, {total/1e3:.2f} ms")


def rep(paths):
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            print(f"## {p}\n(no data)\n")
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"## {p}\n")
        for r in rows[2:]:
            print(f"### `{short(r[idx['Kernel Name']])}`  grid {r[idx.get('Grid Size', 0)]} block {r[idx.get('Block Size', 0)]}\n")
            print("| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in idx:
                    print(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        rep(sys.argv[2:])
