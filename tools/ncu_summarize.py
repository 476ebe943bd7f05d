#!/usr/bin/env python
"""Turns ncu outputs into the small text/JSON summaries committed under profiles/.

  python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/launches_rNN.json
      per-kernel count / total / share of an `ncu --metrics gpu__time_duration.sum --csv` launch list
  python tools/ncu_summarize.py rep gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/ncu_rNN.md
      key metrics of `ncu --set full` captures (needs the ncu CLI; works without a GPU)
  python tools/ncu_summarize.py traffic profiles/ncu_rNN_traffic.json gpurun_out/ncu_rNN_*.ncu-rep
      DRAM bytes per launch next to the algorithmic bytes of the measurement shapes (bench.py's roofline.traffic)
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


# measurement shapes of tools/ncu_targets.py: target -> (description, algorithmic bytes per launch)
_M, _N, _K = 119808, 2048, 512
SHAPES = {
    "gemm_fwd": (f"fwd M={_M} N={_N} K={_K} bias+ReLU bf16", 2 * (_M * _K + _N * _K + _M * _N)),
    "gemm_dgrad": (f"dgrad M={_M} N={_K} K={_N} bf16", 2 * (_M * _N + _N * _K + _M * _K)),
    "gemm_dgrad_mask": (f"dgrad + ReLU mask M={_M} N={_N} K={_K} bf16 (side operand [M, N] bf16)",
                        2 * (_M * _K + _N * _K + 2 * _M * _N)),
    "gemm_wgrad": (f"wgrad M={_N} N={_K} K={_M} fp32 accumulate", 2 * (_M * _N + _M * _K) + 8 * _N * _K),
    "gemm_fwd_bits": (f"fwd M={_M} N={_N} K={_K} bias+ReLU+bit record bf16", 2 * (_M * _K + _N * _K + _M * _N) + _M * _N // 8),
    "gemm_dgrad_bits": (f"dgrad masked by the bit record M={_M} N={_N} K={_K} bf16 (side operand [M, N / 32] words)",
                        2 * (_M * _K + _N * _K + _M * _N) + _M * _N // 8),
    "gemm_res": (f"fwd M={_M} N={_K} K={_K} bias+residual bf16", 2 * (3 * _M * _K + _K * _K)),
    "gemm_x3": (f"fwd M={_M} N={_N} K=3x{_K} split-bf16 operands, fp32 out", 2 * 3 * (_M * _K + _N * _K) + 4 * _M * _N),
    "attn": ("B=1024 S=117 H=8 bf16 (fwd: Q K V O; bwd: Q K V dO dQ dK dV)", None),
    "split": (f"fp32 [{_M}, {_K}] -> bf16 [{_M}, 3x{_K}]", _M * _K * (4 + 6)),
    "ln": (f"LayerNorm bwd rows={_M} D=512 bf16 (dy, x in; dx out)", 3 * 2 * _M * 512),
    "gae": ("T=128 N=65536 both streams", 36 * 128 * 65536),
    "loss": ("R=8388608 A=20", (8 * 20 + 44) * 128 * 65536),
    "adam": ("62.9M parameters, fp32 master + bf16 shadow + grad zeroing", (62_900_000 // 64 * 64) * 34),
}


def traffic(out, paths):
    """{target: kernel, DRAM bytes, duration} of `ncu --set full` captures named ncu_<round>_<target>_<kernel>.ncu-rep."""
    res = OrderedDict()
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        r = rows[2]

        def val(k):
            v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3,
                        "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
        m = re.search(r"ncu_r\d+_(.+?)_(svla_gemm_tc|attn_tc_fwd|attn_tc_bwd|attn_ws_fwd|attn_ws_bwd_x3|attn_ws_bwd|"
                      r"split_concat|gae_march|ppo_lag|clip_adam|layernorm_bwd)", p)
        tgt = m.group(1) if m else p
        key = tgt if tgt.startswith("gemm") or tgt in ("gae", "loss", "adam", "split", "ln") else \
            tgt + ":" + short(r[idx["Kernel Name"]])
        e = {"kernel": short(r[idx["Kernel Name"]]),
             "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
             "ncu_us": val("gpu__time_duration.sum"),
             "tensor_pipe_active_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]])}
        if tgt.startswith("attn"):
            bs, el = 1024 * 117 * 512, (4 if tgt == "attn_x3" else 2)
            n_t = 4 if "fwd" in e["kernel"] else 7
            e["shape"] = f"B=1024 S=117 H=8 {'fp32 (hi, lo) operands' if tgt == 'attn_x3' else 'bf16'}"
            # split mode: operands arrive as two bf16 tensors each (4 B / element), outputs are fp32
            e["algorithmic_bytes"] = n_t * bs * el
            e["traffic_over_algorithmic"] = round(e["dram_bytes"] / e["algorithmic_bytes"], 3)
        elif tgt in SHAPES and SHAPES[tgt][1]:
            e["shape"], e["algorithmic_bytes"] = SHAPES[tgt]
            e["traffic_over_algorithmic"] = round(e["dram_bytes"] / e["algorithmic_bytes"], 3)
        if e.get("algorithmic_bytes"):
            e["algorithmic_gbs"] = round(e["algorithmic_bytes"] / e["ncu_us"] / 1e3, 1)
        res[key] = e
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3:])
    else:
        rep(sys.argv[2:])
