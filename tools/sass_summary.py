#!/usr/bin/env python
"""Writes profiles/sass_r02.md: for every object of libsafevla_b200 the count of Blackwell tensor-core / TMA SASS
mnemonics (cuobjdump -sass) per kernel, plus a short excerpt -- the static proof that the hot kernels are tcgen05 / TMA
code (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops).  Run after `python -m safevla_b200.build`."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA")


def main():
    out = ["# SASS mnemonic census (round 2)", "",
           "`cuobjdump -sass safevla_b200/build/*.o`, sm_100a; counts per kernel (only kernels with tensor-core / TMA "
           "instructions are listed; FFMA shown for contrast).", ""]
    for obj in sorted(glob.glob(os.path.join(ROOT, "safevla_b200", "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernels, cur = {}, None
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
                kernels[cur] = {p: 0 for p in PAT}
                kernels[cur]["_ex"] = []
                continue
            if cur is None:
                continue
            for p in PAT:
                if re.search(r"\b" + p + r"\b|\b" + p + r"\.", ln):
                    kernels[cur][p] += 1
                    if p in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM") and len(kernels[cur]["_ex"]) < 4 and \
                            not any(p in e for e in kernels[cur]["_ex"]):
                        kernels[cur]["_ex"].append(re.sub(r"\s+", " ", ln.split("*/")[1] if "*/" in ln else ln).strip()[:110])
        rows = [(k, v) for k, v in kernels.items() if v["UTCHMMA"] or v["UTMALDG"] or v["UTMASTG"] or v["LDTM"]]
        if not rows:
            continue
        out += [f"## {os.path.basename(obj)}", "", "| kernel | " + " | ".join(PAT) + " |", "|---|" + "---|" * len(PAT)]
        for k, v in rows:
            out.append(f"| `{k[:90]}` | " + " | ".join(str(v[p]) for p in PAT) + " |")
        out.append("")
        k, v = max(rows, key=lambda kv: kv[1]["UTCHMMA"])
        out += ["excerpt (`" + k[:80] + "`):", "```"] + v["_ex"] + ["```", ""]
    dst = os.path.join(ROOT, "profiles", "sass_r02.md")
    open(dst, "w").write("\n".join(out) + "\n")
    print(dst, len(out), "lines")


if __name__ == "__main__":
    sys.exit(main())
