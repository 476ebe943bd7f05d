"""TEST INFRASTRUCTURE ONLY.  Compiles oracle/c/*.c (plain C restatements) into oracle/c/_build/liboracle_c.so with gcc;
`__graft_entry__.build()` calls this so the checker exists wherever the tests run."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle_c.so")


def build(force: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(HERE, "*.c")))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           "-o", OUT, *srcs])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
