"""Build the C restatement of the router (oracle/route_oracle.c) into oracle/_build/.

Test infrastructure only (see the header of route_oracle.c).  Called by
``__graft_entry__.build()`` and lazily by ``oracle.route_oracle_c`` when the .so is stale.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "route_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libroute_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    newest = max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(HERE, "exp_fast.h")))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    # -ffp-contract=off: every rounding point in the source is a rounding point in the binary;
    # fmaf() calls still compile to hardware FMA with -mfma (and are exact in libm otherwise).
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
