"""Compile the C oracle (oracle/gs_oracle.c) into oracle/libgs_oracle.so.

TEST INFRASTRUCTURE -- see the header of gs_oracle.c.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gs_oracle.c")
LIB = os.path.join(HERE, "libgs_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = [
        "gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-mfma", "-mavx2", "-fno-math-errno",
        "-shared", "-fPIC", "-o", LIB, SRC, "-lm",
    ]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
