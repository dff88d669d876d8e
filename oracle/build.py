"""Compile the C restatement (oracle/ma_oracle.c, oracle/pk_oracle.c) into oracle/liboracle.so.

TEST INFRASTRUCTURE ONLY.  Plain gcc, no fast-math: the restatement spells out every
rounding the reference performs; FP contraction is left at gcc's default so the float32
accumulate `number += w` may fuse its last multiply like the reference binary does.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")
SRCS = [os.path.join(HERE, "ma_oracle.c"), os.path.join(HERE, "pk_oracle.c")]


def build(force=False):
    if not force and os.path.exists(SO) and all(
            os.path.getmtime(SO) >= os.path.getmtime(s) for s in SRCS):
        return SO
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-o", SO] + SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True))
