"""Compile the C oracle (oracle/raster_oracle.c) with gcc. Test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libgsd_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "raster_oracle.c")
    if (not force) and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", LIB, src, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
