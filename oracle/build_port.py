#!/usr/bin/env python3
"""Compile the plain-C restatement (oracle/port/*.c) into oracle/_build/liboracle_port.so.  TEST INFRASTRUCTURE ONLY."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "liboracle_port.so")
    srcs = sorted(glob.glob(os.path.join(HERE, "port", "*.c")))
    deps = srcs + glob.glob(os.path.join(HERE, "port", "*.h"))
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    cmd = ["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-Wno-misleading-indentation", "-Wno-unused-function", "-o", lib] + srcs + ["-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle port build failed")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
