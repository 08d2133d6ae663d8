#!/usr/bin/env python3
"""Build kernel-experiment variants of libspectral_b200.so into scratch/var_<name>.so.

    python tools/variants.py [--tu tu_feat2.cu] name1="-DkFoo=1 -DkBar=2" name2="..."

Only the named translation unit is recompiled with the extra flags; the other objects come from the main in-tree
build (run build.py first).  Select a variant at run time with SB200_LIB=scratch/var_<name>.so (see _lib.py)."""
import importlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = importlib.import_module("transtacos-retunegan_b200.build")


def one(arg, tu):
    name, flags = arg.split("=", 1)
    out = os.path.join(ROOT, "scratch", f"var_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    obj = os.path.join(B.OBJ_DIR, f"var_{name}.{os.path.splitext(tu)[0]}.o")
    base = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a"]
    cmd = base + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-diag-suppress", "550,177", "-c", "-o", obj,
                  os.path.join(B.CSRC, tu)] + flags.split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0:
        objs = [obj if u == tu else os.path.join(B.OBJ_DIR, f"libspectral_b200.{os.path.splitext(u)[0]}.o") for u in B.UNITS]
        r = subprocess.run(base + ["-shared", "-o", out] + objs, capture_output=True, text=True)
    return name, r.returncode, r.stderr[-3000:]


if __name__ == "__main__":
    args = sys.argv[1:]
    tu = "tu_feat2.cu"
    if args and args[0] == "--tu":
        tu, args = args[1], args[2:]
    B.build()
    with ThreadPoolExecutor(6) as ex:
        for name, rc, err in ex.map(lambda a: one(a, tu), args):
            print(name, "ok" if rc == 0 else "FAILED\n" + err)
