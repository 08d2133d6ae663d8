"""Build libspectral_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One object per kernel family, compiled in parallel and linked into one shared library.  (--split-compile is not
used: with it ptxas scheduled the hot kernel differently from build to build, a 17 % swing in its run time.)"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
UNITS = ["spectral_b200.cu", "tu_feat2.cu", "tu_gl.cu", "tu_mstft.cu"]
OUT = os.path.join(HERE, "libspectral_b200.so")
OBJ_DIR = os.path.join(HERE, "build")


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "spectral_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stderr[-4000:])


def build(force=False, verbose=False, out=None, extra_flags=None):
    """Compile the CUDA extension.  Returns the path of the shared library."""
    out = out or OUT
    if out == OUT and not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
             "-diag-suppress", "550,177"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    flags += (extra_flags if extra_flags is not None else os.environ.get("SB200_NVCC_FLAGS", "").split())
    tag = os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = [os.path.join(OBJ_DIR, f"{tag}.{os.path.splitext(u)[0]}.o") for u in UNITS]
    with ThreadPoolExecutor(len(UNITS)) as ex:
        list(ex.map(lambda uo: _run([nvcc] + flags + ["-c", "-o", uo[1], os.path.join(CSRC, uo[0])], verbose),
                    zip(UNITS, objs)))
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, verbose)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
