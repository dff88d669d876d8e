"""Build pylians3_b200/libpyl_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pylians3_b200.build [--force] [--verbose]

The shared library is the C-ABI drop-in boundary declared in include/pyl_b200.h.  It is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# developer knobs for A/B timing of kernel variants: PYL_BUILD_TAG=x PYL_BUILD_DEFS="-DPYL_PT=256" builds
# libpyl_b200_x.so beside the product library; PYL_B200_SO=<path> makes _lib.py load it
TAG = os.environ.get("PYL_BUILD_TAG", "")
EXTRA = os.environ.get("PYL_BUILD_DEFS", "").split()
SO = os.path.join(PKG, "libpyl_b200%s.so" % ("_" + TAG if TAG else ""))
BUILD_DIR = os.path.join(PKG, "build" + ("_" + TAG if TAG else ""))
SOURCES = ["common.cu", "deposit.cu", "deposit_atomic.cu", "deposit_tiled.cu", "deposit_sorted.cu", "interp.cu", "fft.cu", "transpose.cu", "route.cu", "pk_bin.cu", "pk_shell.cu",
           "hostapi.cu"]
HEADERS = ["common.cuh", "stencil.cuh", "shell_body.cuh", "deposit_point.cuh", os.path.join(ROOT, "include", "pyl_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_LIB = "/usr/local/cuda/lib64"


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest(deps):
        return SO
    objs = []
    procs = []
    os.makedirs(BUILD_DIR, exist_ok=True)
    for s in srcs:
        o = os.path.join(BUILD_DIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) >= _newest([s] + deps[len(srcs):]):
            continue
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
               "-c", s, "-o", o] + EXTRA
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (os.path.basename(s), out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libpyl_b200.so")
    cmd = [NVCC, "-shared", "-o", SO] + objs + ["-L" + CUDA_LIB, "-lcufft", "-lcudart",
                                                 "-Xlinker", "-rpath," + CUDA_LIB]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
