#!/usr/bin/env python
"""Build the UNMODIFIED reference hot path into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE -- never imported by the product package `pylians3_b200`.

What is built (sources are read where they lie under /root/reference; nothing is
copied into this repository, and every output lands in oracle/_ref/, which is
git-ignored but still travels to the GPU box):

  oracle/_ref/MAS_library/MAS_library*.so  <- library/MAS_library/MAS_library.pyx + MAS_c.c
  oracle/_ref/Pk_library/Pk_library*.so    <- library/Pk_library/Pk_library.pyx
  oracle/_ref/redshift_space_library/redshift_space_library*.so
                                           <- library/redshift_space_library/redshift_space_library.pyx
  oracle/_ref/smoothing_library/smoothing_library*.so
                                           <- library/smoothing_library/smoothing_library.pyx (built with -fopenmp:
                                              its loops are `prange`; it imports Pk_library at run time)
  oracle/_ref/omp/MAS_library*.so          <- same MAS sources, with -fopenmp at compile
                                              time (NON-default build: the reference's
                                              setup.py:22 typo drops -fopenmp)

Flags follow the reference's setup.py:26-31,124-132 (-O3 -ffast-math, directives
legacy_implicit_noexcept / language_level=3) with ONE deliberate deviation:
-march=x86-64-v3 instead of -march=native, because the .so is built in the authoring
container (Sapphire Rapids) but executed on the GPU box's host CPU, which may lack
AVX512-FP16/AMX.  x86-64-v3 keeps AVX2+FMA, i.e. the same FMA contraction the reference
binary shows (SURVEY §8a "vmulss + vfmadd213ss").

The reference's own build system (setup.py, 13 extensions) is NOT run.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

REF = os.environ.get("PYLIANS3_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
CC = "/usr/bin/gcc"          # the venv's gcc wrapper cannot link -fopenmp
CFLAGS = ["-O3", "-ffast-math", "-march=x86-64-v3", "-fPIC", "-fwrapv", "-w",
          "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]


def _ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def _cythonize(pyx, c_out, include):
    os.makedirs(os.path.dirname(c_out), exist_ok=True)
    cmd = [sys.executable, "-m", "cython", "-3", "-X", "legacy_implicit_noexcept=True",
           "-I", include, "-o", c_out, pyx]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _compile(sources, so_out, include_dirs, extra=()):
    import numpy
    os.makedirs(os.path.dirname(so_out), exist_ok=True)
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    inc += ["-I" + d for d in include_dirs]
    cmd = [CC, "-shared"] + CFLAGS + list(extra) + inc + sources + ["-o", so_out, "-lm", "-lgomp"]
    subprocess.check_call(cmd)


def available():
    return os.path.isdir(os.path.join(REF, "library", "MAS_library"))


def build(force=False):
    """Returns True if oracle/_ref is usable afterwards."""
    suf = _ext_suffix()
    mas_so = os.path.join(OUT, "MAS_library", "MAS_library" + suf)
    pk_so = os.path.join(OUT, "Pk_library", "Pk_library" + suf)
    omp_so = os.path.join(OUT, "omp", "MAS_library" + suf)
    rsd_so = os.path.join(OUT, "redshift_space_library", "redshift_space_library" + suf)
    sm_so = os.path.join(OUT, "smoothing_library", "smoothing_library" + suf)
    have = all(os.path.exists(p) for p in (mas_so, pk_so, omp_so, rsd_so, sm_so))
    if have and not force:
        return True
    if not available():
        return have
    mas_dir = os.path.join(REF, "library", "MAS_library")
    pk_dir = os.path.join(REF, "library", "Pk_library")
    gen = os.path.join(OUT, "gen")
    mas_c = os.path.join(gen, "MAS_library.c")
    pk_c = os.path.join(gen, "Pk_library.c")
    _cythonize(os.path.join(mas_dir, "MAS_library.pyx"), mas_c, mas_dir)
    _cythonize(os.path.join(pk_dir, "Pk_library.pyx"), pk_c, pk_dir)
    _compile([mas_c, os.path.join(mas_dir, "MAS_c.c")], mas_so, [mas_dir])
    _compile([mas_c, os.path.join(mas_dir, "MAS_c.c")], omp_so, [mas_dir], extra=["-fopenmp"])
    _compile([pk_c], pk_so, [pk_dir])
    rsd_dir = os.path.join(REF, "library", "redshift_space_library")
    rsd_c = os.path.join(gen, "redshift_space_library.c")
    _cythonize(os.path.join(rsd_dir, "redshift_space_library.pyx"), rsd_c, rsd_dir)
    _compile([rsd_c], rsd_so, [rsd_dir])
    sm_dir = os.path.join(REF, "library", "smoothing_library")
    sm_c = os.path.join(gen, "smoothing_library.c")
    _cythonize(os.path.join(sm_dir, "smoothing_library.pyx"), sm_c, sm_dir)
    _compile([sm_c], sm_so, [sm_dir], extra=["-fopenmp"])
    shutil.rmtree(gen, ignore_errors=True)   # generated C is large; keep only the .so files
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference and no prebuilt .so)")
