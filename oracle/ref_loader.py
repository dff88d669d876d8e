"""Import the compiled, unmodified reference from oracle/_ref (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs
may use this.  The product package never imports anything under oracle/.

The extension modules are loaded straight from their .so files, bypassing the reference
packages' __init__.py (which pull in readgadget -> h5py, not installed here; reference
library/MAS_library/__init__.py:1-4).
"""
import glob
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_cache = {}


def _load(fullname, path):
    spec = importlib.util.spec_from_file_location(fullname, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _find(subdir, stem):
    hits = sorted(glob.glob(os.path.join(_REF, subdir, stem + "*.so")))
    return hits[0] if hits else None


def have_ref():
    return _find("MAS_library", "MAS_library") is not None and _find("Pk_library", "Pk_library") is not None


def ref_MASL(omp=False):
    """The reference's MAS_library extension (MA, NGP, CIC, ..., CICc3D, ...)."""
    key = "MASL_omp" if omp else "MASL"
    if key not in _cache:
        path = _find("omp" if omp else "MAS_library", "MAS_library")
        if path is None:
            raise ImportError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
        # module init name is PyInit_MAS_library -> last component must be MAS_library
        _cache[key] = _load(("ref_omp." if omp else "ref.") + "MAS_library", path)
    return _cache[key]


def ref_RSL():
    """The reference's redshift_space_library extension (pos_redshift_space)."""
    if "RSL" not in _cache:
        path = _find("redshift_space_library", "redshift_space_library")
        if path is None:
            raise ImportError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
        _cache["RSL"] = _load("ref.redshift_space_library", path)
    return _cache["RSL"]


def ref_PKL():
    """The reference's Pk_library extension (Pk, XPk, ...), FFT through the pyfftw shim."""
    if "PKL" not in _cache:
        path = _find("Pk_library", "Pk_library")
        if path is None:
            raise ImportError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
        shim = os.path.join(_HERE, "pyfftw_shim")
        if "pyfftw" not in sys.modules:
            sys.path.insert(0, shim)
            try:
                import pyfftw  # noqa: F401  (the shim, unless a real pyfftw is installed)
            finally:
                sys.path.remove(shim)
        _cache["PKL"] = _load("ref.Pk_library", path)
    return _cache["PKL"]


def ref_SL():
    """The reference's smoothing_library extension (FT_filter, field_smoothing, 2D variants).  It does
    `import Pk_library as PKL` at import time (smoothing_library.pyx:9): the compiled reference Pk_library is
    registered under that bare name for the duration of the import."""
    if "SL" not in _cache:
        path = _find("smoothing_library", "smoothing_library")
        if path is None:
            raise ImportError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
        saved = sys.modules.get("Pk_library")
        sys.modules["Pk_library"] = ref_PKL()
        try:
            _cache["SL"] = _load("ref.smoothing_library", path)
        finally:
            if saved is None:
                del sys.modules["Pk_library"]
            else:
                sys.modules["Pk_library"] = saved
    return _cache["SL"]
