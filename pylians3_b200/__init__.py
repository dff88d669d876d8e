"""pylians3_b200 -- B200-native (sm_100a) implementation of Pylians3's density-field and
power-spectrum hot path:  MAS_library.MA  ->  FFT  ->  Pk_library.Pk / XPk.

    from pylians3_b200 import MAS_library as MASL, Pk_library as PKL

or put `dropin/` on PYTHONPATH to keep the reference's bare module names
(`import MAS_library as MASL`, `import Pk_library as PKL`).

All arithmetic runs in pylians3_b200/libpyl_b200.so (C ABI: include/pyl_b200.h), built by
`python -m pylians3_b200.build`.  There is no CPU fallback: without the built library, or
without a CUDA device, the entry points raise.
"""
__version__ = "0.1.0"

from . import MAS_library, Pk_library, redshift_space_library, smoothing_library  # noqa: E402,F401
from .field import overdensity_, prebias_  # noqa: E402,F401
