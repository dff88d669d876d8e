"""Drop-in for the hot part of Pylians3's `smoothing_library`: `field_smoothing`
(library/smoothing_library/smoothing_library.pyx:215-235) -- FFT, complex product with the filter's transform,
inverse FFT, all on the GPU (cuFFT + pyl_cmul_inplace).  Building the filter itself (FT_filter, :20-120) is not
part of this path."""
from ._pk_more import field_smoothing  # noqa: F401

__all__ = ["field_smoothing"]
