"""Minimal stand-in for the third-party `pyfftw` module (TEST INFRASTRUCTURE ONLY).

The reference imports pyfftw at library/Pk_library/Pk_library.pyx:3 and uses exactly two
names from it in the FFT wrappers (Pk_library.pyx:117-242): `empty_aligned(shape, dtype)`
and `FFTW(a_in, a_out, axes, flags, direction, threads)` called as `plan(a_in, a_out)`.
pyfftw / FFTW3 are not installed in this image (unpinned dependency, reference setup.py:137),
so the compiled reference's *own* deconvolution + binning code is driven through this shim,
which performs the transform with scipy's pocketfft (native single precision for float32
input, like FFTW's float interface).

Backward transforms: pyfftw's `FFTW.__call__(input_array, output_array, normalise_idft=True)`
scales an inverse transform by 1/N by default, and the reference calls the plan with the two
arrays only (Pk_library.pyx:149-163) -- so IFFT3Dr_f returns the NORMALISED inverse.  The
reference's own arithmetic confirms it: Xi (Pk_library.pyx:2221, 2276) multiplies the inverse
transform of |delta_k|^2 by one further 1/dims^3, which is the correlation function only if
the inverse already carried its 1/dims^3.  scipy's irfftn/ifftn are normalised the same way,
so the shim returns them unscaled.
"""
import numpy as np
import scipy.fft as _sfft


def empty_aligned(shape, dtype="float64", order="C", n=None):
    return np.empty(shape, dtype=dtype, order=order)


class FFTW:
    def __init__(self, a_in, a_out, axes=(-1,), flags=("FFTW_MEASURE",),
                 direction="FFTW_FORWARD", threads=1, **kw):
        self.axes = tuple(axes)
        self.direction = direction
        self.threads = int(threads) if threads else 1
        self.real_in = not np.iscomplexobj(a_in)
        self.real_out = not np.iscomplexobj(a_out)

    def __call__(self, a_in, a_out):
        w = self.threads
        if self.direction == "FFTW_FORWARD":
            if self.real_in:
                a_out[...] = _sfft.rfftn(a_in, axes=self.axes, workers=w)
            else:
                a_out[...] = _sfft.fftn(a_in, axes=self.axes, workers=w)
        else:
            if self.real_out:
                s = [a_out.shape[ax] for ax in self.axes]
                a_out[...] = _sfft.irfftn(a_in, s=s, axes=self.axes, workers=w)
            else:
                s = [a_out.shape[ax] for ax in self.axes]
                a_out[...] = _sfft.ifftn(a_in, axes=self.axes, workers=w)
        return a_out
