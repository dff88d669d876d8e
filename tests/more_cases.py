"""Shared driver for the sibling estimators (Pk_plane ... XXi): run every entry point of an implementation
`impl` (a namespace with the reference's names: the CPU oracle `oracle.cpu_more`, or the CUDA product
`pylians3_b200.Pk_library`) on the seeded inputs of tests/golden/make_golden_more.py and return the results
under the golden file's keys."""
import contextlib
import importlib.util
import io
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BOX = 1000.0
_spec = importlib.util.spec_from_file_location("make_golden_more", os.path.join(HERE, "golden", "make_golden_more.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)
inputs = _gen.inputs
SIZES = _gen.SIZES
FILTERS = _gen.FILTERS


def run_smoothing(sl, N, I=None):
    """The smoothing_library entry points of `sl` under the golden file's keys."""
    I = inputs(N) if I is None else I
    t = "N%d_" % N
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for name, (F, R, kmin, kmax) in FILTERS.items():
            out[t + "sm_f3_" + name] = np.asarray(sl.FT_filter(BOX, R, N, F, 2, kmin, kmax))
            out[t + "sm_f2_" + name] = np.asarray(sl.FT_filter_2D(BOX, R, N, F, 2, kmin, kmax))
        out[t + "sm_s3"] = np.asarray(sl.field_smoothing(I["d1"], out[t + "sm_f3_gauss"], 2))
        out[t + "sm_s2"] = np.asarray(sl.field_smoothing_2D(I["img1"], out[t + "sm_f2_tophat"], 2))
    return out


def compare_smoothing(got, ref, tol):
    """Filters and smoothed fields: complex / real arrays, error relative to the array's largest amplitude."""
    bad = []
    for key, g in got.items():
        r = np.asarray(ref[key])
        g = np.asarray(g)
        if g.shape != r.shape or g.dtype != r.dtype:
            bad.append((key, "shape/dtype", g.shape, g.dtype, r.shape, r.dtype))
            continue
        e = float(np.max(np.abs(g - r)) / np.max(np.abs(r)))
        if not e < tol:
            bad.append((key, e, tol))
    return bad


def run_xxi_multi(impl, N, I=None):
    I = inputs(N) if I is None else I
    with contextlib.redirect_stdout(io.StringIO()):
        r = impl.XXi_multi([I["d1"], I["d2"], I["f1"]], BOX, 1, ["CIC", "PCS", "NGP"], 1)
    return {"N%d_xm_r" % N: r.r3D, "N%d_xm_Nm" % N: r.Nmodes3D, "N%d_xm_xi" % N: r.xi, "N%d_xm_Xxi" % N: r.Xxi}


def compare_xxi_multi(got, ref, tol, N):
    """Counts exact, radii 1e-12, multipoles against max(|value|, 0.01 (2l+1) max|xi_0 of the autos|)."""
    t = "N%d_xm_" % N
    bad = []
    if not np.array_equal(got[t + "Nm"], ref[t + "Nm"]):
        bad.append("counts")
    if np.max(np.abs(got[t + "r"] / ref[t + "r"] - 1)) > 1e-12:
        bad.append("r3D")
    floor = 0.01 * np.max(np.abs(ref[t + "xi"][:, 0, :])) * np.array([1.0, 5.0, 9.0])[None, :, None]
    for k in ("xi", "Xxi"):
        g, r = np.asarray(got[t + k]), np.asarray(ref[t + k])
        if g.shape != r.shape:
            bad.append((k, g.shape, r.shape))
            continue
        e = float(np.max(np.abs(g - r) / np.maximum(np.abs(r), floor)))
        if not e < tol:
            bad.append((k, e))
    return bad


def run_all(impl, N, I=None):
    I = inputs(N) if I is None else I
    t = "N%d_" % N
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        r = impl.Pk_plane(I["img1"], BOX, "CIC", 1, False)
        out.update({t + "plane_k": r.k, t + "plane_Pk": r.Pk, t + "plane_Nm": r.Nmodes})
        r = impl.XPk_plane(I["img1"], I["img2"], BOX, "CIC", "PCS", 1)
        out.update({t + "xplane_k": r.k, t + "xplane_Pk": r.Pk, t + "xplane_XPk": r.XPk, t + "xplane_Nm": r.Nmodes,
                    t + "xplane_r": r.r})
        for axis in (0, 1, 2):
            a = t + "imag_a%d_" % axis
            r = impl.XPk_imag([I["f1"], I["f2"], I["f3"]], BOX, axis, ["CIC", "NGP", "PCS"], 1)
            for n in ("k3D", "Nmodes3D", "Pk", "XPk", "k1D", "Nmodes1D", "Pk1D", "PkX1D", "Nmodes2D", "Pk2D", "PkX2D"):
                out[a + n] = np.asarray(getattr(r, n))
            r = impl.XPk_2D(I["f1"], I["f2"], BOX, axis, "TSC", "CIC", 1)
            for i, n in enumerate(("kpar", "kper", "Pk1", "Pk2", "PkX", "Nm")):
                out[t + "x2d_a%d_%s" % (axis, n)] = np.asarray(r[i])
            r = impl.Xi(I["d1"], BOX, "CIC", axis, 1)
            out.update({t + "xi_a%d_r" % axis: r.r3D, t + "xi_a%d_xi" % axis: r.xi, t + "xi_a%d_Nm" % axis: r.Nmodes3D})
        r = impl.XXi(I["d1"], I["d2"], BOX, ["CIC", "PCS"], 2, 1)
        out.update({t + "xxi_r": r.r3D, t + "xxi_xi": r.xi, t + "xxi_Nm": r.Nmodes3D})
        r = impl.Pk_theta(I["Vx1"], I["Vy1"], I["Vz1"], BOX, 2, "CIC", 1)
        out.update({t + "theta_k": r[0], t + "theta_Pk": r[1], t + "theta_Nm": r[2]})
        V = [I[n].copy() for n in ("Vx1", "Vy1", "Vz1")]
        r = impl.XPk_dv(I["d1"], V[0], V[1], V[2], BOX, 2, "TSC", 1)
        for i, n in enumerate(("k", "Pk1", "Pk2", "PkX", "Nm")):
            out[t + "dv_" + n] = np.asarray(r[i])
        out[t + "dv_Vx_after"] = V[0]
        V = [I[n].copy() for n in ("Vx1", "Vy1", "Vz1", "Vx2", "Vy2", "Vz2")]
        r = impl.XPk_vv(I["d1"], V[0], V[1], V[2], I["d2"], V[3], V[4], V[5], BOX, 2, "PCS", 1)
        for i, n in enumerate(("k", "Pk1", "Pk2", "PkX", "Nm")):
            out[t + "vv_" + n] = np.asarray(r[i])
        out[t + "cmas"] = np.asarray(impl.correct_MAS(I["d1"].copy(), BOX, "PCS", 1))
        r = impl.expected_Pk(I["k_in"], I["Pk_in"], BOX, N, 300)
        out.update({t + "exp_k": np.asarray(r[0]), t + "exp_Pk": np.asarray(r[1]), t + "exp_Nm": np.asarray(r[2])})
    return out


_PAIRS = [(0, 1), (0, 2), (1, 2)]


def _cross_floor(key, ref):
    """Natural scale sqrt(P_i P_j) of a cross quantity, from the reference's auto spectra (None for other keys).
    A cross sum is a sum of signed terms: where it cancels, its round-off is set by the autos, not by itself."""
    pre, name = key.rsplit("_", 1)
    with np.errstate(invalid="ignore"):
        if name == "PkX" and (pre.endswith("dv") or pre.endswith("vv") or "x2d" in pre):
            return np.sqrt(np.abs(ref[pre + "_Pk1"] * ref[pre + "_Pk2"]))
        if key.endswith("xplane_XPk"):
            P = ref[pre + "_Pk"]
            return np.sqrt(np.abs(P[:, 0] * P[:, 1]))
        if "imag" in pre and name in ("XPk", "PkX1D", "PkX2D"):
            P = ref[pre + "_" + {"XPk": "Pk", "PkX1D": "Pk1D", "PkX2D": "Pk2D"}[name]]
            if name == "XPk":
                mono = P[:, 0, :]
                fl = np.stack([np.sqrt(np.abs(mono[:, i] * mono[:, j])) for i, j in _PAIRS], axis=-1)
                return fl[:, None, :] * np.array([1.0, 5.0, 9.0])[None, :, None]
            return np.stack([np.sqrt(np.abs(P[:, i] * P[:, j])) for i, j in _PAIRS], axis=-1)
    return None


def compare(got, ref, tol, ktol=1e-12, fft_eps=0.0):
    """Every key of `got` against `ref`: |got - ref| <= tol * den + fft_eps * sqrt(den * peak).

    Counts (...Nm, ...Nmodes*) bit-exact; wavenumbers / radii to `ktol`.  For everything else `den` is the
    natural scale of the entry: max(|value|, 1e-6 * the array's largest |value|), and
      * cross terms: also 0.1*sqrt(P_i P_j) -- a sum of signed terms, where it cancels its round-off is set
        by the autos (SURVEY 8a note on degenerate bins);
      * multipoles l = 2, 4: also (2l+1) * |monopole| (the corner-mode quadrupole is pure round-off);
      * correlation-function bins: also 0.01 (2l+1) max|xi_0| (zero crossings);
      * real-space fields: the field's amplitude.
    `fft_eps` (GPU tests: 1e-5, as in test_gpu_pk.py) is the single-precision FFT floor: two float32 FFT
    libraries agree per mode to ~1e-6 of the LARGEST amplitudes, so a bin far below the peak moves by more than
    `tol` of itself whichever library is used.  The kernels themselves are held to 1e-10 on identical input
    (test_shell_kernels_same_input)."""
    bad = []
    ell = np.array([1.0, 5.0, 9.0])
    for key, g in got.items():
        r = np.asarray(ref[key], dtype=np.float64)
        g = np.asarray(g, dtype=np.float64)
        if g.shape != r.shape:
            bad.append((key, "shape", g.shape, r.shape))
            continue
        name = key.split("_")[-1]
        if name in ("Nm", "Nmodes3D", "Nmodes1D", "Nmodes2D"):
            if not np.array_equal(g, r):
                bad.append((key, "counts differ"))
            continue
        nan_g, nan_r = np.isnan(g), np.isnan(r)
        if not np.array_equal(nan_g, nan_r):
            bad.append((key, "NaN pattern"))
            continue
        if r.size == 0 or nan_r.all():
            continue
        is_k = name in ("k", "k3D", "k1D", "kpar", "kper", "r") and "xplane_r" not in key
        t = ktol if is_k else tol
        if "exp_k" in key:
            t = max(t, 1e-6)          # float32 k in the reference
        scale = np.nanmax(np.abs(r))
        den = np.maximum(np.abs(r), 0.0 if is_k else 1e-6 * scale)
        fl = _cross_floor(key, ref)
        if fl is not None:
            den = np.maximum(den, 0.1 * np.nan_to_num(fl))
        if "imag" in key and name == "Pk":
            den = np.maximum(den, np.abs(r[:, :1, :]) * ell[None, :, None])
        if name == "xi":
            den = np.maximum(den, 0.01 * np.nanmax(np.abs(r[:, 0])) * ell[None, :])
        if name == "cmas":
            den = np.full_like(r, scale)                 # a real-space field: error relative to its amplitude
        if "xplane_r" in key:
            den = np.maximum(den, 0.1)                   # correlation coefficient: |r| <= 1
        den = np.where(den == 0, 1.0, den)
        allowed = t * den
        if not is_k and "exp_" not in key:
            allowed = allowed + fft_eps * np.sqrt(den * np.nanmax(den))
        excess = np.abs(g - r) - allowed
        if np.nanmax(excess) > 0:
            i = int(np.nanargmax(excess))
            bad.append((key, "worst |d|/den = %.3e at flat index %d" % (float(np.abs(g - r).flat[i] / den.flat[i]), i), t))
    return bad
