"""ctypes binding of libpyl_b200.so (the C ABI declared in include/pyl_b200.h).

There is NO fallback: if the CUDA library has not been built, importing any compute entry
point raises.  Build it with `python -m pylians3_b200.build` (or __graft_entry__.build()).
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PYL_B200_SO") or os.path.join(_PKG, "libpyl_b200.so")

PYL_OK = 0
MAS_IDS = {"NGP": 0, "CIC": 1, "TSC": 2, "PCS": 3}
MODE_IDS = {"auto": -1, "atomic": 0, "tiled": 1, "deterministic": 2}
MAX_FIELDS = 4


class PylError(RuntimeError):
    """A pyl_* entry point returned a negative status."""

    def __init__(self, status, where, detail):
        self.status = status
        super().__init__("%s failed (%d: %s)%s" % (where, status, _error_string(status),
                                                    (" -- " + detail) if detail else ""))


SHELL_KINDS = {"theta": 0, "dv": 1, "vv": 2, "expected": 3, "plane": 4, "xplane": 5, "xi": 6}
PK_PHASE, PK_CROSS_IMAG, PK_KY_MAJOR = 1, 2, 4


class ShellTable(ctypes.Structure):
    """Mirror of pyl_shell_table_t."""
    _fields_ = [("k", ctypes.c_void_p), ("P", ctypes.c_void_p), ("n", ctypes.c_int32), ("kF", ctypes.c_float),
                ("log10_kmin", ctypes.c_double), ("deltak", ctypes.c_double)]


class PkLayout(ctypes.Structure):
    """Mirror of pyl_pk_layout_t."""
    _fields_ = [(n, ctypes.c_int32) for n in ("dims", "fields", "xfields", "kmax_par", "kmax_per", "kmax")] + \
               [(n, ctypes.c_int64) for n in ("n2d", "k3D", "Nm3D", "Pk3D", "PkX3D", "phase", "Nm1D", "Pk1D",
                                              "PkX1D", "Nm2D", "Pk2D", "PkX2D", "total_words")]


# every symbol include/pyl_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
_l = ctypes.c_long
PROTOTYPES = {
    "pyl_error_string": (ctypes.c_char_p, [_i]),
    "pyl_last_error": (ctypes.c_char_p, []),
    "pyl_version": (ctypes.c_char_p, []),
    "pyl_kernel_launches": (ctypes.c_ulonglong, []),
    "pyl_deposit_workspace_bytes": (_sz, [_i, _i64, _i, _i, _i]),
    "pyl_deposit": (_i, [_i, _vp, _vp, _vp, _i64, _i, _i, _f, _i, _vp, _sz, _vp]),
    "pyl_deposit_slab_workspace_bytes": (_sz, [_i, _i64, _i, _i]),
    "pyl_deposit_slab": (_i, [_i, _vp, _vp, _vp, _i64, _i, _f, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "pyl_deposit_slab_counted": (_i, [_i, _vp, _vp, _vp, _i64, _vp, _i, _f, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "pyl_route_block_bytes": (_sz, [_i64]),
    "pyl_route_scatter": (_i, [_i, _vp, _vp, _i64, _i, _f, _i, _vp, _vp, _i64, _vp, _vp]),
    "pyl_stencil_base_plane": (_i, [_i, _vp, _i64, _i, _f, _vp, _vp]),
    "pyl_cic_interp": (_i, [_vp, _i, _f, _vp, _i64, _vp, _vp]),
    "pyl_pos_redshift_space": (_i, [_vp, _vp, _i64, _f, _f, _f, _i, _vp]),
    "pyl_divide_inplace": (_i, [_vp, _i64, _f, _vp]),
    "pyl_scale_inplace": (_i, [_vp, _i64, _f, _vp]),
    "pyl_affine_inplace": (_i, [_vp, _i64, _f, _f, _vp]),
    "pyl_sum_f64": (_i, [_vp, _i64, _vp, _vp]),
    "pyl_overdensity_inplace": (_i, [_vp, _i64, _vp, ctypes.c_double, _vp]),
    "pyl_fill_negative": (_i, [_vp, _i64, _vp, ctypes.c_double, _vp, _vp]),
    "pyl_add_inplace": (_i, [_vp, _vp, _i64, _vp]),
    "pyl_fft_r2c_workspace_bytes": (_sz, [_i]),
    "pyl_fft_r2c": (_i, [_vp, _vp, _i, _vp, _sz, _vp]),
    "pyl_fft_slab_workspace_bytes": (_sz, [_i, _i, _i]),
    "pyl_fft_slab_yz": (_i, [_vp, _vp, _i, _i, _vp, _sz, _vp]),
    "pyl_fft_slab_x": (_i, [_vp, _i, _i, _vp, _sz, _vp]),
    "pyl_transpose_scatter": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_i), _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "pyl_transpose_scatter_kymajor": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_i), _vp, _vp, _vp, _i, _i, _i, _i,
                                           _vp]),
    "pyl_fft_slab_x_kymajor_workspace_bytes": (_sz, [_i]),
    "pyl_fft_slab_x_kymajor": (_i, [_vp, _i, _i, _vp, _sz, _vp]),
    "pyl_fft_clear_plans": (_i, []),
    "pyl_pk_layout": (_i, [_i, _i, ctypes.POINTER(PkLayout)]),
    "pyl_pk_bin_workspace_bytes": (_sz, [_i, _i]),
    "pyl_pk_bin": (_i, [ctypes.POINTER(_vp), _i, ctypes.POINTER(_i), _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "pyl_pk_finalize": (_i, [_vp, _i, _i, ctypes.c_double, _i, _vp, _vp, _vp]),
    "pyl_pk_counts_to_f64": (_i, [_vp, _i, _i, _vp]),
    "pyl_pk_take_dc": (_i, [_vp, _i, _i, _vp, _vp]),
    "pyl_pk_density_scale": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "pyl_pk_clear_cache": (_i, []),
    "pyl_pk_mirrored_rows": (_i, [_i, _i, _i, ctypes.POINTER(_i)]),
    "pyl_pk_bin_mirrored": (_i, [ctypes.POINTER(_vp), _i, ctypes.POINTER(_i), _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "pyl_fft_c2r_workspace_bytes": (_sz, [_i]),
    "pyl_fft_c2r": (_i, [_vp, _vp, _i, _vp, _sz, _vp]),
    "pyl_shell_layout": (_i, [_i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "pyl_shell_bin_workspace_bytes": (_sz, [_i, _i]),
    "pyl_shell_bin": (_i, [_i, ctypes.POINTER(_vp), _i, ctypes.POINTER(_i), _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "pyl_modes_workspace_bytes": (_sz, [_i]),
    "pyl_modes_deconvolve": (_i, [_vp, _i, _i, _vp, _sz, _vp]),
    "pyl_modes_power": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "pyl_modes_power_2d": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "pyl_radial_bin_2d": (_i, [_vp, _i, _f, _vp, _vp]),
    "pyl_cmul_inplace": (_i, [_vp, _vp, _i64, _vp]),
    "pyl_mul_one_plus": (_i, [_vp, _vp, _i64, _vp]),
    "pyl_filter_fill": (_i, [_i, _vp, _i, _i, _f, _f, _f, _f, _vp]),
    "pyl_divide_by_f64": (_i, [_vp, _i64, _vp, _vp]),
    "pyl_fft2d_c2r_workspace_bytes": (_sz, [_i]),
    "pyl_fft2d_c2r": (_i, [_vp, _vp, _i, _vp, _sz, _vp]),
    "pyl_NGP": (_i, [_vp, _vp, _vp, _l, _i, _i, _f, _i]),
    "pyl_CIC": (_i, [_vp, _vp, _vp, _l, _i, _i, _f, _i]),
    "pyl_TSC": (_i, [_vp, _vp, _vp, _l, _i, _i, _f, _i]),
    "pyl_PCS": (_i, [_vp, _vp, _vp, _l, _i, _i, _f, _i]),
    "pyl_host_arena_release": (_i, []),
}

_lib = None


def load():
    """Load the shared library once; raise loudly if it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "pylians3_b200: %s is missing -- the CUDA extension has not been built "
            "(run `python -m pylians3_b200.build`). There is no CPU fallback." % SO_PATH)
    lib = ctypes.CDLL(SO_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError("pylians3_b200: %s does not export %s (stale build?)" % (SO_PATH, name)) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _error_string(status):
    try:
        return load().pyl_error_string(status).decode()
    except Exception:
        return "?"


def check(status, where):
    if status != PYL_OK:
        detail = load().pyl_last_error().decode(errors="replace")
        raise PylError(status, where, detail)


def pk_layout(dims, fields):
    L = PkLayout()
    check(load().pyl_pk_layout(int(dims), int(fields), ctypes.byref(L)), "pyl_pk_layout")
    return L
