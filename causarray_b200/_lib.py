"""ctypes binding of ``libcausarray_b200.so`` (the C ABI in ``include/causarray_b200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is
visible every numeric entry point raises ``CausarrayCudaError``.
"""
import ctypes
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CAUSARRAY_B200_LIB", os.path.join(_HERE, "libcausarray_b200.so"))

FAMILY = {"poisson": 0, "nb": 1}


class CausarrayCudaError(RuntimeError):
    pass


class AltminParams(ctypes.Structure):
    _fields_ = [
        ("alpha", ctypes.c_double), ("beta", ctypes.c_double), ("tol", ctypes.c_double), ("C", ctypes.c_double),
        ("ls_max_iters", ctypes.c_int),
        ("tol_gene", ctypes.c_double), ("tol_cell", ctypes.c_double),
        ("recheck_interval", ctypes.c_int),
        ("sparsity_threshold", ctypes.c_double), ("sparsity_boost", ctypes.c_double),
        ("warmup_iters", ctypes.c_int),
        ("es_max_iters", ctypes.c_int), ("es_warmup", ctypes.c_int), ("es_patience", ctypes.c_int),
        ("es_tolerance", ctypes.c_double), ("es_rel_tol", ctypes.c_double),
        ("lam", ctypes.c_double), ("a", ctypes.c_int),
    ]


class AltminResult(ctypes.Structure):
    _fields_ = [
        ("n_iter", ctypes.c_int), ("func_val", ctypes.c_double), ("resid", ctypes.c_double),
        ("hist_len", ctypes.c_int),
        ("total_gene_updates", ctypes.c_int64), ("total_cell_updates", ctypes.c_int64),
        ("cell_evals", ctypes.c_int64), ("gene_evals", ctypes.c_int64),
        ("loop_seconds", ctypes.c_double),
        ("gene_evals_onehot", ctypes.c_int64), ("onehot_treated_cells", ctypes.c_int64),
    ]


class LfcParams(ctypes.Structure):
    _fields_ = [
        ("usevar", ctypes.c_int), ("thres_min", ctypes.c_double), ("thres_diff", ctypes.c_double),
        ("eps_var", ctypes.c_double), ("yhat_cap", ctypes.c_double),
    ]


_P = ctypes.c_void_p
_D = ctypes.POINTER(ctypes.c_double)
_U8 = ctypes.POINTER(ctypes.c_uint8)
_I32 = ctypes.POINTER(ctypes.c_int32)
_c_int, _c_double = ctypes.c_int, ctypes.c_double

# name -> argtypes (restype is always int unless listed in _RESTYPE)
_SIGNATURES = {
    "ca_version": [],
    "ca_last_error": [],
    "ca_device_info": [ctypes.POINTER(_c_int)] * 3 + [ctypes.POINTER(ctypes.c_uint64)],
    "ca_set_device": [_c_int],
    "ca_release_cached": [],
    "ca_gcate_create": [ctypes.POINTER(_P), _c_int, _c_int, _c_int, _c_int, _c_int, _c_double],
    "ca_gcate_destroy": [_P],
    "ca_comm_unique_id": [_U8],
    "ca_gcate_set_shard": [_P, _c_int, _c_int, ctypes.c_int64, _U8],
    "ca_gcate_transport": [_P],
    "ca_gcate_load_counts": [_P, _D, _D, _D],
    "ca_gcate_upload_counts": [_P, _D] + [ctypes.POINTER(ctypes.c_int64)] * 3,
    "ca_gcate_upload_counts_typed": [_P, ctypes.c_void_p, _c_int] + [ctypes.POINTER(ctypes.c_int64)] * 3,
    "ca_gcate_upload_counts_rows": [_P, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), _c_int] + [ctypes.POINTER(ctypes.c_int64)] * 3,
    "ca_gcate_upload_counts_csr": [_P, ctypes.POINTER(ctypes.c_int64), _I32, _D, ctypes.c_int64] + [ctypes.POINTER(ctypes.c_int64)] * 3,
    "ca_gcate_size_factor_medians": [_P, _D],
    "ca_gcate_prepare": [_P, _D, _D],
    "ca_gcate_load_normalized": [_P, _D, _D, _D],
    "ca_gcate_set_factors": [_P, _D, _D],
    "ca_gcate_get_factors": [_P, _D, _D],
    "ca_gcate_set_stage": [_P, _D, _c_int, _D, _c_int, _D, _c_int],
    "ca_gcate_set_masks": [_P, _U8, _U8, _D],
    "ca_gcate_get_masks": [_P, _U8, _U8],
    "ca_gcate_update": [_P, _c_double, _c_double, _c_double, _c_int, _c_double, _D],
    "ca_gcate_nll": [_P, _D],
    "ca_gcate_grad": [_P, _c_int, _D],
    "ca_gcate_alter_min": [_P, ctypes.POINTER(AltminParams), ctypes.POINTER(AltminResult), _D],
    "ca_glm_create": [ctypes.POINTER(_P), _c_int, _c_int],
    "ca_glm_destroy": [_P],
    "ca_glm_load_counts": [_P, _D],
    "ca_glm_share_counts": [_P, _P],
    "ca_glm_fit": [_P, _c_int, _D, _c_int, _D, _D, _c_double, _c_int, _c_double, _c_int, _D, _D, _I32, _I32],
    "ca_glm_fit_ex": [_P, _c_int, _D, _c_int, _D, _D, _c_double, _c_int, _c_double, _c_int, _D, _U8, _D, _D, _I32, _I32],
    "ca_glm_refit_l1": [_P, _c_int, _D, _c_int, _D, _D, _c_double, _c_int, _c_double, _D, _U8, _D, _D, _I32, _I32],
    "ca_glm_fit_treatment": [_P, _c_int, _D, _c_int, _D, _D, _c_double, _c_int, _c_double, _D, _D, _I32, _I32],
    "ca_glm_resid": [_P, _D],
    "ca_glm_resid_device": [_P, ctypes.POINTER(_P)],
    "ca_glm_resid_device2": [_P, ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int64)],
    "ca_svd_topr": [_P, _P, ctypes.c_int64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_double, ctypes.c_uint32, _U8, _D, _D,
                    _D, _I32],
    "ca_glm_fitted": [_P, _D],
    "ca_glm_disp_moments": [_P, _D, _D],
    "ca_aipw_lfc": [_c_int] * 4 + [_D] * 8 + [_U8, ctypes.POINTER(LfcParams)] + [_D] * 5 + [_U8],
    "ca_aipw_lfc_shared": [_P, _c_int, _c_int] + [_D] * 7 + [_U8, ctypes.POINTER(LfcParams)] + [_D] * 5 + [_U8],
    "ca_aipw_lfc_yhat": [_c_int] * 3 + [_D] * 5 + [_U8, ctypes.POINTER(LfcParams)] + [_D] * 5 + [_U8],
    "ca_profile_enable": [_c_int],
    "ca_profile_reset": [],
    "ca_profile_num_classes": [],
    "ca_profile_class_name": [_c_int],
    "ca_profile_get": [_c_int, _D, ctypes.POINTER(ctypes.c_int64)],
    "ca_launch_count": [],
    "ca_nvtx_push": [ctypes.c_char_p],
    "ca_nvtx_pop": [],
    "ca_test_fastmath": [_D, _c_int, _c_int, _D],
    "ca_host_bh_rows": [_D, _c_int, _c_int, _D],
    "ca_host_median_mad_rows": [_D, _c_int, _c_int, _D, _D],
    "ca_inference_tail": [_D, _D, _c_int, _c_int, _D, _D, _D, _D, _D],
}
_RESTYPE = {"ca_last_error": ctypes.c_char_p, "ca_profile_class_name": ctypes.c_char_p,
            "ca_launch_count": ctypes.c_int64}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_gpu_ok = False


def load(require_gpu=True):
    """Load the shared library (once).  Raises ``CausarrayCudaError`` when it is absent."""
    global _lib, _gpu_ok
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CausarrayCudaError(
                "libcausarray_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, ctypes.c_int)
        _lib = lib
    if require_gpu and not _gpu_ok:
        sm = ctypes.c_int(0)
        for attempt in range(5):   # the first CUDA call on a freshly booted box can race the driver
            rc = _lib.ca_device_info(ctypes.byref(sm), None, None, None)
            if rc == 0:
                break
            time.sleep(1.0)
        if rc != 0:
            raise CausarrayCudaError("no usable CUDA device: %s" % _lib.ca_last_error().decode())
        _gpu_ok = True
    return _lib


def check(rc):
    if rc != 0:
        msg = _lib.ca_last_error().decode() if _lib is not None else "library not loaded"
        if rc == -2:
            raise ValueError(msg)
        raise CausarrayCudaError(msg)


def dptr(a):
    """Pointer to a C-contiguous float64 array (or NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_D)


def u8ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_U8)


def i32ptr(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_I32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def profile_enable(on=True):
    load().ca_profile_enable(1 if on else 0)


def profile_reset():
    load().ca_profile_reset()


def profile_snapshot():
    """{class name: (total device ms, launches)} since the last reset."""
    lib = load()
    out = {}
    for c in range(lib.ca_profile_num_classes()):
        ms = ctypes.c_double(0.0)
        cnt = ctypes.c_int64(0)
        lib.ca_profile_get(c, ctypes.byref(ms), ctypes.byref(cnt))
        out[lib.ca_profile_class_name(c).decode()] = (ms.value, int(cnt.value))
    return out


def launch_count():
    return int(load().ca_launch_count())
