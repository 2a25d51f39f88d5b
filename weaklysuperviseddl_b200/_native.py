"""ctypes binding of libwsdl_b200.so (include/wsdl_b200.h).  There is no CPU or eager fallback:
if the library cannot be loaded every hot-path call raises."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwsdl_b200.so")

c_void_pp = ctypes.POINTER(ctypes.c_void_p)
c_int_p = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); mirrors include/wsdl_b200.h declaration by declaration
SIGNATURES = {
    "wsdl_version": (ctypes.c_int, []),
    "wsdl_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "wsdl_layercam_workspace_bytes": (ctypes.c_size_t, [c_int_p, c_int_p, c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "wsdl_layercam_fused": (
        ctypes.c_int,
        [c_void_pp, c_void_pp, c_int_p, c_int_p, c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
         ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    ),
    "wsdl_threshold_mask": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "wsdl_keep_largest_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "wsdl_keep_largest": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_size_t, ctypes.c_void_p],
    ),
    "wsdl_pairwise_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "wsdl_pairwise_fwd_bwd": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
         ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    ),
    "wsdl_pairwise_workspace_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "wsdl_pairwise_fwd_bwd_prepared": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
         ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    ),
    "wsdl_pairwise_dual_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "wsdl_pairwise_dual_fwd_bwd": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
         ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    ),
    "wsdl_weak_loss_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "wsdl_weak_loss_fwd_bwd": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong,
         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
         ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
         ctypes.c_void_p],
    ),
    "wsdl_count_valid_labels": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p],
    ),
    "wsdl_affinities": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "wsdl_labels_from_masks": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "wsdl_iou_acc_counts": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "wsdl_refine_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "wsdl_refine_step": (
        ctypes.c_int,
        [ctypes.c_void_p] * 10 + [ctypes.c_int] * 3 + [ctypes.c_float] * 5 + [ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                                                           ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    ),
    "wsdl_scale": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    ),
}

_lock = threading.Lock()
_lib = None


class WsdlError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Loads the native library once; raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise WsdlError(
                    f"{LIB_PATH} is missing: build it with `python -m weaklysuperviseddl_b200.build` "
                    "(needs nvcc; there is no CPU/eager fallback for the hot path)"
                )
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the .so is stale
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().wsdl_strerror(rc)
        raise WsdlError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
