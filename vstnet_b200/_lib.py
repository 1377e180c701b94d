"""ctypes binding of libvstb200.so (the C ABI declared in include/vstb200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets an exception — never a silent eager-PyTorch or CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvstb200.so")

MAX_STAGES = 8
MAX_LABELS = 256
MAX_STYLES = 16

CONV_FP32, CONV_TF32X3, CONV_TF32X2, CONV_TF32, CONV_F16X2 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": CONV_FP32, "tf32x3": CONV_TF32X3, "tf32x2": CONV_TF32X2, "tf32": CONV_TF32, "f16x2": CONV_F16X2}


class RevnetConfig(C.Structure):
    _fields_ = [
        ("n_stages", C.c_int),
        ("n_blocks", C.c_int * MAX_STAGES),
        ("n_strides", C.c_int * MAX_STAGES),
        ("n_channels", C.c_int * MAX_STAGES),
        ("in_channel", C.c_int),
        ("mult", C.c_int),
        ("hidden_dim", C.c_int),
        ("sp_steps", C.c_int),
        ("n_cr_blocks", C.c_int),
    ]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("ms", C.c_double), ("launches", C.c_longlong), ("flops", C.c_double),
                ("bytes", C.c_double)]


# name -> (restype, argtypes); every symbol include/vstb200.h declares
SIGNATURES = {
    "vst_last_error": (C.c_char_p, []),
    "vst_version": (C.c_int, []),
    "vst_launch_count": (C.c_ulonglong, []),
    "vst_profile_enable": (C.c_int, [C.c_int]),
    "vst_profile_collect": (C.c_int, [C.POINTER(ProfileEntry), C.c_int, C.POINTER(C.c_int)]),
    "vst_revnet_create": (C.c_int, [C.POINTER(RevnetConfig), C.POINTER(C.c_void_p)]),
    "vst_revnet_destroy": (None, [C.c_void_p]),
    "vst_revnet_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "vst_revnet_latent_channels": (C.c_int, [C.c_void_p]),
    "vst_revnet_down_scale": (C.c_int, [C.c_void_p]),
    "vst_revnet_param_floats": (C.c_size_t, [C.c_void_p]),
    "vst_revnet_packed_bytes": (C.c_size_t, [C.c_void_p]),
    "vst_revnet_pack_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vst_revnet_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vst_revnet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "vst_revnet_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "vst_revnet_stylize_supported": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vst_revnet_stylize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vst_cwct_stats_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "vst_cwct_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vst_cwct_stats2d": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vst_cwct_factor": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_float), C.c_int, C.c_float,
                                  C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "vst_cwct_whiten_factor": (C.c_int, [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "vst_cwct_color_factor": (C.c_int, [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "vst_cwct_cholesky": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vst_cwct_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vst_frame_u8_to_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vst_mask_resize_nearest": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vst_seg_scratch_bytes": (C.c_size_t, []),
    "vst_seg_remap": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "vst_seg_labels_from_colors": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "vst_frame_f32_to_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


class VstError(RuntimeError):
    pass


def load():
    """Load libvstb200.so (once).  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise VstError(
            "libvstb200.so not found at %s — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C vstnet_b200/csrc`. vstnet_b200 has no CPU / eager fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().vst_last_error()
        raise VstError("%s failed (rc=%d): %s" % (what, rc, (msg or b"").decode("utf-8", "replace")))


def launch_count():
    return int(load().vst_launch_count())


def profile_enable(on):
    check(load().vst_profile_enable(1 if on else 0), "vst_profile_enable")


def profile_collect(max_entries=128):
    """-> {kernel class: dict(ms, launches, flops, bytes)}; synchronises on the recorded events."""
    arr = (ProfileEntry * max_entries)()
    n = C.c_int(0)
    check(load().vst_profile_collect(arr, max_entries, C.byref(n)), "vst_profile_collect")
    return {arr[i].name.decode(): dict(ms=arr[i].ms, launches=int(arr[i].launches), flops=arr[i].flops,
                                       bytes=arr[i].bytes) for i in range(n.value)}
