"""ctypes binding of libgvn.so (include/gvn.h).  Fails loudly when the library is missing:
there is no CPU or torch fallback for the hot path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libgvn.so")

GVN_FRAME_ALIGN = 32
GVN_HIDDEN = 128
GVN_VS_TILE = 8
GVN_COST_TILE = 8
GVN_MAX_K = 32
GVN_MAX_L = 64
PREC_FP32, PREC_F16 = 0, 2
GVN_PREC_XV_CURRENT = 0x100
GVN_PREC_XV_BF16 = 0x200
PRECISIONS = {"fp32": PREC_FP32, "f16": PREC_F16, "fp32_xvbf16": PREC_FP32 | GVN_PREC_XV_BF16}
E_INVALID, E_UNSUPPORTED_SHAPE, E_CUDA, E_UNSUPPORTED_MODEL, E_BAD_WINDOW = -1, -2, -3, -4, -5

_p = C.c_void_p
_i = C.c_int32


class GvnBatch(C.Structure):
    _fields_ = [("B", _i), ("F", _i), ("K", _i), ("L", _i), ("NP", _i), ("R_cap", _i),
                ("frame_off", _p), ("n_frames", _p), ("frame_utt", _p),
                ("X2", _p), ("Xc", _p), ("W", _p), ("Wun", _p), ("H", _p), ("g", _p),
                ("Vb", _p), ("Z", _p), ("Vs", _p), ("yproj", _p), ("Vs_w", _p), ("XV", _p), ("X2t", _p)]


class GvnNoise(C.Structure):
    _fields_ = [("eps", _p), ("u", _p), ("forced_accept", _p), ("seed", C.c_uint64), ("chain", C.c_uint64)]


class GvnTrace(C.Structure):
    _fields_ = [("acc_prob", _p), ("accepted", _p), ("n_accepted", _p), ("z_samples", _p)]


# name -> (restype, argtypes); must list every symbol include/gvn.h declares
SIGNATURES = {
    "gvn_version": (_i, []),
    "gvn_last_error": (C.c_char_p, []),
    "gvn_decoder_packed_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "gvn_pack_decoder": (_i, [_p] * 6 + [_i, _i, _i, _i, _p, _p]),
    "gvn_label_projection": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "gvn_estep": (_i, [C.POINTER(GvnBatch), _p, _i, _i, C.c_float, C.POINTER(GvnNoise), C.POINTER(GvnTrace), _i, _p]),
    "gvn_mstep_workspace_bytes": (C.c_size_t, [C.POINTER(GvnBatch)]),
    "gvn_mstep": (_i, [C.POINTER(GvnBatch), _i, _p, _p, _i, _p]),
    "gvn_mstep_gain": (_i, [C.POINTER(GvnBatch), _i, _p, _p]),
    "gvn_cost_reduce": (_i, [C.POINTER(GvnBatch), _i, _i, _p, _p, _p]),
    "gvn_wiener": (_i, [C.POINTER(GvnBatch), _i, _p, _p, _p, _p, _p]),
    "gvn_stft_power": (_i, [C.POINTER(GvnBatch), _p, _i, _p, _p, _i, _i, _p]),
    "gvn_istft_workspace_bytes": (C.c_size_t, [C.POINTER(GvnBatch), _i]),
    "gvn_istft": (_i, [C.POINTER(GvnBatch), _p, _i, _i, _p, _p, _i, _p, _p]),
    "gvn_dense": (_i, [_p, _p, _p, _i, _p, _i, _p, _p, C.c_float, _i, _i, _i, _p, _p]),
    "gvn_spp_mask": (_i, [C.POINTER(GvnBatch), C.c_float, C.c_float, C.c_float, C.c_float, _i, _p, _p, _p]),
    "gvn_init_nmf": (_i, [C.POINTER(GvnBatch), _p, _p, C.c_float, _p]),
    "gvn_speech_labels_workspace_bytes": (C.c_size_t, [C.POINTER(GvnBatch)]),
    "gvn_speech_labels": (_i, [C.POINTER(GvnBatch), _p, _i, _i, C.c_float, C.c_float, _p, _p, _p]),
    "gvn_energy_ratios": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "gvn_selftest_umma": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "gvn_debug_profile_buffer": (None, [_p]),
    "gvn_launch_count": (C.c_uint64, []),
}


class GvnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libgvn error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Loads libgvn.so once.  Raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libgvn.so not found at %s -- build it with `make -C %s` or "
                          "`python __graft_entry__.py`; there is no fallback path"
                          % (LIB_PATH, os.path.join(os.path.dirname(_HERE), "csrc")))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Maps a negative status to the exception the reference raises for the same condition."""
    if rc == 0:
        return
    msg = load().gvn_last_error().decode("utf-8", "replace")
    if rc == E_UNSUPPORTED_MODEL:
        raise NameError(msg)
    if rc == E_BAD_WINDOW:
        raise ValueError(msg)
    raise GvnError(rc, msg)
