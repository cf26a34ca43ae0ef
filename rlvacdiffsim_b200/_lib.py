"""ctypes binding of the C ABI in ``include/rlvd_b200.h`` (``librlvd_b200.so``, sm_100a).

There is no CPU fallback: if the library is missing or cannot be loaded, :func:`load` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librlvd_b200.so")

ACT_CODES = {"silu": 0, "relu": 1, "tanh": 2, "leaky_relu": 3}
QX_CODES = {"kT_x10": 0, "delta_e": 1}


class QParams(C.Structure):
    """``rlvd_q_params`` (include/rlvd_b200.h)."""
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("dqn", C.c_int32), ("temperature", C.c_float),
                ("delta_e_standardised", C.c_int32), ("painn_head_act", C.c_int32), ("dqn_head_act", C.c_int32),
                ("q_extra", C.c_int32 * 2),
                ("mean_barrier", C.c_float), ("std_barrier", C.c_float), ("mean_freq", C.c_float),
                ("std_freq", C.c_float), ("mean_delta_e", C.c_float), ("std_delta_e", C.c_float)]


class TimeHead(C.Structure):
    """``rlvd_time_head`` (include/rlvd_b200.h)."""
    _fields_ = [("d_w0", C.c_void_p), ("d_b0", C.c_void_p), ("d_w1", C.c_void_p), ("d_b1", C.c_void_p),
                ("d_w2", C.c_void_p), ("d_b2", C.c_void_p), ("n_feat", C.c_int32), ("pooling", C.c_int32),
                ("act", C.c_int32), ("tau", C.c_float), ("T_scale", C.c_float), ("D_scale", C.c_float),
                ("output", C.c_int32)]


# name -> (restype, argtypes); every symbol include/rlvd_b200.h declares
_P = C.c_void_p
SYMBOLS = {
    "rlvd_abi_version": (C.c_int, []),
    "rlvd_last_error": (C.c_char_p, []),
    "rlvd_engine_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "rlvd_engine_destroy": (None, [_P]),
    "rlvd_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, C.c_int32]),
    "rlvd_finalize_weights": (C.c_int, [_P]),
    "rlvd_radius_graph_count": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_float, _P, _P, _P]),
    "rlvd_radius_graph_fill": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_float, _P,
                                         C.c_int64, _P, _P]),
    "rlvd_painn_representation": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _P, _P,
                                            _P, _P, _P, _P, _P]),
    "rlvd_reaction_readout": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.POINTER(QParams), _P,
                                        _P, _P, _P, _P, _P, _P]),
    "rlvd_score_reactions_host": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_float,
                                            C.c_int32, _P, _P, C.POINTER(QParams), _P, _P, _P, _P, _P, _P, _P, _P]),
    "rlvd_action_space": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.c_double, C.c_int32, C.c_int32, _P, _P, _P]),
    "rlvd_candidate_offsets": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "rlvd_expand_candidates": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int32,
                                         _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rlvd_select_apply": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    "rlvd_select_apply_philox": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint32, _P, C.c_int32,
                                           C.c_int32, _P, _P, _P, _P, _P]),
    "rlvd_action_space_general": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_double, C.c_int32, C.c_int32, _P, _P, _P]),
    "rlvd_expand_candidates_ex": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P,
                                            _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rlvd_pack_step_records": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "rlvd_sro_counts": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, _P, _P]),
    "rlvd_reaction_readout_ex": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.POINTER(QParams), _P, _P, _P, _P, _P, _P, _P, _P]),
    "rlvd_q_head_train_step": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "rlvd_get_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "rlvd_time_readout": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, C.POINTER(TimeHead), _P]),
    "rlvd_linear": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P, C.c_int32, _P]),
    "rlvd_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "rlvd_set_hyper": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "rlvd_train_param_count": (C.c_int64, []),
    "rlvd_train_layout": (C.c_int, [C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "rlvd_train_forward_backward": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32,
                                              C.POINTER(QParams), _P, C.c_int32, _P, _P, _P, C.c_float, _P, _P, _P, _P, _P]),
    "rlvd_adam_step": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int32, _P]),
    "rlvd_check_status": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "rlvd_set_species": (C.c_int, [_P, _P, C.c_int32]),
    "rlvd_launch_count": (C.c_int64, [_P]),
    "rlvd_profile_enable": (C.c_int, [_P, C.c_int32]),
    "rlvd_profile_read": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "rlvd_profile_reset": (C.c_int, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """Load ``librlvd_b200.so`` and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m rlvacdiffsim_b200.build` (nvcc, sm_100a). "
            "This package has no CPU or eager-PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.rlvd_abi_version() != 2:
        raise RuntimeError("librlvd_b200.so has an unexpected ABI version")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("rlvd_b200: " + load().rlvd_last_error().decode(errors="replace"))
