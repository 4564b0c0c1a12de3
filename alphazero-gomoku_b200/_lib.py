"""ctypes binding of libazgomoku_b200.so (the C ABI in include/azgomoku_b200.h).

The library is built in-tree by ``make -C alphazero-gomoku_b200/csrc`` (or
``__graft_entry__.build()``).  There is no fallback of any kind: a missing
library raises at import time and every compute call fails when no sm_100
device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libazgomoku_b200.so")


class AzgError(RuntimeError):
    pass


class azg_pos(C.Structure):
    _fields_ = [("stones", (C.c_uint32 * 8) * 2), ("player", C.c_int32), ("last", C.c_int32),
                ("caps", C.c_int32 * 2), ("plies", C.c_int32), ("pad", C.c_int32 * 3)]


class azg_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("rule", C.c_int32), ("n_games", C.c_int32), ("queue_len", C.c_int32),
                ("node_capacity", C.c_int32), ("noise_on", C.c_int32), ("noise_plies", C.c_int32),
                ("game_base", C.c_int32), ("cpuct", C.c_double), ("alpha", C.c_double), ("eps", C.c_double),
                ("seed", C.c_uint64), ("fast_warps", C.c_int32), ("virtual_loss", C.c_int32)]


MAX_LAYERS = 80


class azg_net_weights(C.Structure):
    _fields_ = [("conv_w", C.c_void_p), ("bn", C.c_void_p * 4),
                ("res_conv_w", C.c_void_p * MAX_LAYERS), ("res_bn", (C.c_void_p * 4) * MAX_LAYERS),
                ("policy_conv_w", C.c_void_p), ("policy_bn", C.c_void_p * 4), ("policy_fc_w", C.c_void_p),
                ("policy_fc_b", C.c_void_p), ("value_conv_w", C.c_void_p), ("value_bn", C.c_void_p * 4),
                ("value_fc1_w", C.c_void_p), ("value_fc1_b", C.c_void_p), ("value_fc2_w", C.c_void_p),
                ("value_fc2_b", C.c_void_p)]


class azg_train_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_blocks", C.c_int32), ("channels", C.c_int32), ("max_batch", C.c_int32),
                ("lr", C.c_double), ("weight_decay", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("clip", C.c_double), ("bn_momentum", C.c_double), ("bn_eps", C.c_double)]


assert C.sizeof(azg_pos) == 96

_P = C.c_void_p
_I = C.c_int
# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
PROTOTYPES = {
    "azg_last_error": (C.c_char_p, []),
    "azg_abi_version": (_I, []),
    "azg_device_count": (_I, []),
    "azg_rules_pack": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "azg_rules_unpack": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "azg_rules_play": (_I, [_I, _P, _P, _P, _I, _P]),
    "azg_rules_status": (_I, [_I, _P, _P, _I, _P]),
    "azg_rules_legal": (_I, [_P, _P, _I, _P]),
    "azg_rules_encode": (_I, [_P, _P, _I, _P]),
    "azg_rules_play_host": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _P, _I]),
    "azg_rules_query_host": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I]),
    "azg_create": (_I, [C.POINTER(azg_config), C.POINTER(_P)]),
    "azg_destroy": (_I, [_P]),
    "azg_set_stream": (_I, [_P, _P]),
    "azg_memory_bytes": (C.c_int64, [_P]),
    "azg_set_roots": (_I, [_P, _P, _P, _I]),
    "azg_get_roots": (_I, [_P, _P]),
    "azg_search_begin": (_I, [_P, _P, _I]),
    "azg_search_begin_masked": (_I, [_P, _P, _I, _P]),
    "azg_search_fill": (_I, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "azg_search_read_counters": (_I, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "azg_search_counters": (_P, [_P]),
    "azg_search_leaf_planes": (_I, [_P, _P]),
    "azg_search_commit": (_I, [_P, _P, _P]),
    "azg_search_result": (_I, [_P, _P, _P]),
    "azg_search_advance": (_I, [_P, _P, _I, _I, _P]),
    "azg_search_stats": (_I, [_P, _P]),
    "azg_selfplay_enable": (_I, [_P, _I]),
    "azg_selfplay_set_active": (_I, [_P, _P]),
    "azg_selfplay_noise": (_I, [_P, C.c_uint64, _P]),
    "azg_selfplay_choose": (_I, [_P, _P, C.c_float, C.c_uint64, _P]),
    "azg_selfplay_finish": (_I, [_P, _P, _I, _I, _P, C.c_int64, _P, _P, _P]),
    "azg_selfplay_finish_packed": (_I, [_P, _P, _I, _P, C.c_int64, _P, _P, _P]),
    "azg_examples_expand": (_I, [_P, C.c_int64, _I, _P, _P]),
    "azg_net_create": (_I, [_I, _I, _I, _I, C.POINTER(_P)]),
    "azg_net_destroy": (_I, [_P]),
    "azg_net_memory_bytes": (C.c_int64, [_P]),
    "azg_net_load": (_I, [_P, C.POINTER(azg_net_weights), _P]),
    "azg_net_forward_planes": (_I, [_P, _P, _I, _P, _P, _P, _P]),
    "azg_net_forward_leaves": (_I, [_P, _P, _P, _P]),
    "azg_net_trunk_debug": (_I, [_P, _P, _I, _I, _P, _P]),
    "azg_net_profile": (_I, [_P, _I]),
    "azg_net_profile_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "azg_net_profile_counters": (_I, [_P, _P]),
    "azg_net_check": (_I, [_P, _P]),
    "azg_train_create": (_I, [C.POINTER(azg_train_config), C.POINTER(_P)]),
    "azg_train_destroy": (_I, [_P]),
    "azg_train_param_count": (C.c_int64, [_P]),
    "azg_train_memory_bytes": (C.c_int64, [_P]),
    "azg_train_bind": (_I, [_P, _P, _P, _P, _P, C.POINTER(azg_net_weights), C.c_int64, _P]),
    "azg_train_pack": (_I, [_P, _P]),
    "azg_train_forward_backward": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "azg_train_apply": (_I, [_P, _I, _P]),
    "azg_train_check": (_I, [_P, _P, _P]),
    "azg_train_read_activation": (_I, [_P, _I, _I, _P, _P]),
    "azg_train_export_grads": (_I, [_P, _P, _P]),
    "azg_train_debug_conv_grads": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(this package has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise AzgError(f"azgomoku_b200 error {rc}: {lib.azg_last_error().decode()}")


def ptr(t) -> C.c_void_p:
    """Device (or host) address of a torch tensor / numpy array / None."""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)
