"""ctypes binding of libyacht_b200.so (the C ABI declared in include/yacht_b200.h).

There is deliberately no fallback: if the CUDA library has not been built, importing a symbol
raises; if it is called without a CUDA device the CUDA runtime error is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libyacht_b200.so")

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32
_int = ctypes.c_int

# name -> argtypes; every function returns int (0 / cudaError_t).  Must list every symbol of
# include/yacht_b200.h (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    "ya_abi_version": [],
    "ya_set_device": [_int],
    "ya_init_states": [_vp, _i64, _vp, _vp, _vp, _i64, _u64, _u64, _vp],
    "ya_next_state": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _vp, _u64, _u64, _vp, _vp, _u32, _vp, _vp],
    "ya_valid_moves": [_vp, _i64, _vp, _vp, _i64, _vp],
    "ya_game_ended": [_vp, _i64, _vp, _vp, _i64, _vp],
    "ya_canonical_form": [_vp, _i64, _vp, _vp, _i64, _i64, _vp],
    "ya_features": [_vp, _i64, _vp, _i64, _vp],
    "ya_random_action": [_vp, _i64, _vp, _vp, _i64, _u64, _u64, _vp, _vp, _vp],
    "ya_enumerate_scores": [_vp, _i64, _vp, _vp, _i64, _vp],
    "ya_greedy_action": [_vp, _i64, _vp, _vp, _vp, _i64, _int, _u64, _u64, _vp, _vp, _vp],
    "ya_play_ply": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _u64, _u64, _int, _vp],
    "ya_mcts_cursor_words": [],
    "ya_mcts_node_words": [],
    "ya_mcts_reset": [_vp, _vp, _vp],
    "ya_mcts_select": [_vp, _vp, _i64, _vp, _vp, _vp, _u64, _u64, _u32, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _int, _vp, _vp,
                       _vp, _vp],
    "ya_mcts_select_injected": [_vp, _vp, _i64, _vp, _u32, ctypes.c_float, _vp, _int, _vp, _vp, _vp, _vp, _vp],
    "ya_mcts_expand": [_vp, _vp, _vp, _int, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp],
    "ya_mcts_expand_logits": [_vp, _vp, _int, _i64, _vp, _vp, _vp, _vp, _vp],
    "ya_mcts_search_uniform": [_vp, _vp, _i64, _vp, _vp, _vp, _u64, _u64, _int, ctypes.c_float, ctypes.c_float,
                               ctypes.c_float, _vp, _vp, _vp],
    "ya_mcts_root_counts": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "ya_mcts_root_sparse": [_vp, _vp, _i64, _vp, _int, _vp, _vp, _vp, _vp],
    "ya_mcts_pick_action": [_vp, _vp, _vp, _i64, _u64, _u64, _int, _vp, _vp],
    "ya_nn_forward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _i64, ctypes.c_float, _int, _vp, _vp, _vp],
    "ya_nn_forward_tiles": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _i64, ctypes.c_float, _int, _vp, _vp, _int, _vp],
    "ya_host_create": [_i64, _int, ctypes.POINTER(ctypes.c_void_p)],
    "ya_host_destroy": [_vp],
    "ya_host_play_ply": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u64, _int],
    "ya_host_play_ply_records": [_vp, _vp, _vp, _vp, _u64, _u64, _int],
    "ya_host_play_plies_records": [_vp, _vp, _int, _vp, _vp, _u64, _u64, _int],
    "ya_play_ply_records": [_vp, _vp, _vp, _i64, _u64, _u64, _int, _vp],
}



class MctsTreeStruct(ctypes.Structure):
    """ya_mcts_tree of include/yacht_b200.h."""
    _fields_ = [("nodes", _vp), ("ht", _vp), ("arena", _vp), ("meta", _vp), ("cursor", _vp),
                ("n", _i64), ("max_nodes", ctypes.c_int32), ("ht_size", ctypes.c_int32), ("arena_words", _i64)]


_lib = None


class YachtB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise YachtB200Error(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        raise YachtB200Error("%s failed with CUDA error %d" % (what, code))


def ptr(t):
    """Device (or pinned-host) pointer of a torch tensor, or NULL."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
