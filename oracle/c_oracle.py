"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): ctypes wrapper of oracle/yacht_oracle.c."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libyacht_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "yacht_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        lib = ctypes.CDLL(LIB)
        lib.yo_play_random.restype = ctypes.c_longlong
        lib.yo_play_random.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int]
        lib.yo_category_points.restype = ctypes.c_int
        lib.yo_category_points.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        lib.yo_subset_position.restype = ctypes.c_int
        lib.yo_subset_position.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.yo_philox.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


def philox(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    load().yo_philox(c, k, o)
    return tuple(int(x) for x in o)


def category_points(cat, five):
    return load().yo_category_points(cat, (ctypes.c_int * 5)(*five))


def play_random(n, plies, seed, game_base=0, auto_reset=False, want_keys=False):
    """Returns dict(packed uint32[plies,n,8], actions int32[plies,n], legal int32[plies,n], result int8[plies,n],
    steps, keys)."""
    packed = np.zeros((plies, n, 8), dtype=np.uint32)
    actions = np.zeros((plies, n), dtype=np.int32)
    legal = np.zeros((plies, n), dtype=np.int32)
    result = np.zeros((plies, n), dtype=np.int8)
    cap = 256
    keys = ctypes.create_string_buffer(n * cap) if want_keys else None
    steps = load().yo_play_random(n, plies, seed, game_base, 1 if auto_reset else 0, packed.ctypes.data, actions.ctypes.data,
                                  legal.ctypes.data, result.ctypes.data, ctypes.addressof(keys) if keys else None, cap)
    out = {"packed": packed, "actions": actions, "legal": legal, "result": result, "steps": int(steps)}
    if want_keys:
        out["keys"] = [keys.raw[i * cap:(i + 1) * cap].split(b"\0", 1)[0].decode() for i in range(n)]
    return out


def timed_steps(n, plies, seed=0):
    """Plays n games x plies on all host cores (OpenMP) without recording anything; returns steps."""
    return int(load().yo_play_random(n, plies, seed, 0, 1, None, None, None, None, None, 0))
