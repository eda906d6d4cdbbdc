"""ctypes wrapper of oracle/fps_oracle.c (TEST INFRASTRUCTURE).  Built by `make -C oracle` /
`__graft_entry__.build()` into oracle/_build/liboracle.so."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", _HERE, "c_oracle"], stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(_SO)
        _lib.fps_oracle.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5
        _lib.fps_oracle.restype = None
        _lib.fps_oracle_block_size.argtypes = [ctypes.c_int]
        _lib.fps_oracle_block_size.restype = ctypes.c_int
    return _lib


def block_size(n_max: int) -> int:
    return _load().fps_oracle_block_size(int(n_max))


def furthestsampling(xyz, offset, new_offset):
    """xyz [N,3] f32, offset/new_offset cumulative [b] -> idx int32 [new_offset[-1]]
    (lib/pointops2/functions/pointops.py:14-31 semantics)."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    offset = np.ascontiguousarray(offset, np.int32)
    new_offset = np.ascontiguousarray(new_offset, np.int32)
    sizes = np.diff(np.concatenate([[0], offset]))
    tmp = np.empty(xyz.shape[0], np.float32)
    idx = np.zeros(int(new_offset[-1]), np.int32)
    _load().fps_oracle(offset.shape[0], int(sizes.max()), xyz.ctypes.data, offset.ctypes.data,
                       new_offset.ctypes.data, tmp.ctypes.data, idx.ctypes.data)
    return idx


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """-> (idx int32 [m,nsample], dist2 float32 [m,nsample]) exactly as the reference kernel produces them."""
    lib = _load()
    lib.knn_oracle.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6
    lib.knn_oracle.restype = None
    xyz = np.ascontiguousarray(xyz, np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    offset = np.ascontiguousarray(offset, np.int32)
    new_offset = np.ascontiguousarray(new_offset, np.int32)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), np.int32)
    d2 = np.zeros((m, nsample), np.float32)
    lib.knn_oracle(m, offset.shape[0], nsample, xyz.ctypes.data, new_xyz.ctypes.data, offset.ctypes.data,
                   new_offset.ctypes.data, idx.ctypes.data, d2.ctypes.data)
    return idx, d2
