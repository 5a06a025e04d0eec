"""ctypes wrapper of oracle/pair_count.c (CPU restatement of the reference's pair counting).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpaircount_oracle.so")


def _load():
    if not os.path.exists(_SO):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    h = C.CDLL(_SO)
    h.pair_count_oracle.restype = C.c_long
    h.pair_count_oracle.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    return h


def count_pairs_c(data) -> dict:
    """bytes / uint8 ndarray -> {(a, b): count}, single thread."""
    arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    asc = np.zeros(128 * 128, dtype=np.uint64)
    cap = 1 << 20
    keys = np.zeros(cap, dtype=np.uint64)
    vals = np.zeros(cap, dtype=np.uint64)
    n = _load().pair_count_oracle(arr.ctypes.data, arr.size, asc.ctypes.data, keys.ctypes.data, vals.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("too many distinct non-ASCII pairs")
    out = {(chr(k >> 7), chr(k & 127)): int(asc[k]) for k in np.nonzero(asc)[0].tolist()}
    for k, v in zip(keys[:n].tolist(), vals[:n].tolist()):
        out[(chr(k >> 32), chr(k & 0xffffffff))] = int(v)
    return out
