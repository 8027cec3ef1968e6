"""ctypes loader of the plain-C oracle (oracle/c/spg_oracle.c).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libspg_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.spgo_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().spgo_num_threads()


def use_all_cores():
    """All host cores for the OpenMP loops, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    lib().spgo_set_threads(C.c_int(os.cpu_count() or 1))
    return num_threads()


def mul_batch(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    lib().spgo_mul_batch(_p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def ntt(data, log_n, inverse=False, order=2):
    arr = np.array(data, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    n = 1 << log_n
    lib().spgo_ntt(_p(arr), C.c_uint(log_n), C.c_size_t(arr.shape[0] // n), C.c_int(int(inverse)), C.c_int(order))
    return arr


def lde(trace, log_n, n_cols, log_blowup, offset=3):
    """trace (n_cols * 2^log_n, 4) -> (2^log_blowup * n_cols * 2^log_n, 4), layout [B][C][N]."""
    tr = np.ascontiguousarray(trace, dtype=np.uint64).reshape(-1, 4)
    n = 1 << log_n
    assert tr.shape[0] == n_cols * n
    out = np.empty(((n_cols * n) << log_blowup, 4), dtype=np.uint64)
    g = np.frombuffer(int(offset).to_bytes(32, "little"), dtype="<u8").copy()
    lib().spgo_lde(_p(tr), C.c_uint(log_n), C.c_size_t(n_cols), C.c_uint(log_blowup), _p(g), _p(out))
    return out


def poly_eval(coefs, n, points):
    """coefs (batch * n, 4) canonical natural-order coefficients, points: list of ints (one per polynomial) -> list of ints"""
    cf = np.ascontiguousarray(coefs, dtype=np.uint64).reshape(-1, 4)
    batch = cf.shape[0] // n
    assert len(points) == batch
    pts = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in points), dtype="<u8").reshape(-1, 4).copy()
    out = np.empty((batch, 4), dtype=np.uint64)
    lib().spgo_poly_eval(_p(cf), C.c_size_t(n), _p(pts), C.c_size_t(batch), _p(out))
    raw = out.tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(batch)]
