"""ctypes binding of oracle/_build/liboracle.so (the C restatement of the reference's CPU path).
TEST INFRASTRUCTURE ONLY -- see the header of oracle/ref_cpu.c."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = ctypes.CDLL(LIB)
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.oracle_bucket_msm.restype = ci
        L.oracle_bucket_msm.argtypes = [vp, sz, vp, sz, sz, sz, vp]
        L.oracle_bucket_msm_mt.restype = ci
        L.oracle_bucket_msm_mt.argtypes = [vp, sz, vp, sz, sz, sz, ci, vp]
        for name in ("oracle_ntt_381", "oracle_i_ntt_381"):
            getattr(L, name).restype = ci
            getattr(L, name).argtypes = [vp, vp, sz]
        L.oracle_ntt_381_rows.restype = ci
        L.oracle_ntt_381_rows.argtypes = [vp, vp, sz, ci, sz, sz]
        for name in ("oracle_fp_mul", "oracle_fr_mul", "oracle_fp_add", "oracle_fp_sub", "oracle_g1_add"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, vp, vp]
        L.oracle_g1_double.restype = None
        L.oracle_g1_double.argtypes = [vp, vp]
        L.oracle_fr_horner.restype = None
        L.oracle_fr_horner.argtypes = [vp, sz, vp, vp]
        _lib = L
    return _lib


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def bucket_msm(points_xyz, scalars, b=256, c=4, threads=1):
    """points uint64[n,18], scalars uint64[m,4] (Montgomery) -> uint64[18] projective (not normalised)"""
    p, s = _u64(points_xyz).reshape(-1, 18), _u64(scalars).reshape(-1, 4)
    out = np.zeros(18, dtype=np.uint64)
    if threads == 1:
        rc = lib().oracle_bucket_msm(p.ctypes.data, p.shape[0], s.ctypes.data, s.shape[0], b, c, out.ctypes.data)
    else:
        rc = lib().oracle_bucket_msm_mt(p.ctypes.data, p.shape[0], s.ctypes.data, s.shape[0], b, c, threads,
                                        out.ctypes.data)
    if rc != 0:
        raise IndexError("the reference panics on these (b, c)")
    return out


def ntt_381(elements, inverse=False):
    e = _u64(elements).reshape(-1, 4)
    out = np.zeros_like(e)
    fn = lib().oracle_i_ntt_381 if inverse else lib().oracle_ntt_381
    if fn(e.ctypes.data, out.ctypes.data, e.shape[0]) != 0:
        raise AssertionError("assertion failed: is_power_of_two(n)")
    return out


def ntt_381_rows(elements, row_lo, row_hi, inverse=False):
    e = _u64(elements).reshape(-1, 4)
    out = np.zeros_like(e)
    rc = lib().oracle_ntt_381_rows(e.ctypes.data, out.ctypes.data, e.shape[0], 1 if inverse else 0, row_lo, row_hi)
    assert rc == 0
    return out[row_lo:row_hi]


def fr_horner(scalars, tau_mont):
    s = _u64(scalars).reshape(-1, 4)
    t = _u64(tau_mont).reshape(4)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_horner(s.ctypes.data, s.shape[0], t.ctypes.data, out.ctypes.data)
    return out


def binop(name, a, b, n):
    a, b = _u64(a), _u64(b)
    out = np.zeros(n, dtype=np.uint64)
    getattr(lib(), name)(a.ctypes.data, b.ctypes.data, out.ctypes.data)
    return out


def g1_iota(n):
    """[1]G .. [n]G as uint64[n,18] projective points (not normalised)"""
    L = lib()
    L.oracle_g1_iota.restype = None
    L.oracle_g1_iota.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    out = np.zeros((n, 18), dtype=np.uint64)
    L.oracle_g1_iota(n, out.ctypes.data)
    return out
