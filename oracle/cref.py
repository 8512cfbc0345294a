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
        L.oracle_poly_mul.restype = ci
        L.oracle_poly_mul.argtypes = [vp, sz, vp, sz, vp]
        L.oracle_g1_powers_small.restype = None
        L.oracle_g1_powers_small.argtypes = [vp, ctypes.c_uint64, sz, vp]
        L.oracle_fr_ntt_fast.restype = ci
        L.oracle_fr_ntt_fast.argtypes = [vp, vp, sz, ci, ci]
        L.oracle_fr_poly_at.restype = ci
        L.oracle_fr_poly_at.argtypes = [vp, sz, vp, ci, vp]
        L.oracle_msm_pippenger_mt.restype = ci
        L.oracle_msm_pippenger_mt.argtypes = [vp, sz, vp, sz, ci, ci, vp]
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


def poly_mul(a, b):
    """impl Mul for Polynomial (Monomial) by the reference's own algorithm -> uint64[la + lb - 1, 4]"""
    a, b = _u64(a).reshape(-1, 4), _u64(b).reshape(-1, 4)
    out = np.zeros((a.shape[0] + b.shape[0] - 1, 4), dtype=np.uint64)
    rc = lib().oracle_poly_mul(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data)
    assert rc == 0
    return out


def g1_powers_small(start_xyz, k, n):
    """start, [k]start, [k^2]start, ... as uint64[n,18] projective points (not normalised)"""
    s = _u64(start_xyz).reshape(18)
    out = np.zeros((n, 18), dtype=np.uint64)
    lib().oracle_g1_powers_small(s.ctypes.data, k, n, out.ctypes.data)
    return out


def g1_iota(n):
    """[1]G .. [n]G as uint64[n,18] projective points (not normalised)"""
    L = lib()
    L.oracle_g1_iota.restype = None
    L.oracle_g1_iota.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    out = np.zeros((n, 18), dtype=np.uint64)
    L.oracle_g1_iota(n, out.ctypes.data)
    return out


# ---- optimised CPU baseline (oracle/fast_cpu.c): NOT the reference's algorithms -------------------------------
def ntt_fast(elements, inverse=False, threads=0):
    """radix-2 NTT on all host cores; same contract as ntt_381 / i_ntt_381 (natural order, Montgomery limbs)"""
    e = _u64(elements).reshape(-1, 4)
    out = np.empty_like(e)
    rc = lib().oracle_fr_ntt_fast(e.ctypes.data, out.ctypes.data, e.shape[0], 1 if inverse else 0,
                                  threads or os.cpu_count() or 1)
    if rc != 0:
        raise AssertionError("assertion failed: is_power_of_two(n)")
    return out


def poly_at(values, x_mont, threads=0):
    """p(x) for p given by its values on the n-th roots of unity (inverse NTT + Horner) -> uint64[4] Montgomery"""
    v, x = _u64(values).reshape(-1, 4), _u64(x_mont).reshape(4)
    out = np.zeros(4, dtype=np.uint64)
    rc = lib().oracle_fr_poly_at(v.ctypes.data, v.shape[0], x.ctypes.data, threads or os.cpu_count() or 1,
                                 out.ctypes.data)
    assert rc == 0
    return out


def msm_pippenger(points_xyz, scalars, threads=0, window=0):
    """signed-digit Pippenger on all host cores; points must be normalised (Z = R) or the identity"""
    p, s = _u64(points_xyz).reshape(-1, 18), _u64(scalars).reshape(-1, 4)
    out = np.zeros(18, dtype=np.uint64)
    rc = lib().oracle_msm_pippenger_mt(p.ctypes.data, p.shape[0], s.ctypes.data, s.shape[0],
                                       threads or os.cpu_count() or 1, window, out.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_msm_pippenger_mt: rc %d" % rc)
    return out
