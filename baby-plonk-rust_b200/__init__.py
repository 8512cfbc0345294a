"""baby-plonk-rust_b200: B200-native MSM / NTT hot path of ChainUpZero/baby-plonk-rust.

This package is the host-side mirror of the reference's call surfaces for the hot path
(SURVEY.md section 8b), in Python because the reference's Rust toolchain is not available in
this image; the Rust FFI shim a maintainer would use instead is in INTEGRATION.md.  Everything
here is a thin layer over the C ABI of ``libbpk.so`` (include/bpk.h):

    reference (Rust)                                   here
    -------------------------------------------------  ---------------------------------------
    BucketMSM::bucket_msm(points, scalars, b, c)       BucketMSM.bucket_msm(points, scalars, b, c)
      src/msm.rs:76-118
    Setup::generate_srs / Setup::commit                Setup.generate_srs / Setup.commit
      src/setup.rs:12-37
    ntt_381 / i_ntt_381                                ntt_381 / i_ntt_381
      src/utils.rs:63-81, 106-129
    root_of_unity / roots_of_unity /                   same names
      find_next_power_of_two  src/utils.rs:39-61
    Polynomial { values, basis }, ntt, i_ntt, Mul      Polynomial
      src/polynomial.rs:14-55, 189-276

Data layout is the reference's in-memory layout: a Scalar is 4 little-endian u64 limbs in
Montgomery form (numpy ``uint64[n, 4]``), a G1Projective is 18 u64 limbs X|Y|Z (``uint64[n, 18]``).
Rust panics (assert!, unwrap) surface as ``BpkPanic``.

There is NO CPU fallback: if ``libbpk.so`` is missing or no sm_100 GPU is present, every call
raises.  Nothing in this package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPK_LIB") or os.path.join(_HERE, "libbpk.so")  # BPK_LIB: A/B kernel builds

# ---------------------------------------------------------------------------------------------
# field constants needed by the host layer (value conversion only; no data-path arithmetic)
# ---------------------------------------------------------------------------------------------
FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001   # scalar.rs:83-88
FP_MODULUS = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_FR_R = (1 << 256) % FR_MODULUS
_FR_RINV = pow(_FR_R, -1, FR_MODULUS)
_FP_R = (1 << 384) % FP_MODULUS
_FP_RINV = pow(_FP_R, -1, FP_MODULUS)
_ROOT_OF_UNITY = pow(7, (FR_MODULUS - 1) >> 32, FR_MODULUS)                     # scalar.rs:201-213
_MASK64 = (1 << 64) - 1


class BpkPanic(RuntimeError):
    """A condition on which the reference panics (assert!, slice bounds, unwrap) or a CUDA failure."""


# ---------------------------------------------------------------------------------------------
# library loading
# ---------------------------------------------------------------------------------------------
_lib = None

_SIGNATURES = {
    "bpk_init": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "bpk_destroy": (None, [ctypes.c_void_p]),
    "bpk_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "bpk_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "bpk_abi_version": (ctypes.c_int, []),
    "bpk_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "bpk_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "bpk_srs_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]),
    "bpk_srs_generate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]),
    "bpk_srs_generate_range": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                              ctypes.POINTER(ctypes.c_uint64)]),
    "bpk_srs_precompute": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint]),
    "bpk_srs_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_srs_len": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_size_t)]),
    "bpk_srs_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64]),
    "bpk_srs_table_bytes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_size_t)]),
    "bpk_bucket_msm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_msm_g1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_msm_g1_points": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                         ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_msm_g1_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "bpk_msm_g1_from_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "bpk_msm_g1_dev_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "bpk_g1_sum": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_g1_sum_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_ntt_fr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]),
    "bpk_intt_fr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]),
    "bpk_coset_ntt_fr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_coset_intt_fr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_ntt_fr_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "bpk_poly_mul_fr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_poly_mul_fr_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.c_void_p]),
    "bpk_dev_alloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "bpk_dev_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "bpk_dev_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "bpk_dev_download": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "bpk_dev_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "bpk_dev_zero": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "bpk_fr_vec_op": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_size_t]),
    "bpk_fr_scale_powers": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_size_t]),
    "bpk_fr_poly_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "bpk_fr_poly_eval_many": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "bpk_fr_poly_div_linear": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "bpk_fr_poly_div_vanishing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                                 ctypes.c_void_p]),
    "bpk_plonk_grand_product": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 6 + [ctypes.c_size_t] +
                                [ctypes.c_void_p] * 5),
    "bpk_plonk_quotient_evals": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                ctypes.c_size_t] + [ctypes.c_void_p] * 7),
    "bpk_fr_fold": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                   ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "bpk_plonk_quotient_evals_shard": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                      ctypes.c_uint] + [ctypes.c_void_p] * 7),
    "bpk_keccak_f1600": (None, [ctypes.c_void_p]),
    "bpk_synthetic_chain_circuit": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_void_p,
                                                   ctypes.c_void_p]),
    "bpk_profile_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "bpk_profile_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "bpk_profile_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_uint64)]),
    "bpk_launch_count": (ctypes.c_uint64, [ctypes.c_void_p]),
    "bpk_msm_last_plan": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint)]),
    "bpk_msm_last_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "bpk_imad_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "bpk_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_long]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def load_library() -> ctypes.CDLL:
    """dlopen libbpk.so and bind every entry point of include/bpk.h.  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BpkPanic(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(there is no CPU fallback for the MSM / NTT path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# ---------------------------------------------------------------------------------------------
# value conversion helpers (host): Python ints <-> Montgomery limb arrays
# ---------------------------------------------------------------------------------------------
def scalars_from_ints(values: Iterable[int]) -> np.ndarray:
    """canonical integers -> uint64[n, 4] Montgomery limbs (what Vec<Scalar> holds)"""
    raw = b"".join(((v % FR_MODULUS) * _FR_R % FR_MODULUS).to_bytes(32, "little") for v in values)
    return np.frombuffer(raw, dtype="<u8").reshape(-1, 4).astype(np.uint64)


def scalars_to_ints(arr: np.ndarray) -> list:
    raw = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 4).tobytes()
    return [int.from_bytes(raw[k:k + 32], "little") * _FR_RINV % FR_MODULUS for k in range(0, len(raw), 32)]


def _fp_limbs(v: int) -> list:
    m = (v % FP_MODULUS) * _FP_R % FP_MODULUS
    return [(m >> (64 * k)) & _MASK64 for k in range(6)]


def _fp_value(limbs) -> int:
    m = 0
    for k, l in enumerate(limbs):
        m |= int(l) << (64 * k)
    return m * _FP_RINV % FP_MODULUS


def points_from_affine(points: Sequence[Optional[tuple]], z_scale: Optional[Sequence[int]] = None) -> np.ndarray:
    """affine (x, y) tuples / None -> uint64[n, 18] G1Projective limbs.  With z_scale the points are
    handed over as the non-normalised representatives (x z, y z, z) a Rust scalar multiplication leaves."""
    out = np.empty((len(points), 18), dtype=np.uint64)
    for i, pt in enumerate(points):
        z = 1 if z_scale is None else z_scale[i]
        if pt is None:
            out[i] = _fp_limbs(0) + _fp_limbs(z) + _fp_limbs(0)
        else:
            out[i] = _fp_limbs(pt[0] * z) + _fp_limbs(pt[1] * z) + _fp_limbs(z)
    return out


def point_to_affine(xyz: np.ndarray) -> Optional[tuple]:
    """G1Affine::from on an 18-limb projective point: (x, y) canonical ints, None for the identity"""
    xyz = np.asarray(xyz, dtype=np.uint64).reshape(18)
    X, Y, Z = _fp_value(xyz[0:6]), _fp_value(xyz[6:12]), _fp_value(xyz[12:18])
    if Z == 0:
        return None
    zi = pow(Z, -1, FP_MODULUS)
    return (X * zi % FP_MODULUS, Y * zi % FP_MODULUS)


def point_to_compressed(xyz: np.ndarray) -> bytes:
    """G1Affine::to_compressed (g1.rs:221-242) of an 18-limb projective point"""
    pt = point_to_affine(xyz)
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    b = bytearray(pt[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if pt[1] > (FP_MODULUS - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def _as_u64(a, cols: int) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.uint64)
    if arr.ndim == 1 and cols and arr.size % cols == 0:
        arr = arr.reshape(-1, cols)
    if arr.ndim != 2 or arr.shape[1] != cols:
        raise ValueError(f"expected uint64[n, {cols}]")
    return arr


# ---------------------------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------------------------
class Context:
    """One bpk_ctx (one GPU).  Owns device tables, workspaces and the SRS cache."""

    def __init__(self, device: Optional[int] = None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = ctypes.c_void_p()
        st = self.lib.bpk_init(ctypes.byref(h), int(device))
        if st != 0:
            raise BpkPanic(f"bpk_init(device={device}) failed: {self.lib.bpk_strerror(st).decode()}")
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.bpk_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, status: int, what: str = ""):
        if status != 0:
            msg = self.lib.bpk_strerror(status).decode()
            detail = self.lib.bpk_last_error(self.handle).decode() if status == -2 else ""
            raise BpkPanic(f"{what}: {msg} {detail}".strip())

    # -- instrumentation --
    def set_option(self, key: str, value: int):
        self.check(self.lib.bpk_set_option(self.handle, key.encode(), int(value)), "bpk_set_option")

    def profile_enable(self, on: bool = True):
        self.check(self.lib.bpk_profile_enable(self.handle, 1 if on else 0))

    def profile_reset(self):
        self.check(self.lib.bpk_profile_reset(self.handle))

    def profile_get(self, name: str):
        ms = ctypes.c_double()
        n = ctypes.c_uint64()
        self.check(self.lib.bpk_profile_get(self.handle, name.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.bpk_launch_count(self.handle))

    def msm_last_plan(self) -> dict:
        arr = (ctypes.c_uint * 4)()
        self.check(self.lib.bpk_msm_last_plan(self.handle, arr), "bpk_msm_last_plan")
        return {"window_bits": arr[0], "windows": arr[1], "pairs_per_thread": arr[2], "buckets": arr[3]}

    def msm_last_stats(self) -> dict:
        """counters of the most recent MSM (synchronises): how its additions split between the batched-affine
        pairwise tree and the XYZZ tail"""
        arr = (ctypes.c_uint64 * 6)()
        self.check(self.lib.bpk_msm_last_stats(self.handle, arr), "bpk_msm_last_stats")
        return {"entries": int(arr[0]), "affine_adds": int(arr[1]), "xyzz_adds_bound": int(arr[2]),
                "nonempty_buckets": int(arr[3]), "tree_levels": int(arr[4]), "batch": int(arr[5])}

    def synchronize(self):
        self.check(self.lib.bpk_synchronize(self.handle), "bpk_synchronize")

    def set_stream(self, cuda_stream: int):
        """run all later work of this context on the given cudaStream_t (0 = the legacy default stream)"""
        if getattr(self, "_stream", 0) != int(cuda_stream):
            self.check(self.lib.bpk_set_stream(self.handle, ctypes.c_void_p(int(cuda_stream))), "bpk_set_stream")
            self._stream = int(cuda_stream)

    def bind_torch_stream(self, torch):
        """Callers that interleave torch ops (copies, NCCL collectives) with this library's kernels must keep both on
        ONE stream: torch's current stream.  Outside `torch.cuda.stream(...)` that is the legacy default stream, which
        is also the context's default, so this is a no-op then; inside, the context follows torch."""
        s = torch.cuda.current_stream(self.device)
        self.set_stream(0 if s == torch.cuda.default_stream(self.device) else s.cuda_stream)

    def imad_peak(self, mode: int = 0):
        self.set_option("imad.mode", mode)
        rate = ctypes.c_double()
        sec = ctypes.c_double()
        self.check(self.lib.bpk_imad_peak(self.handle, ctypes.byref(rate), ctypes.byref(sec)), "bpk_imad_peak")
        return rate.value, sec.value


_default_ctx: Optional[Context] = None


def get_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# ---------------------------------------------------------------------------------------------
# src/utils.rs host helpers
# ---------------------------------------------------------------------------------------------
def is_power_of_two(n: int) -> bool:
    """utils.rs:82-84"""
    return n != 0 and (n & (n - 1)) == 0


def root_of_unity(group_order: int) -> int:
    """utils.rs:39-43 (canonical value)"""
    return pow(_ROOT_OF_UNITY, (1 << 32) // group_order, FR_MODULUS)


def roots_of_unity(group_order: int) -> list:
    """utils.rs:45-52"""
    g = root_of_unity(group_order)
    res = [1]
    for _ in range(1, group_order):
        res.append(res[-1] * g % FR_MODULUS)
    return res


def find_next_power_of_two(n: int, m: int) -> int:
    """utils.rs:54-61"""
    power = 1
    target = n + m + 1
    while power < target:
        power <<= 1
    return power


def _ntt_call(fn_name: str, elements: np.ndarray, ctx: Optional[Context], shift=None) -> np.ndarray:
    ctx = ctx or get_context()
    arr = np.ascontiguousarray(elements, dtype=np.uint64)
    batched = arr.ndim == 3
    a = arr if batched else arr.reshape(1, -1, 4)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("expected uint64[n, 4] or uint64[batch, n, 4]")
    batch, n = a.shape[0], a.shape[1]
    # assert!(is_power_of_two(n))  utils.rs:65,108
    if not is_power_of_two(n):
        raise BpkPanic("assertion failed: is_power_of_two(n)")
    out = np.empty_like(a)
    fn = getattr(ctx.lib, fn_name)
    if shift is None:
        st = fn(ctx.handle, a.ctypes.data, out.ctypes.data, n, batch)
    else:
        sh = scalars_from_ints([shift]) if isinstance(shift, int) else np.ascontiguousarray(shift, dtype=np.uint64)
        st = fn(ctx.handle, a.ctypes.data, out.ctypes.data, n, batch, sh.ctypes.data)
    ctx.check(st, fn_name)
    return out if batched else out.reshape(n, 4)


def ntt_381(elements: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """utils.rs:63-81: coefficients -> evaluations on {w^i}; uint64[n,4] (or [batch,n,4]) Montgomery"""
    return _ntt_call("bpk_ntt_fr", elements, ctx)


def i_ntt_381(elements: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """utils.rs:106-129: evaluations -> coefficients (scaled by n^-1)"""
    return _ntt_call("bpk_intt_fr", elements, ctx)


def coset_ntt(elements: np.ndarray, shift, ctx: Optional[Context] = None) -> np.ndarray:
    """evaluate on shift * w^i (quotient pipeline; no reference counterpart)"""
    return _ntt_call("bpk_coset_ntt_fr", elements, ctx, shift)


def coset_intt(elements: np.ndarray, shift, ctx: Optional[Context] = None) -> np.ndarray:
    return _ntt_call("bpk_coset_intt_fr", elements, ctx, shift)


# ---------------------------------------------------------------------------------------------
# src/polynomial.rs
# ---------------------------------------------------------------------------------------------
class Basis:
    Monomial = "Monomial"
    Lagrange = "Lagrange"


class Polynomial:
    """polynomial.rs:14-17.  values: uint64[n, 4] Montgomery limbs."""

    def __init__(self, values, basis: str = Basis.Monomial, ctx: Optional[Context] = None):
        self.values = _as_u64(values, 4)
        self.basis = basis
        self._ctx = ctx

    @classmethod
    def from_ints(cls, ints: Iterable[int], basis: str = Basis.Monomial, ctx: Optional[Context] = None):
        return cls(scalars_from_ints(ints), basis, ctx)

    def to_ints(self) -> list:
        return scalars_to_ints(self.values)

    def ntt(self) -> "Polynomial":
        """polynomial.rs:47-51"""
        if self.basis != Basis.Monomial:
            raise BpkPanic("assertion failed: self.basis == Basis::Monomial")
        return Polynomial(ntt_381(self.values, self._ctx), Basis.Lagrange, self._ctx)

    def i_ntt(self) -> "Polynomial":
        """polynomial.rs:52-55"""
        if self.basis != Basis.Lagrange:
            raise BpkPanic("assertion failed: self.basis == Basis::Lagrange")
        return Polynomial(i_ntt_381(self.values, self._ctx), Basis.Monomial, self._ctx)

    def __mul__(self, rhs: "Polynomial") -> "Polynomial":
        """impl Mul for Polynomial (polynomial.rs:189-276), Monomial x Monomial"""
        if self.basis != rhs.basis:
            raise BpkPanic("assertion failed: self.basis == rhs.basis")
        if self.basis == Basis.Lagrange:
            raise BpkPanic("not yet implemented")  # todo!() polynomial.rs:274-276
        ctx = self._ctx or get_context()
        la, lb = self.values.shape[0], rhs.values.shape[0]
        out = np.empty((la + lb - 1, 4), dtype=np.uint64)
        st = ctx.lib.bpk_poly_mul_fr(ctx.handle, self.values.ctypes.data, la, rhs.values.ctypes.data, lb,
                                     out.ctypes.data)
        ctx.check(st, "bpk_poly_mul_fr")
        return Polynomial(out, Basis.Monomial, self._ctx)


# ---------------------------------------------------------------------------------------------
# src/msm.rs, src/setup.rs
# ---------------------------------------------------------------------------------------------
class BucketMSM:
    @staticmethod
    def bucket_msm(points: np.ndarray, scalars: np.ndarray, b: int = 256, c: int = 4,
                   ctx: Optional[Context] = None) -> np.ndarray:
        """msm.rs:76-118 with the reference signature: points uint64[n,18], scalars uint64[m,4].
        Returns the normalised G1Projective limbs (uint64[18])."""
        ctx = ctx or get_context()
        pts = _as_u64(points, 18)
        sc = _as_u64(scalars, 4)
        n = min(pts.shape[0], sc.shape[0])
        handle = ctypes.c_uint64()
        ctx.check(ctx.lib.bpk_srs_load(ctx.handle, pts.ctypes.data, n, ctypes.byref(handle)), "bpk_srs_load")
        try:
            out = np.empty(18, dtype=np.uint64)
            st = ctx.lib.bpk_bucket_msm(ctx.handle, handle.value, sc.ctypes.data, sc.shape[0], b, c, out.ctypes.data)
            ctx.check(st, "bpk_bucket_msm")
        finally:
            ctx.lib.bpk_srs_free(ctx.handle, handle.value)
        return out


class Setup:
    """setup.rs:7-37.  powers_of_x stay resident on the GPU behind an SRS handle."""

    def __init__(self, ctx: Context, handle: int, n: int):
        self.ctx = ctx
        self.handle = handle
        self.n = n

    @classmethod
    def generate_srs(cls, powers: int, tau: int, ctx: Optional[Context] = None, first: int = 0) -> "Setup":
        """setup.rs:12-31: [tau^i]G for first <= i < first + powers (first != 0: the shard one rank of a
        multi-GPU job owns).  The G2 part stays with the CPU verifier (out of scope)."""
        ctx = ctx or get_context()
        t = scalars_from_ints([tau])
        h = ctypes.c_uint64()
        ctx.check(ctx.lib.bpk_srs_generate_range(ctx.handle, t.ctypes.data, first, powers, ctypes.byref(h)),
                  "bpk_srs_generate_range")
        return cls(ctx, h.value, powers)

    @classmethod
    def from_points(cls, points: np.ndarray, ctx: Optional[Context] = None) -> "Setup":
        ctx = ctx or get_context()
        pts = _as_u64(points, 18)
        h = ctypes.c_uint64()
        ctx.check(ctx.lib.bpk_srs_load(ctx.handle, pts.ctypes.data, pts.shape[0], ctypes.byref(h)), "bpk_srs_load")
        return cls(ctx, h.value, pts.shape[0])

    def precompute(self, window_bits: int = 0) -> "Setup":
        """keep [2^(c w)] P_i for every window next to the SRS (W x its size in HBM): later commits use one
        shared bucket set, fewer windows, no final doubling chain.  Results are bit-identical."""
        self.ctx.check(self.ctx.lib.bpk_srs_precompute(self.ctx.handle, self.handle, int(window_bits)),
                       "bpk_srs_precompute")
        return self

    def table_bytes(self) -> int:
        """HBM bytes of the point table behind this Setup (W x the SRS after precompute)"""
        out = ctypes.c_size_t()
        self.ctx.check(self.ctx.lib.bpk_srs_table_bytes(self.ctx.handle, self.handle, ctypes.byref(out)),
                       "bpk_srs_table_bytes")
        return int(out.value)

    def powers_of_x(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.empty((count, 18), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.bpk_srs_read(self.ctx.handle, self.handle, first, count, out.ctypes.data),
                       "bpk_srs_read")
        return out

    def commit(self, polynomial: Polynomial) -> np.ndarray:
        """setup.rs:32-37"""
        if polynomial.basis != Basis.Monomial:
            raise BpkPanic("assertion `left == right` failed: polynomial.basis == Basis::Monomial")
        return self.commit_scalars(polynomial.values)

    def commit_scalars(self, scalars: np.ndarray, b: int = 256, c: int = 4) -> np.ndarray:
        sc = _as_u64(scalars, 4)
        out = np.empty(18, dtype=np.uint64)
        st = self.ctx.lib.bpk_bucket_msm(self.ctx.handle, self.handle, sc.ctypes.data, sc.shape[0], b, c,
                                         out.ctypes.data)
        self.ctx.check(st, "bpk_bucket_msm")
        return out

    def free(self):
        if self.handle:
            self.ctx.lib.bpk_srs_free(self.ctx.handle, self.handle)
            self.handle = 0


def g1_sum(points: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """sum of projective points (the post-gather step of the sharded MSM)"""
    ctx = ctx or get_context()
    pts = _as_u64(points, 18)
    out = np.empty(18, dtype=np.uint64)
    ctx.check(ctx.lib.bpk_g1_sum(ctx.handle, pts.ctypes.data, pts.shape[0], out.ctypes.data), "bpk_g1_sum")
    return out


__all__ = [
    "BpkPanic", "Context", "get_context", "load_library", "LIB_PATH", "EXPORTED_SYMBOLS",
    "BucketMSM", "Setup", "Polynomial", "Basis",
    "ntt_381", "i_ntt_381", "coset_ntt", "coset_intt",
    "root_of_unity", "roots_of_unity", "find_next_power_of_two", "is_power_of_two",
    "scalars_from_ints", "scalars_to_ints", "points_from_affine", "point_to_affine", "point_to_compressed",
    "g1_sum", "FR_MODULUS", "FP_MODULUS",
]
