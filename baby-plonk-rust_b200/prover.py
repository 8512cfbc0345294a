"""Device-resident PLONK prover: the caller of the MSM / NTT hot path, kept in HBM end to end.

Mirrors ``Prover::prove`` of the reference (src/prover.rs:106-176, rounds 1-5 at :177-647) with every
polynomial living on the GPU between the transforms and the commitments:

    reference step (serial Rust on Vec<Scalar>)            here (include/bpk.h)
    -----------------------------------------------------  ------------------------------------------
    i_ntt_381 of witness / selector / sigma columns        bpk_ntt_fr_dev (batched, inverse)
    Setup::commit x 9                                      bpk_msm_g1_dev on the resident SRS
    round 2 accumulator loop (prover.rs:286-317)           bpk_plonk_grand_product (ratio + product scan)
    round 3: 16 Polynomial::mul + Div by Z_H               coset transforms on a 4n domain +
      (prover.rs:370-452)                                    bpk_plonk_quotient_evals (one fused pass)
    coeffs_evaluate at zeta (prover.rs:502-541)            bpk_fr_poly_eval
    linearisation r(X), opening numerators                 bpk_fr_vec_op (a + s*b passes)
    Div by X - zeta, X - zeta*omega (prover.rs:623-638)    bpk_fr_poly_div_linear (scaled prefix scan)

The proof is the same group / field elements the reference computes: t(X) is formed from evaluations on
a coset instead of by long division, which yields the same polynomial whenever the division is exact
(it is for a satisfying witness; the reference drops the remainder otherwise and its own
``assert r(zeta) == 0`` -- kept here -- fires).  One documented difference: the reference's Div loses
interior zero quotient coefficients (polynomial.rs:314-380); with blinded polynomials that event has
probability ~ n / 2^255 and is not reproduced.

Only a few hundred bytes per round cross PCIe (commitments, evaluations, challenges).  Nothing here
imports ``oracle/``; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import hashlib
import time
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import (BpkPanic, Context, Setup, FR_MODULUS, _FR_RINV, is_power_of_two, point_to_compressed, root_of_unity,
               scalars_from_ints)
from .transcript import PlonkTranscript

Q = FR_MODULUS
COSET_SHIFT = 7           # multiplicative generator of Fr (scalar.rs GENERATOR): outside every 2-power subgroup
K1, K2 = 2, 3             # coset representatives of the permutation argument (prover.rs:286-317)


def _mont(v: int) -> np.ndarray:
    return scalars_from_ints([v])[0]


class _Scalars:
    """Montgomery limb arrays for a C call; holds the arrays so the pointers stay valid during the call"""

    def __init__(self, *values: int):
        self.arrays = [_mont(v) for v in values]

    def ptrs(self):
        return [a.ctypes.data for a in self.arrays]


def _from_mont(limbs) -> int:
    m = 0
    for k in range(4):
        m |= int(limbs[k]) << (64 * k)
    return m * _FR_RINV % Q


@dataclass
class Proof:
    """src/verifier.rs:23-40 field order; points as 48-byte compressed G1, scalars canonical ints"""
    a_1: bytes
    b_1: bytes
    c_1: bytes
    z_1: bytes
    t_lo_1: bytes
    t_mid_1: bytes
    t_hi_1: bytes
    w_zeta_1: bytes
    w_zeta_omega_1: bytes
    a_bar: int
    b_bar: int
    c_bar: int
    s1_bar: int
    s2_bar: int
    z_omega_bar: int

    POINTS = ("a_1", "b_1", "c_1", "z_1", "t_lo_1", "t_mid_1", "t_hi_1", "w_zeta_1", "w_zeta_omega_1")
    SCALARS = ("a_bar", "b_bar", "c_bar", "s1_bar", "s2_bar", "z_omega_bar")

    def to_bytes(self) -> bytes:
        out = b"".join(getattr(self, k) for k in self.POINTS)
        out += b"".join(int(getattr(self, k)).to_bytes(32, "little") for k in self.SCALARS)
        assert len(out) == 624
        return out

    def sha256(self) -> str:
        return hashlib.sha256(self.to_bytes()).hexdigest()


class DeviceProver:
    """Prover { group_order, setup, pk } (src/prover.rs:90-104) on one GPU.

    ``selectors`` = (QL, QR, QM, QO, QC) and ``sigmas`` = (S1, S2, S3) are the pre-processed Lagrange
    columns of src/program.rs:51-147 as uint64[n, 4] Montgomery limbs (what the reference's
    CommonPreprocessedInput holds).  They are uploaded once; their coefficient forms are recomputed in
    every ``prove`` exactly as the reference does unless ``cache_preprocessed`` is set.
    """

    def __init__(self, setup: Setup, group_order: int, selectors: Sequence[np.ndarray], sigmas: Sequence[np.ndarray],
                 cache_preprocessed: bool = False, committer=None, round3_shards: int = 0):
        import torch  # device memory only

        if not is_power_of_two(group_order):
            raise BpkPanic("assertion failed: is_power_of_two(group_order)")
        self.torch = torch
        self.setup = setup
        self.ctx: Context = setup.ctx
        self.lib = self.ctx.lib
        self.n = n = int(group_order)
        # multi-GPU: `committer` (multi_gpu.ShardedCommitter) owns this rank's slice of the SRS; every rank
        # runs the same prover on replicated polynomials and only the nine commitments are sharded
        self.committer = committer
        srs_len = setup.n if committer is None else self._committer_total(committer)
        if srs_len < n + 6:
            raise BpkPanic(f"SRS too short: {srs_len} powers for polynomials of {n + 6} coefficients")
        self.dev = torch.device("cuda", self.ctx.device)
        self.domain = 1
        while self.domain < 3 * n + 6:
            self.domain <<= 1
        self.ratio = self.domain // n
        self.stride = n + 8                       # room for the blinded degrees (n + 6 at most)
        self.omega = root_of_unity(n)
        cols = [np.ascontiguousarray(c, dtype=np.uint64).reshape(n, 4) for c in list(selectors) + list(sigmas)]
        self.pk_lagrange = self._upload(np.stack(cols))          # [8, n, 4]: ql qr qm qo qc s1 s2 s3
        self.cache_preprocessed = cache_preprocessed
        self._pre = None
        # Round 3 dealt over the ranks ("independent NTTs of a round are dealt to different GPUs"): the quotient
        # domain g<w_D> is the union of R sub-cosets (g w_D^j)<w_m>, m = D / R; rank j transforms the sixteen
        # polynomials of the round onto ITS sub-coset only (fold mod X^m - s, one size-m coset transform each),
        # evaluates t there, and the ranks exchange m x 32 bytes each.  Without a committer, round3_shards = R > 1 runs
        # the R parts one after the other on this GPU (same arithmetic, used by the single-GPU parity tests).
        world = 1 if committer is None else committer.world
        self.shards = int(round3_shards) if round3_shards else world
        if self.shards < 1 or self.shards & (self.shards - 1) or self.shards > self.domain:
            raise ValueError("round3_shards must be a power of two <= the quotient domain")
        if committer is not None and world > 1 and self.shards != world:
            raise ValueError("with a committer, round 3 is dealt over exactly its ranks")
        w = root_of_unity(self.domain)
        gn = pow(COSET_SHIFT, n, Q)
        wn = pow(w, n, Q)
        self._zh_inv = scalars_from_ints([pow((gn * pow(wn, i, Q) - 1) % Q, -1, Q) for i in range(self.ratio)])
        self._shift = _mont(COSET_SHIFT)
        self._one = _mont(1)

    @staticmethod
    def _committer_total(committer) -> int:
        import torch.distributed as dist

        if committer.world == 1:
            return committer.hi
        t = committer.d_partial.new_tensor([committer.hi])
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=committer.group)
        return int(t.item())

    # ---- plumbing -------------------------------------------------------------------------------
    def _upload(self, arr: np.ndarray):
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint64).view(np.int64))
        return t.to(self.dev, non_blocking=False)

    def _zeros(self, *shape):
        return self.torch.zeros(*shape, 4, dtype=self.torch.int64, device=self.dev)

    def _empty(self, *shape):
        return self.torch.empty(*shape, 4, dtype=self.torch.int64, device=self.dev)

    def _ck(self, st: int, what: str):
        self.ctx.check(st, what)

    def _vec(self, op: int, a, b, s, out, n: int):
        sp = None if s is None else s.ctypes.data   # s is a live array owned by the caller's frame
        self._ck(self.lib.bpk_fr_vec_op(self.ctx.handle, op, a.data_ptr(), 0 if b is None else b.data_ptr(), sp,
                                        out.data_ptr(), n), "bpk_fr_vec_op")

    def _axpy(self, acc, s: int, p, length: int):
        """acc[:length] += s * p[:length]"""
        s %= Q
        if s == 0:
            return
        sm = _mont(s)
        self._vec(4, acc, p, sm, acc, length)

    def _add_const(self, acc, index: int, s: int):
        sm = _mont(s % Q)
        self._vec(5, acc[index:index + 1], None, sm, acc[index:index + 1], 1)

    def _intt(self, src, dst, n: int, batch: int = 1):
        self._ck(self.lib.bpk_ntt_fr_dev(self.ctx.handle, src.data_ptr(), dst.data_ptr(), n, batch, 1, None),
                 "bpk_ntt_fr_dev")

    def _eval(self, coeffs, length: int, x: int) -> int:
        out = np.empty(4, dtype=np.uint64)
        xm = _mont(x)
        self._ck(self.lib.bpk_fr_poly_eval(self.ctx.handle, coeffs.data_ptr(), length, xm.ctypes.data,
                                           out.ctypes.data), "bpk_fr_poly_eval")
        return _from_mont(out)

    def _eval_many(self, items, x: int) -> list:
        """[(coeffs, length), ...] at one point: one pair of power tables, one device -> host copy"""
        k = len(items)
        ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t, _ in items])
        lens = (ctypes.c_size_t * k)(*[length for _, length in items])
        out = np.empty((k, 4), dtype=np.uint64)
        xm = _mont(x)
        self._ck(self.lib.bpk_fr_poly_eval_many(self.ctx.handle, k, ptrs, lens, xm.ctypes.data, out.ctypes.data),
                 "bpk_fr_poly_eval_many")
        return [_from_mont(out[i]) for i in range(k)]

    def _commit(self, coeffs, length: int) -> bytes:
        """Setup::commit (src/setup.rs:32-37) on device-resident coefficients -> compressed G1"""
        if self.committer is not None:
            xyz = self.committer.commit_prefix(coeffs, length).cpu().numpy().view(np.uint64)
            return point_to_compressed(xyz)
        self._ck(self.lib.bpk_msm_g1_dev(self.ctx.handle, self.setup.handle, 0, coeffs.data_ptr(), length, 1,
                                         self._pt.data_ptr()), "bpk_msm_g1_dev")
        xyz = self._pt.cpu().numpy().view(np.uint64)
        return point_to_compressed(xyz)

    def _commit_many(self, items) -> list:
        """the independent commitments of one round, [(coeffs, length), ...], in one batched call"""
        k = len(items)
        if self.committer is not None:
            xyz = self.committer.commit_prefix_many(items)
        else:
            ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t, _ in items])
            firsts = (ctypes.c_size_t * k)(*([0] * k))
            lens = (ctypes.c_size_t * k)(*[length for _, length in items])
            out = self.torch.empty((k, 18), dtype=self.torch.int64, device=self.dev)
            self._ck(self.lib.bpk_msm_g1_dev_batch(self.ctx.handle, self.setup.handle, k, ptrs, firsts, lens, 1,
                                                   out.data_ptr()), "bpk_msm_g1_dev_batch")
            xyz = out.cpu().numpy().view(np.uint64)
        return [point_to_compressed(xyz[i]) for i in range(k)]

    def _div_linear(self, coeffs, length: int, root: int, out):
        rm = _mont(root)
        self._ck(self.lib.bpk_fr_poly_div_linear(self.ctx.handle, coeffs.data_ptr(), length, rm.ctypes.data,
                                                 out.data_ptr()), "bpk_fr_poly_div_linear")

    # ---- round 3 on sub-cosets ---------------------------------------------------------------------
    def _subcoset_shift(self, j: int) -> int:
        return COSET_SHIFT * pow(root_of_unity(self.domain), j, Q) % Q

    def _subcoset_evals(self, rows, length: int, j: int):
        """rows: [k, stride, 4] coefficient rows (`length` coefficients each) -> their values on sub-coset j, [k, m, 4]:
        p mod (X^m - s^m) has p's values wherever X^m = s^m, so fold, then one size-m coset transform per row"""
        m = self.domain // self.shards
        k, stride = rows.shape[0], rows.shape[1]
        s = self._subcoset_shift(j)
        out = self._empty(k, m)
        sm = _mont(pow(s, m, Q))
        self._ck(self.lib.bpk_fr_fold(self.ctx.handle, rows.data_ptr(), k, stride, length, m, sm.ctypes.data,
                                      out.data_ptr()), "bpk_fr_fold")
        sh = _mont(s)
        self._ck(self.lib.bpk_ntt_fr_dev(self.ctx.handle, out.data_ptr(), out.data_ptr(), m, k, 2, sh.ctypes.data),
                 "bpk_ntt_fr_dev")
        return out

    def _zh_inv_subcoset(self, j: int):
        """1 / Z_H on sub-coset j: x^n = s^n (w_m^n)^i takes max(1, ratio / R) values"""
        n, m = self.n, self.domain // self.shards
        period = max(1, self.ratio // self.shards)
        sn = pow(self._subcoset_shift(j), n, Q)
        wmn = pow(root_of_unity(self.domain), self.shards * n, Q)
        return period, scalars_from_ints([pow((sn * pow(wmn, i, Q) - 1) % Q, -1, Q) for i in range(period)])

    def _circuit_rows(self, coeffs):
        """coefficient rows of the ten per-circuit polynomials of round 3: ql qr qm qo qc s1 s2 s3 L1 X"""
        n = self.n
        rows = self._zeros(10, n)
        rows[0:8].copy_(coeffs)
        rows[8, :n] = self.torch.from_numpy(_mont(pow(n, -1, Q)).view(np.int64)).to(self.dev)   # L1 = (1/n) sum X^i
        rows[9, 1] = self.torch.from_numpy(self._one.view(np.int64)).to(self.dev)              # the polynomial X
        return rows

    def _quotient_sharded(self, wv, coeffs, beta, gamma, alpha):
        """t on the whole quotient domain (natural order), computed sub-coset by sub-coset: this rank's part and an
        all-gather with a committer, all R parts locally without one"""
        torch, n, D, L, R = self.torch, self.n, self.domain, self.stride, self.shards
        m = D // R
        rows6 = self._zeros(6, L)
        rows6[0:5].copy_(wv[:, :L])
        one = self._one
        wm = _mont(self.omega)
        self._ck(self.lib.bpk_fr_scale_powers(self.ctx.handle, rows6[3].data_ptr(), wm.ctypes.data, one.ctypes.data,
                                              rows6[5].data_ptr(), n + 3), "bpk_fr_scale_powers")     # z(wX)
        crow = self._circuit_rows(coeffs)
        sc = _Scalars(beta, gamma, alpha, K1, K2)
        multi = self.committer is not None and self.committer.world > 1
        mine = [self.committer.rank] if multi else list(range(R))
        parts = self._empty(len(mine), m)
        for slot, j in enumerate(mine):
            if self.cache_preprocessed and self._pre is not None and j in self._pre[2]:
                ec = self._pre[2][j]
            else:
                ec = self._subcoset_evals(crow, n, j)
                if self.cache_preprocessed and self._pre is not None:
                    self._pre[2][j] = ec
            ew = self._subcoset_evals(rows6, L, j)
            period, zh = self._zh_inv_subcoset(j)
            self._ck(self.lib.bpk_plonk_quotient_evals_shard(
                self.ctx.handle, ew.data_ptr(), ec.data_ptr(), m, period, *sc.ptrs(), zh.ctypes.data,
                parts[slot].data_ptr()), "bpk_plonk_quotient_evals_shard")
        if multi:
            import torch.distributed as dist
            gathered = self._empty(R, m)
            dist.all_gather_into_tensor(gathered.view(-1), parts.view(-1), group=self.committer.group)
            parts = gathered
        # point j + R i of the domain is point i of sub-coset j
        return parts.permute(1, 0, 2).contiguous().view(D, 4)

    def _preprocessed(self):
        """Per-circuit data of round 3: coefficient forms of the eight pre-processed columns (the reference
        runs these i_ntt_381 on every prove, prover.rs round 3) and the coset evaluations of
        ql qr qm qo qc s1 s2 s3 L1 X on the quotient domain.  Kept across proofs with cache_preprocessed."""
        if self._pre is not None:
            return self._pre
        n, D = self.n, self.domain
        coeffs = self._empty(8, n)
        self._intt(self.pk_lagrange, coeffs, n, 8)
        if self.shards > 1:   # the per-circuit evaluations are made per sub-coset (and cached there)
            pre = (coeffs, None, {})
            if self.cache_preprocessed:
                self._pre = pre
            return pre
        cv = self._zeros(10, D)
        cv[0:8, :n].copy_(coeffs)
        cv[8, :n] = self.torch.from_numpy(_mont(pow(n, -1, Q)).view(np.int64)).to(self.dev)   # L1 = (1/n) sum X^i
        cv[9, 1] = self.torch.from_numpy(self._one.view(np.int64)).to(self.dev)              # the polynomial X
        self._ck(self.lib.bpk_ntt_fr_dev(self.ctx.handle, cv.data_ptr(), cv.data_ptr(), D, 10, 2,    # 2 = coset
                                         self._shift.ctypes.data), "bpk_ntt_fr_dev")
        pre = (coeffs, cv, {})
        if self.cache_preprocessed:
            self._pre = pre
        return pre

    def _wires_on_device(self, wires):
        """(A, B, C) as one [3, n, 4] device tensor; accepts numpy columns, a host int64 tensor [3, n, 4]
        (pinned memory makes the upload a single DMA at PCIe speed) or a CUDA int64 tensor"""
        torch, n = self.torch, self.n
        if isinstance(wires, torch.Tensor):
            if wires.dtype != torch.int64 or tuple(wires.shape) != (3, n, 4):
                raise ValueError("expected an int64 tensor of shape [3, n, 4]")
            if wires.device.type == "cpu":
                return wires.contiguous().to(self.dev, non_blocking=True)
            if wires.device != self.dev:
                raise ValueError("witness tensor lives on another device")
            return wires.contiguous()
        W = self._empty(3, n)
        for k in range(3):
            col = np.ascontiguousarray(wires[k], dtype=np.uint64).reshape(n, 4)
            W[k].copy_(torch.from_numpy(col.view(np.int64)))
        return W

    # ---- Prover::prove --------------------------------------------------------------------------
    def prove(self, wires, public_inputs: Sequence[int], blinding: Sequence[int],
              trace: Optional[dict] = None) -> Proof:
        """``wires`` = the (A, B, C) witness columns on H (uint64[n, 4] Montgomery each, or one CUDA int64
        tensor [3, n, 4]; prover.rs:177-214 fills them from the witness map); ``public_inputs`` the public
        values in declaration order; ``blinding`` the 11 scalars b_1..b_11 the reference draws from
        thread_rng (prover.rs:108-110)."""
        torch = self.torch
        self.ctx.bind_torch_stream(torch)   # torch copies / collectives and the library's kernels share one stream
        n, D, L = self.n, self.domain, self.stride
        b = [int(x) % Q for x in blinding]
        if len(b) != 11:
            raise BpkPanic("11 blinding scalars expected")
        tr = PlonkTranscript()
        marks = [("start", time.perf_counter())]
        self._pt = torch.empty(18, dtype=torch.int64, device=self.dev)
        # blinding values in the order they are patched in: (b2 + b1 X) Z_H etc.
        blind = self._upload(scalars_from_ints([b[1], b[0], b[3], b[2], b[5], b[4], b[8], b[7], b[6], b[9], b[10]]))

        # coefficient forms of the per-proof polynomials, zero-padded to the quotient domain so that the same
        # rows are the input of the coset transform of round 3
        wv = self._zeros(5, D)
        row = {name: wv[i] for i, name in enumerate(("a", "b", "c", "z", "pi"))}

        # ---- round 1 (prover.rs:177-277)
        W = self._wires_on_device(wires)
        tmp = self._empty(3, n)
        self._intt(W, tmp, n, 3)
        for k, name in enumerate(("a", "b", "c")):
            r = row[name]
            r[:n].copy_(tmp[k])
            bl = blind[2 * k:2 * k + 2]
            self._vec(1, r[0:2], bl, None, r[0:2], 2)          # -(b_lo + b_hi X)
            r[n:n + 2].copy_(bl)                               # +(b_lo + b_hi X) X^n
        a_1, b_1, c_1 = self._commit_many([(row["a"], n + 2), (row["b"], n + 2), (row["c"], n + 2)])
        tr.append_point(b"a_1", a_1)
        tr.append_point(b"b_1", b_1)
        tr.append_point(b"c_1", c_1)
        beta = tr.get_and_append_challenge(b"beta")
        gamma = tr.get_and_append_challenge(b"gamma")
        marks.append(("round1", time.perf_counter()))

        # ---- round 2 (prover.rs:279-368)
        Z = self._empty(n + 1)
        s_l = self.pk_lagrange
        sc = _Scalars(beta, gamma, K1, K2)
        self._ck(self.lib.bpk_plonk_grand_product(
            self.ctx.handle, W[0].data_ptr(), W[1].data_ptr(), W[2].data_ptr(), s_l[5].data_ptr(), s_l[6].data_ptr(),
            s_l[7].data_ptr(), n, *sc.ptrs(), Z.data_ptr()), "bpk_plonk_grand_product")
        if _from_mont(Z[n].cpu().numpy().view(np.uint64)) != 1:
            raise BpkPanic("assertion `left == right` failed: z_values.pop() == Scalar::one()")  # prover.rs:317
        z = row["z"]
        self._intt(Z, z, n)
        bl = blind[6:9]
        self._vec(1, z[0:3], bl, None, z[0:3], 3)
        z[n:n + 3].copy_(bl)
        z_1 = self._commit(z, n + 3)
        tr.append_point(b"z_1", z_1)
        alpha = tr.get_and_append_challenge(b"z_1")  # sic: src/transcript.rs:24 labels alpha "z_1"
        marks.append(("round2", time.perf_counter()))

        # ---- round 3 (prover.rs:370-500)
        pk, cv, _ = self._preprocessed()
        ql, qr, qm, qo, qc, s1c, s2c, s3c = (pk[i] for i in range(8))
        pi_l = self._zeros(n)
        if len(public_inputs):
            pi_l[:len(public_inputs)].copy_(self._upload(scalars_from_ints([(-int(v)) % Q for v in public_inputs])))
        self._intt(pi_l, row["pi"], n)
        # keep the coefficient forms (rounds 4-5 need them), transform a copy
        keep = self._empty(5, L)
        keep.copy_(wv[:, :L])
        if self.shards > 1:
            t = self._quotient_sharded(wv, pk, beta, gamma, alpha)
        else:
            self._ck(self.lib.bpk_ntt_fr_dev(self.ctx.handle, wv.data_ptr(), wv.data_ptr(), D, 5, 2,     # 2 = coset
                                             self._shift.ctypes.data), "bpk_ntt_fr_dev")
            t = self._empty(D)
            sc = _Scalars(beta, gamma, alpha, K1, K2)
            self._ck(self.lib.bpk_plonk_quotient_evals(
                self.ctx.handle, wv.data_ptr(), cv.data_ptr(), D, n, *sc.ptrs(), self._zh_inv.ctypes.data, t.data_ptr()),
                "bpk_plonk_quotient_evals")
        del wv, cv
        row = {name: keep[i] for i, name in enumerate(("a", "b", "c", "z", "pi"))}
        z = row["z"]
        self._ck(self.lib.bpk_ntt_fr_dev(self.ctx.handle, t.data_ptr(), t.data_ptr(), D, 1, 3,      # inverse | coset
                                         self._shift.ctypes.data), "bpk_ntt_fr_dev")
        # split_t_to_3pieces (prover.rs:454-500): t_lo + b10 X^n | t_mid - b10 + b11 X^n | t_hi - b11
        parts = self._zeros(3, L)
        t_lo, t_mid, t_hi = parts[0], parts[1], parts[2]
        t_lo[:n].copy_(t[0:n])
        t_lo[n].copy_(blind[9])
        t_mid[:n].copy_(t[n:2 * n])
        self._vec(1, t_mid[0:1], blind[9:10], None, t_mid[0:1], 1)
        t_mid[n].copy_(blind[10])
        t_hi[:n + 6].copy_(t[2 * n:3 * n + 6])
        self._vec(1, t_hi[0:1], blind[10:11], None, t_hi[0:1], 1)
        del t
        t_lo_1, t_mid_1, t_hi_1 = self._commit_many([(t_lo, n + 1), (t_mid, n + 1), (t_hi, n + 6)])
        tr.append_point(b"t_lo_1", t_lo_1)
        tr.append_point(b"t_mid_1", t_mid_1)
        tr.append_point(b"t_hi_1", t_hi_1)
        zeta = tr.get_and_append_challenge(b"zeta")
        marks.append(("round3", time.perf_counter()))

        # ---- round 4 (prover.rs:502-541)
        a_bar, b_bar, c_bar, s1_bar, s2_bar, pi_zeta = self._eval_many(
            [(row["a"], n + 2), (row["b"], n + 2), (row["c"], n + 2), (s1c, n), (s2c, n), (row["pi"], n)], zeta)
        z_omega_bar = self._eval(z, n + 3, zeta * self.omega % Q)       # z(omega X) at zeta
        for lab, v in ((b"a_eval", a_bar), (b"b_eval", b_bar), (b"c_eval", c_bar), (b"s1_eval", s1_bar),
                       (b"s2_eval", s2_bar), (b"z_shifted_eval", z_omega_bar)):
            tr.append_scalar(lab, v)
        nu = tr.get_and_append_challenge(b"nu")
        marks.append(("round4", time.perf_counter()))

        # ---- round 5 (prover.rs:543-647): linearisation polynomial r(X), then the two opening quotients
        zeta_n = pow(zeta, n, Q)
        zh_zeta = (zeta_n - 1) % Q
        l1_zeta = zh_zeta * pow(n * (zeta - 1) % Q, -1, Q) % Q            # = (1/n) sum zeta^i
        f = (a_bar + zeta * beta + gamma) * (b_bar + zeta * beta * K1 + gamma) % Q * (c_bar + zeta * beta * K2 + gamma) % Q
        g = (a_bar + s1_bar * beta + gamma) * (b_bar + s2_bar * beta + gamma) % Q * z_omega_bar % Q
        a2 = alpha * alpha % Q
        r = self._zeros(L)
        self._axpy(r, a_bar * b_bar, qm, n)
        self._axpy(r, a_bar, ql, n)
        self._axpy(r, b_bar, qr, n)
        self._axpy(r, c_bar, qo, n)
        self._axpy(r, 1, qc, n)
        self._axpy(r, alpha * f + a2 * l1_zeta, z, n + 3)
        self._axpy(r, -alpha * g * beta, s3c, n)
        self._axpy(r, -zh_zeta, t_lo, n + 1)
        self._axpy(r, -zh_zeta * zeta_n, t_mid, n + 1)
        self._axpy(r, -zh_zeta * zeta_n * zeta_n, t_hi, n + 6)
        self._add_const(r, 0, pi_zeta - alpha * g * (c_bar + gamma) - a2 * l1_zeta)
        if self._eval(r, n + 6, zeta) != 0:
            raise BpkPanic("assertion `left == right` failed: r.coeffs_evaluate(zeta) == Scalar::zero()")
        nus = [pow(nu, k, Q) for k in range(6)]
        for k, (poly, length) in enumerate(((row["a"], n + 2), (row["b"], n + 2), (row["c"], n + 2), (s1c, n), (s2c, n)),
                                           start=1):
            self._axpy(r, nus[k], poly, length)
        self._add_const(r, 0, -(nus[1] * a_bar + nus[2] * b_bar + nus[3] * c_bar + nus[4] * s1_bar + nus[5] * s2_bar))
        w_zeta = self._empty(L)
        self._div_linear(r, n + 6, zeta, w_zeta)
        self._add_const(z, 0, -z_omega_bar)
        w_zeta_omega = self._empty(L)
        self._div_linear(z, n + 3, zeta * self.omega % Q, w_zeta_omega)
        w_zeta_1, w_zeta_omega_1 = self._commit_many([(w_zeta, n + 5), (w_zeta_omega, n + 2)])
        tr.append_point(b"w_zeta_1", w_zeta_1)
        tr.append_point(b"w_zeta_omega_1", w_zeta_omega_1)
        mu = tr.get_and_append_challenge(b"mu")
        marks.append(("round5", time.perf_counter()))
        self.last_round_seconds = {marks[i][0]: marks[i][1] - marks[i - 1][1] for i in range(1, len(marks))}
        if trace is not None:
            trace.update(beta=beta, gamma=gamma, alpha=alpha, zeta=zeta, nu=nu, mu=mu)
        return Proof(a_1=a_1, b_1=b_1, c_1=c_1, z_1=z_1, t_lo_1=t_lo_1, t_mid_1=t_mid_1, t_hi_1=t_hi_1,
                     w_zeta_1=w_zeta_1, w_zeta_omega_1=w_zeta_omega_1, a_bar=a_bar, b_bar=b_bar, c_bar=c_bar,
                     s1_bar=s1_bar, s2_bar=s2_bar, z_omega_bar=z_omega_bar)
