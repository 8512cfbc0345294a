"""CPU ORACLE (test infrastructure only): restatement of the reference's PLONK prover / verifier algebra,
its circuit pre-processing and its Fiat-Shamir transcript -- the CALLERS of the MSM / NTT hot path.

Purpose: config[0] of BASELINE.json (tests/verify_proof_test.rs: setup + prove + verify of the 3-gate
circuit, proof bytes as golden output) and end-to-end drop-in checks: the prover below runs on a
pluggable backend for exactly the three surfaces the reference routes through the hot path
(`Setup::commit`, `i_ntt_381`, `impl Mul for Polynomial`), so the same proof can be produced with the CPU
oracle backend and with the GPU library and compared byte for byte.

What is restated (paths relative to the reference root):
  src/prover.rs:106-675        Prover::prove, round_1..5, split_t_to_3pieces, monomial_z_to_z_omega
  src/verifier.rs:80-192       verifier equation, in trapdoor form (tau known -> no pairing needed)
  src/program.rs:51-194        selector / sigma polynomials, public-variable discovery
  src/assembly.rs:30-81        gate coefficient sign convention (gates are given as tuples; the string
                               parser itself is out of scope)
  src/polynomial.rs:57-380     Add / Sub / Mul<Scalar> / Div semantics incl. length rules and the Div quirk
  src/transcript.rs:8-86       label schedule, rejection-sampled challenges
  merlin 3.0.0 / keccak 0.1.5  (Cargo.lock; NOT vendored in the reference): STROBE-128 + keccak-f[1600],
                               restated from the published construction.
Parity status: the reference has no test pinning transcript or proof bytes and draws its blinding from
thread_rng (prover.rs:108-110), so this layer is pinned only against (a) merlin's published
conformance vector, (b) SURVEY.md 8c's survey-time model values for the n = 8 circuit with blinding
1..11 (challenges, a_1, SHA-256 of the 624-byte proof) and (c) self-verification.  "parity unpinned"
with respect to rustc output.
"""
from __future__ import annotations

import hashlib

from oracle import bls12_381 as O

Q = O.Q

# --------------------------------------------------------------------------------------------------
# keccak-f[1600] and merlin's STROBE-128 subset
# --------------------------------------------------------------------------------------------------
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f1600(state: bytearray) -> None:
    a = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")


class Strobe128:
    """merlin::strobe::Strobe128 (meta-AD, AD, PRF only)"""
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state = st
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data: bytes):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & self.FLAG_T == 0
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (self.FLAG_C | self.FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class MerlinTranscript:
    """merlin::Transcript (new / append_message / challenge_bytes)"""

    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)


class PlonkTranscript(MerlinTranscript):
    """src/transcript.rs:65-86 on top of merlin"""

    def __init__(self):
        super().__init__(b"plonk")  # prover.rs:112

    def append_point(self, label: bytes, pt):
        self.append_message(label, O.g1_to_compressed(pt))

    def append_scalar(self, label: bytes, s: int):
        self.append_message(label, O.fr_to_bytes(s))

    def get_and_append_challenge(self, label: bytes) -> int:
        while True:
            b = self.challenge_bytes(label, 32)
            v = int.from_bytes(b, "little")
            if v < Q and v != 0:  # Scalar::from_bytes is_some and != 0
                self.append_message(label, b)
                return v


# --------------------------------------------------------------------------------------------------
# circuit pre-processing  (src/program.rs, src/assembly.rs sign convention)
# --------------------------------------------------------------------------------------------------
class Gate:
    """one constraint row: wires (L, R, O variable names or None) and (L, R, M, O, C) coefficients
    of  ql*a + qr*b + qm*a*b + qo*c + qc + PI = 0"""

    def __init__(self, wires, coeffs, public=None):
        self.wires = tuple(wires)
        self.coeffs = tuple(c % Q for c in coeffs)
        self.public = public  # variable name if this row declares a public input

    @staticmethod
    def public_input(var):  # "v public": L = 1, wires (v, None, None)   assembly.rs / SURVEY App. A
        return Gate((var, None, None), (1, 0, 0, 0, 0), public=var)

    @staticmethod
    def mul(out, x, y):     # "out <== x * y": M = -1, O = 1
        return Gate((x, y, out), (0, 0, -1, 1, 0))

    @staticmethod
    def add(out, x, y):     # "out <== x + y": L = -1, R = -1, O = 1
        return Gate((x, y, out), (-1, -1, 0, 1, 0))

    @staticmethod
    def mul_add(out, x, y):  # "out <== x * y + y": R = -1, M = -1, O = 1
        return Gate((x, y, out), (0, -1, -1, 1, 0))


class Program:
    def __init__(self, gates, group_order: int):
        assert O.is_power_of_two(group_order) and len(gates) <= group_order
        self.gates = list(gates)
        self.n = group_order

    def selectors(self):
        """program.rs:51-75: (ql, qr, qm, qo, qc) on H"""
        cols = [[0] * self.n for _ in range(5)]
        for i, g in enumerate(self.gates):
            for k in range(5):
                cols[k][i] = g.coeffs[k]
        return cols

    def sigmas(self):
        """program.rs:76-147"""
        n = self.n
        roots = O.roots_of_unity(n)
        uses = {}
        for row, g in enumerate(self.gates):
            for col, var in enumerate(g.wires):
                uses.setdefault(var, []).append((col, row))
        for row in range(len(self.gates), n):
            for col in range(3):
                uses.setdefault(None, []).append((col, row))
        s = [list(roots), [r * 2 % Q for r in roots], [0] * n]
        for _, cells in uses.items():
            for i, (col, row) in enumerate(cells):
                ncol, nrow = cells[(i + 1) % len(cells)]
                s[ncol][nrow] = roots[row] * (col + 1) % Q  # Cell::label, utils.rs:29-36
        return s

    def public_vars(self):
        """program.rs:172-194 (declarations must be at the top)"""
        out, no_more = [], False
        for g in self.gates:
            if g.public is not None:
                assert not no_more, "Public var declarations must be at the top"
                out.append(g.public)
            else:
                no_more = True
        return out


# --------------------------------------------------------------------------------------------------
# polynomial helpers on coefficient lists (src/polynomial.rs semantics)
# --------------------------------------------------------------------------------------------------
def p_add(a, b):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % Q for i in range(n)]


def p_sub(a, b):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % Q for i in range(n)]


def p_scale(a, s):
    return [x * s % Q for x in a]


def p_add_scalar(a, s):   # Add<Scalar>, Monomial: coefficient 0 only (polynomial.rs:57-72)
    r = list(a)
    r[0] = (r[0] + s) % Q
    return r


def p_sub_scalar(a, s):   # Sub<Scalar>, Monomial (polynomial.rs:119-132)
    r = list(a)
    r[0] = (r[0] - s) % Q
    return r


def p_eval(a, x):         # coeffs_evaluate (polynomial.rs:34-45)
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % Q
    return acc


def p_div(c1, c2):
    """impl Div (polynomial.rs:314-380): schoolbook long division, trailing zeros stripped, remainder
    dropped, and the reference's quirk: a quotient coefficient is only recorded for the steps the loop
    executes (q.insert(0, coeff)), so interior zero quotient coefficients are lost."""
    c1, c2 = list(c1), list(c2)
    while c1 and c1[-1] == 0:
        c1.pop()
    while c2 and c2[-1] == 0:
        c2.pop()
    assert c2, "Division by zero polynomial"
    q, r = [], c1
    lead_inv = pow(c2[-1], -1, Q)
    nz = [(i, ci) for i, ci in enumerate(c2) if ci]  # the prover only divides by sparse polynomials
    while len(r) >= len(c2) and r[-1] != 0:
        coeff = r[-1] * lead_inv % Q
        diff = len(r) - len(c2)
        for i, ci in nz:
            r[diff + i] = (r[diff + i] - ci * coeff) % Q
        while r and r[-1] == 0:
            r.pop()
        q.insert(0, coeff)
    return q


# --------------------------------------------------------------------------------------------------
# backends: the three surfaces the reference routes through the hot path
# --------------------------------------------------------------------------------------------------
class OracleBackend:
    """CPU oracle: i_ntt_381 / impl Mul / Setup::commit by the restatements in oracle/bls12_381.py"""

    def __init__(self, srs_points, reference_msm=True):
        self.srs = srs_points
        self.reference_msm = reference_msm

    def i_ntt(self, values):
        return O.ntt_fast(values, inverse=True)

    def mul(self, a, b):
        return (O.Polynomial(a) * O.Polynomial(b)).values

    def commit(self, coeffs):
        assert len(coeffs) <= len(self.srs), "SRS too short"
        if self.reference_msm:
            return O.bucket_msm(self.srs, coeffs, 256, 4)
        return O.msm_naive(self.srs, coeffs)


class CRefBackend:
    """The three hot-path surfaces through the C restatement of the reference's OWN algorithms (oracle/ref_cpu.c):
    naive O(n^2) i_ntt_381 (utils.rs:106-129), impl Mul by coeffs_evaluate + naive inverse DFT
    (polynomial.rs:241-273), bucket_msm(256, 4) with complete projective additions (msm.rs:76-118).  This is the
    reference-algorithm prover whose time bench.py reports as the CPU baseline of the prove metric."""

    def __init__(self, srs_points):
        import numpy as np
        self.np = np
        self.srs = np.array([O.g1_scale_proj(p, 1) for p in srs_points], dtype=np.uint64)

    def _m(self, vals):
        return self.np.array([O.fr_to_mont(v) for v in vals], dtype=self.np.uint64).reshape(-1, 4)

    def _i(self, arr):
        return [O.fr_from_mont([int(x) for x in row]) for row in arr]

    def i_ntt(self, values):
        from oracle import cref
        return self._i(cref.ntt_381(self._m(values), inverse=True))

    def mul(self, a, b):
        from oracle import cref
        return self._i(cref.poly_mul(self._m(a), self._m(b)))

    def commit(self, coeffs):
        from oracle import cref
        assert len(coeffs) <= len(self.srs), "SRS too short"
        return O.g1_proj_limbs_to_affine([int(v) for v in cref.bucket_msm(self.srs, self._m(coeffs), 256, 4)])


# --------------------------------------------------------------------------------------------------
# prover (src/prover.rs) -- SURVEY.md Appendix A
# --------------------------------------------------------------------------------------------------
class Proof:
    POINTS = ("a_1", "b_1", "c_1", "z_1", "t_lo_1", "t_mid_1", "t_hi_1", "w_zeta_1", "w_zeta_omega_1")
    SCALARS = ("a_bar", "b_bar", "c_bar", "s1_bar", "s2_bar", "z_omega_bar")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to_bytes(self) -> bytes:
        """9 x to_compressed (48 B) in Proof field order (verifier.rs:23-40) + 6 x Scalar::to_bytes (32 B)"""
        out = b"".join(O.g1_to_compressed(getattr(self, k)) for k in self.POINTS)
        out += b"".join(O.fr_to_bytes(getattr(self, k)) for k in self.SCALARS)
        assert len(out) == 624
        return out

    def sha256(self) -> str:
        return hashlib.sha256(self.to_bytes()).hexdigest()


def prove(program: Program, witness: dict, blinding, backend, trace=None) -> Proof:
    """Prover::prove (prover.rs:106-176) with the 11 blinding scalars injected (the reference draws them
    from thread_rng).  `trace`, if a dict, receives the challenges."""
    n = program.n
    b = [x % Q for x in blinding]
    assert len(b) == 11
    omega = O.root_of_unity(n)
    roots = O.roots_of_unity(n)
    k1, k2 = 2, 3
    ql, qr, qm, qo, qc = program.selectors()
    s1, s2, s3 = program.sigmas()
    pub = program.public_vars()
    pi = [(-witness[v]) % Q for v in pub] + [0] * (n - len(pub))
    z_h = [Q - 1] + [0] * (n - 1) + [1]
    tr = PlonkTranscript()

    # ---- round 1 (prover.rs:177-277)
    A = [0] * n
    B = [0] * n
    C = [0] * n
    for i, g in enumerate(program.gates):
        A[i] = witness[g.wires[0]] % Q if g.wires[0] is not None else 0
        B[i] = witness[g.wires[1]] % Q if g.wires[1] is not None else 0
        C[i] = witness[g.wires[2]] % Q if g.wires[2] is not None else 0
    a = p_add(backend.mul([b[1], b[0]], z_h), backend.i_ntt(A))
    bb = p_add(backend.mul([b[3], b[2]], z_h), backend.i_ntt(B))
    c = p_add(backend.mul([b[5], b[4]], z_h), backend.i_ntt(C))
    a_1, b_1, c_1 = backend.commit(a), backend.commit(bb), backend.commit(c)
    tr.append_point(b"a_1", a_1)
    tr.append_point(b"b_1", b_1)
    tr.append_point(b"c_1", c_1)
    beta = tr.get_and_append_challenge(b"beta")
    gamma = tr.get_and_append_challenge(b"gamma")

    # ---- round 2 (prover.rs:279-368)
    def rlc(x, y):
        return (x + y * beta + gamma) % Q

    Z = [1]
    for i in range(n):
        num = rlc(A[i], roots[i]) * rlc(B[i], roots[i] * k1 % Q) % Q * rlc(C[i], roots[i] * k2 % Q) % Q
        den = rlc(A[i], s1[i]) * rlc(B[i], s2[i]) % Q * rlc(C[i], s3[i]) % Q
        Z.append(Z[-1] * num % Q * pow(den, -1, Q) % Q)
    assert Z.pop() == 1
    z = p_add(backend.mul([b[8], b[7], b[6]], z_h), backend.i_ntt(Z))
    z_1 = backend.commit(z)
    tr.append_point(b"z_1", z_1)
    alpha = tr.get_and_append_challenge(b"z_1")  # sic: transcript.rs:24

    # ---- round 3 (prover.rs:370-500)
    s1c, s2c, s3c = backend.i_ntt(s1), backend.i_ntt(s2), backend.i_ntt(s3)
    qlc, qrc, qmc, qoc, qcc = (backend.i_ntt(v) for v in (ql, qr, qm, qo, qc))
    pic = backend.i_ntt(pi)
    mul = backend.mul
    gate = p_add(p_add(p_add(p_add(p_add(mul(a, qlc), mul(bb, qrc)), mul(mul(a, bb), qmc)), mul(c, qoc)), pic), qcc)
    rho = backend.i_ntt(roots)
    z_omega = [cf * pow(omega, i, Q) % Q for i, cf in enumerate(z)]

    def prlc(p, o):  # Rlc for Polynomial (utils.rs:169-174): p + o*beta + gamma (gamma on coefficient 0)
        return p_add_scalar(p_add(p, p_scale(o, beta)), gamma)

    perm = p_sub(
        mul(mul(mul(prlc(a, rho), prlc(bb, p_scale(rho, k1))), prlc(c, p_scale(rho, k2))), z),
        mul(mul(mul(prlc(a, s1c), prlc(bb, s2c)), prlc(c, s3c)), z_omega))
    l1c = backend.i_ntt([1] + [0] * (n - 1))
    first = mul(p_sub_scalar(z, 1), l1c)
    allc = p_add(p_add(gate, p_scale(perm, alpha)), p_scale(first, alpha * alpha % Q))
    t = p_div(allc, z_h)
    assert len(t) >= 2 * n
    t_lo, t_mid, t_hi = t[0:n], t[n:2 * n], t[2 * n:]
    x_n = [0] * n + [1]
    t_lo = p_add(t_lo, p_scale(x_n, b[9]))
    t_mid = p_add(t_mid, p_sub_scalar(p_scale(x_n, b[10]), b[9]))
    t_hi = p_add_scalar(t_hi, (-b[10]) % Q)
    t_lo_1, t_mid_1, t_hi_1 = backend.commit(t_lo), backend.commit(t_mid), backend.commit(t_hi)
    tr.append_point(b"t_lo_1", t_lo_1)
    tr.append_point(b"t_mid_1", t_mid_1)
    tr.append_point(b"t_hi_1", t_hi_1)
    zeta = tr.get_and_append_challenge(b"zeta")

    # ---- round 4 (prover.rs:502-541)
    a_bar, b_bar, c_bar = p_eval(a, zeta), p_eval(bb, zeta), p_eval(c, zeta)
    s1_bar, s2_bar = p_eval(s1c, zeta), p_eval(s2c, zeta)
    z_omega_bar = p_eval(z_omega, zeta)
    for lab, v in ((b"a_eval", a_bar), (b"b_eval", b_bar), (b"c_eval", c_bar), (b"s1_eval", s1_bar),
                   (b"s2_eval", s2_bar), (b"z_shifted_eval", z_omega_bar)):
        tr.append_scalar(lab, v)
    nu = tr.get_and_append_challenge(b"nu")

    # ---- round 5 (prover.rs:543-647)
    r1 = p_add(p_add_scalar(p_add(p_add(p_add(p_scale(p_scale(qmc, a_bar), b_bar), p_scale(qlc, a_bar)),
                                        p_scale(qrc, b_bar)), p_scale(qoc, c_bar)), p_eval(pic, zeta)), qcc)
    f1 = (a_bar + zeta * beta + gamma) % Q
    f2 = (b_bar + zeta * beta * k1 + gamma) % Q
    f3 = (c_bar + zeta * beta * k2 + gamma) % Q
    g1 = (a_bar + s1_bar * beta + gamma) % Q
    g2 = (b_bar + s2_bar * beta + gamma) % Q
    r2 = p_sub(p_scale(p_scale(p_scale(z, f1), f2), f3),
               p_scale(p_scale(p_scale(p_add_scalar(p_add_scalar(p_scale(s3c, beta), c_bar), gamma), g1), g2), z_omega_bar))
    r3 = p_scale(p_sub_scalar(z, 1), p_eval(l1c, zeta))
    zh_zeta = p_eval(z_h, zeta)
    r4 = p_scale(p_add(p_add(t_lo, p_scale(t_mid, pow(zeta, n, Q))), p_scale(t_hi, pow(zeta, 2 * n, Q))), zh_zeta)
    r = p_sub(p_add(p_add(r1, p_scale(r2, alpha)), p_scale(p_scale(r3, alpha), alpha)), r4)
    assert p_eval(r, zeta) == 0
    w_num = p_add(p_add(p_add(p_add(p_add(r, p_scale(p_sub_scalar(a, a_bar), nu)),
                                    p_scale(p_scale(p_sub_scalar(bb, b_bar), nu), nu)),
                              p_scale(p_sub_scalar(c, c_bar), pow(nu, 3, Q))),
                        p_scale(p_sub_scalar(s1c, s1_bar), pow(nu, 4, Q))),
                  p_scale(p_sub_scalar(s2c, s2_bar), pow(nu, 5, Q)))
    w_zeta = p_div(w_num, [(-zeta) % Q, 1])
    w_zeta_omega = p_div(p_sub_scalar(z, z_omega_bar), [(-zeta * omega) % Q, 1])
    w_zeta_1, w_zeta_omega_1 = backend.commit(w_zeta), backend.commit(w_zeta_omega)
    tr.append_point(b"w_zeta_1", w_zeta_1)
    tr.append_point(b"w_zeta_omega_1", w_zeta_omega_1)
    mu = tr.get_and_append_challenge(b"mu")
    if trace is not None:
        trace.update(beta=beta, gamma=gamma, alpha=alpha, zeta=zeta, nu=nu, mu=mu)
    return Proof(a_1=a_1, b_1=b_1, c_1=c_1, z_1=z_1, t_lo_1=t_lo_1, t_mid_1=t_mid_1, t_hi_1=t_hi_1,
                 w_zeta_1=w_zeta_1, w_zeta_omega_1=w_zeta_omega_1, a_bar=a_bar, b_bar=b_bar, c_bar=c_bar,
                 s1_bar=s1_bar, s2_bar=s2_bar, z_omega_bar=z_omega_bar)


# --------------------------------------------------------------------------------------------------
# verifier (src/verifier.rs:80-192) in trapdoor form: tau known, no pairing
# --------------------------------------------------------------------------------------------------
def verify(program: Program, proof: Proof, public_inputs, tau: int, commit) -> bool:
    """Recomputes the challenges from the proof, then checks
         tau * (W_zeta + mu W_zeta_omega) == zeta W_zeta + mu zeta omega W_zeta_omega + F - E   in G1,
    which is the pairing equation of verifier.rs:186-190 with the trapdoor known.  `commit` commits the
    eight pre-processed polynomials (verifier.rs:49-79 does that through Setup::commit, i.e. the hot path)."""
    ql, qr, qm, qo, qc = program.selectors()
    s1, s2, s3 = program.sigmas()
    inv = lambda v: O.ntt_fast(v, inverse=True)
    commitments = [commit(inv(v)) for v in (qm, ql, qr, qo, qc, s1, s2, s3)]
    return verify_with_commitments(program.n, proof, public_inputs, tau, commitments)


def lagrange_public_input_eval(n: int, public_inputs, zeta: int) -> int:
    """PI(zeta) for PI = sum_i (-pub_i) L_i (verifier.rs:95-104 interpolates the column and evaluates it):
    L_i(zeta) = w^i (zeta^n - 1) / (n (zeta - w^i)), O(len(public_inputs)) instead of an n-point transform"""
    omega = O.root_of_unity(n)
    zh = (pow(zeta, n, Q) - 1) % Q
    acc, wi = 0, 1
    for v in public_inputs:
        acc = (acc - v * wi % Q * zh % Q * pow(n * (zeta - wi) % Q, -1, Q)) % Q
        wi = wi * omega % Q
    return acc


def verify_columns(n: int, selectors_mont, sigmas_mont, proof: Proof, public_inputs, tau: int, threads: int = 0) -> bool:
    """verify() for circuits given as pre-processed COLUMNS (uint64[n, 4] Montgomery arrays: [QL, QR, QM, QO, QC],
    [S1, S2, S3]) at sizes where Python transforms are out of reach (2^20+ gates).  The eight pre-processed
    commitments are formed on the ORACLE side in closed form, [p(tau)]G with p(tau) from the C inverse transform +
    Horner (oracle/fast_cpu.c::oracle_fr_poly_at) -- no GPU result enters the check except the proof itself."""
    import numpy as np
    from oracle import cref
    tau_m = np.array(O.fr_to_mont(tau % Q), dtype=np.uint64)
    ql, qr, qm, qo, qc = selectors_mont
    s1, s2, s3 = sigmas_mont
    commitments = []
    for col in (qm, ql, qr, qo, qc, s1, s2, s3):
        assert col.shape == (n, 4)
        e = O.fr_from_mont([int(x) for x in cref.poly_at(col, tau_m, threads)])
        commitments.append(O.g1_mul(O.G1_GEN, e))
    return verify_with_commitments(n, proof, public_inputs, tau, commitments, fast_pi=True)


def verify_with_commitments(n: int, proof: Proof, public_inputs, tau: int, commitments, fast_pi: bool = False) -> bool:
    """the verifier equation given the commitments [QM, QL, QR, QO, QC, S1, S2, S3] of the pre-processed polynomials"""
    omega = O.root_of_unity(n)
    G = O.G1_GEN
    inv = lambda v: O.ntt_fast(v, inverse=True)
    cqm, cql, cqr, cqo, cqc, cs1, cs2, cs3 = commitments
    tr = PlonkTranscript()
    for lab in ("a_1", "b_1", "c_1"):
        tr.append_point(lab.encode(), getattr(proof, lab))
    beta = tr.get_and_append_challenge(b"beta")
    gamma = tr.get_and_append_challenge(b"gamma")
    tr.append_point(b"z_1", proof.z_1)
    alpha = tr.get_and_append_challenge(b"z_1")
    for lab in ("t_lo_1", "t_mid_1", "t_hi_1"):
        tr.append_point(lab.encode(), getattr(proof, lab))
    zeta = tr.get_and_append_challenge(b"zeta")
    for lab, k in ((b"a_eval", "a_bar"), (b"b_eval", "b_bar"), (b"c_eval", "c_bar"), (b"s1_eval", "s1_bar"),
                   (b"s2_eval", "s2_bar"), (b"z_shifted_eval", "z_omega_bar")):
        tr.append_scalar(lab, getattr(proof, k))
    nu = tr.get_and_append_challenge(b"nu")
    tr.append_point(b"w_zeta_1", proof.w_zeta_1)
    tr.append_point(b"w_zeta_omega_1", proof.w_zeta_omega_1)
    mu = tr.get_and_append_challenge(b"mu")

    zh = (pow(zeta, n, Q) - 1) % Q
    l1 = zh * pow(n * (zeta - 1) % Q, -1, Q) % Q
    if fast_pi:
        pi_zeta = lagrange_public_input_eval(n, public_inputs, zeta)
    else:
        pi_vals = [(-v) % Q for v in public_inputs] + [0] * (n - len(public_inputs))
        pi_zeta = p_eval(inv(pi_vals), zeta)
    ab, bb_, cb, s1b, s2b, zwb = (proof.a_bar, proof.b_bar, proof.c_bar, proof.s1_bar, proof.s2_bar, proof.z_omega_bar)
    r0 = (pi_zeta - l1 * alpha * alpha - alpha * (ab + beta * s1b + gamma) * (bb_ + beta * s2b + gamma) % Q
          * (cb + gamma) % Q * zwb) % Q
    mulp, addp = O.g1_mul, O.g1_add
    D = None
    for pt, k in ((cqm, ab * bb_), (cql, ab), (cqr, bb_), (cqo, cb), (cqc, 1)):
        D = addp(D, mulp(pt, k % Q))
    zc = ((ab + beta * zeta + gamma) * (bb_ + beta * 2 * zeta + gamma) % Q * (cb + beta * 3 * zeta + gamma) % Q * alpha
          + l1 * alpha * alpha + mu) % Q
    D = addp(D, mulp(proof.z_1, zc))
    D = addp(D, O.g1_neg(mulp(cs3, (ab + beta * s1b + gamma) * (bb_ + beta * s2b + gamma) % Q * alpha % Q * beta % Q * zwb % Q)))
    tsum = addp(addp(proof.t_lo_1, mulp(proof.t_mid_1, pow(zeta, n, Q))), mulp(proof.t_hi_1, pow(zeta, 2 * n, Q)))
    D = addp(D, O.g1_neg(mulp(tsum, zh)))
    F = D
    for pt, k in ((proof.a_1, nu), (proof.b_1, pow(nu, 2, Q)), (proof.c_1, pow(nu, 3, Q)), (cs1, pow(nu, 4, Q)),
                  (cs2, pow(nu, 5, Q))):
        F = addp(F, mulp(pt, k))
    e = (nu * ab + pow(nu, 2, Q) * bb_ + pow(nu, 3, Q) * cb + pow(nu, 4, Q) * s1b + pow(nu, 5, Q) * s2b + mu * zwb - r0) % Q
    E = mulp(G, e)
    lhs = mulp(addp(proof.w_zeta_1, mulp(proof.w_zeta_omega_1, mu)), tau)
    rhs = addp(addp(addp(mulp(proof.w_zeta_1, zeta), mulp(proof.w_zeta_omega_1, mu * zeta % Q * omega % Q)), F), O.g1_neg(E))
    return lhs == rhs


# --------------------------------------------------------------------------------------------------
# the reference's test program (tests/verify_proof_test.rs:13-50) and a synthetic circuit family
# --------------------------------------------------------------------------------------------------
def reference_test_circuit():
    """constraints "e public", "c <== a * b + b", "e <== c * d", group order 8; witness a=3 b=4 c=16 d=5 e=80"""
    gates = [Gate.public_input("e"), Gate.mul_add("c", "a", "b"), Gate.mul("e", "c", "d")]
    witness = {"a": 3, "b": 4, "c": 16, "d": 5, "e": 80}
    return Program(gates, 8), witness, [80]


def synthetic_circuit(n: int, gates_used: int, seed: int = 1):
    """row 0 declares the public output; rows alternate x_k <== x_i * x_j and x_k <== x_i + x_j, each
    output feeding later gates; the last gate's output is the public variable."""
    assert gates_used >= 3 and gates_used <= n
    rnd = O.splitmix64_stream(seed, 4 * gates_used)
    w = {"x0": rnd[0] % 1000 + 2, "x1": rnd[1] % 1000 + 3}
    names = ["x0", "x1"]
    body = []
    for k in range(gates_used - 1):
        i = names[rnd[2 + 2 * k] % len(names)]
        j = names[-1]
        out = "x%d" % (len(names))
        if k % 2 == 0:
            body.append(Gate.mul(out, i, j))
            w[out] = w[i] * w[j] % Q
        else:
            body.append(Gate.add(out, i, j))
            w[out] = (w[i] + w[j]) % Q
        names.append(out)
    pub = names[-1]
    gates = [Gate.public_input(pub)] + body
    return Program(gates, n), w, [w[pub]]
