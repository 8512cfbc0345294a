"""CPU ORACLE (test infrastructure only) for the baby-plonk-rust MSM / NTT hot path.

This file is a big-integer restatement of the arithmetic the reference performs on the
hot path.  It is TEST INFRASTRUCTURE: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.  The product path
(``baby-plonk-rust_b200``) never imports anything from ``oracle/``.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this module against
 * the 6-limb Fp and 4-limb Fr Montgomery known-answer vectors of the vendored curve
   library (lib/bls12_381/src/fp.rs:700-941, scalar.rs:160-221,795-1046),
 * the 1000-entry G1 encoding vectors [i]G (lib/bls12_381/src/tests/
   g1_{compressed,uncompressed}_valid_test_vectors.dat, harness tests/mod.rs:3-58),
 * [2]G affine limbs (g1.rs:1263-1297),
 * the algebraic MSM / NTT pins of src/setup.rs:45-116, src/polynomial.rs:437-451,
   src/utils.rs:238-242.
Fixtures live in tests/golden/ and were produced by tests/golden/make_golden.py from
/root/reference.

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

# --------------------------------------------------------------------------------------
# Constants
# --------------------------------------------------------------------------------------
# lib/bls12_381/src/fp.rs:69-77
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
# lib/bls12_381/src/scalar.rs:80-88
Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FP_R = (1 << 384) % P      # fp.rs:83-90
FR_R = (1 << 256) % Q      # scalar.rs:167-172
FP_RINV = pow(FP_R, -1, P)
FR_RINV = pow(FR_R, -1, Q)
FR_S = 32                  # scalar.rs:199
FR_GENERATOR = 7           # scalar.rs:106-113
# scalar.rs:201-213: ROOT_OF_UNITY = GENERATOR^t, t*2^32 + 1 = q
ROOT_OF_UNITY = pow(FR_GENERATOR, (Q - 1) >> FR_S, Q)
ROOT_OF_UNITY_INV = pow(ROOT_OF_UNITY, -1, Q)
CURVE_B = 4                # g1.rs:176-183 (Montgomery limbs of 4)
# g1.rs:199-214 generator (canonical values; the golden test re-derives the limbs)
G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G1_GEN = (G1_X, G1_Y)
MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------
# Limb codecs: the representation Rust holds in memory (u64 little-endian limbs, Montgomery)
# --------------------------------------------------------------------------------------
def int_to_limbs(x: int, n: int) -> list[int]:
    return [(x >> (64 * i)) & MASK64 for i in range(n)]


def limbs_to_int(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def fr_to_mont(x: int) -> list[int]:
    """canonical Fr -> Scalar.0 ([u64;4], value x*R mod q).  scalar.rs:282-284"""
    return int_to_limbs((x % Q) * FR_R % Q, 4)


def fr_from_mont(limbs) -> int:
    """Scalar.0 -> canonical value (what to_bytes yields).  scalar.rs:292-304"""
    return limbs_to_int(limbs) * FR_RINV % Q


def fp_to_mont(x: int) -> list[int]:
    """canonical Fp -> Fp.0 ([u64;6]).  fp.rs:199-201"""
    return int_to_limbs((x % P) * FP_R % P, 6)


def fp_from_mont(limbs) -> int:
    """fp.rs:206-227"""
    return limbs_to_int(limbs) * FP_RINV % P


def fr_mont_mul_limbs(a, b) -> list[int]:
    """Scalar::mul on raw Montgomery limbs: (a*b)/R mod q.  scalar.rs:562-586 + 514-558"""
    return int_to_limbs(limbs_to_int(a) * limbs_to_int(b) * FR_RINV % Q, 4)


def fp_mont_mul_limbs(a, b) -> list[int]:
    """Fp::mul on raw Montgomery limbs.  fp.rs:565-609 + 487-562"""
    return int_to_limbs(limbs_to_int(a) * limbs_to_int(b) * FP_RINV % P, 6)


def fr_from_bytes_wide(b: bytes) -> int:
    """Scalar::from_bytes_wide: 512-bit LE integer reduced mod q.  scalar.rs:308-339"""
    assert len(b) == 64
    return int.from_bytes(b, "little") % Q


def fr_to_bytes(x: int) -> bytes:
    """Scalar::to_bytes: 32 B little-endian canonical.  scalar.rs:292-304"""
    return (x % Q).to_bytes(32, "little")


# --------------------------------------------------------------------------------------
# G1: affine group law on canonical integers.  None = point at infinity.
# Parity on the reference's G1Projective is judged on G1Affine::from (g1.rs:49-63), so an
# affine model is a complete description of every value the reference can output.
# --------------------------------------------------------------------------------------
def g1_is_on_curve(pt) -> bool:
    """g1.rs:427-430: y^2 = x^3 + 4 or infinity"""
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - CURVE_B) % P == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(p1, p2):
    """group addition, all degenerate cases (what the complete formulas of g1.rs:670-712 give
    after G1Affine::from)"""
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_double(pt):
    """g1.rs:638-667"""
    return g1_add(pt, pt)


# Jacobian helpers (internal speed-up only; results are converted back to affine)
def _jac_double(X, Y, Z):
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % P
    B = Y * Y % P
    C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jac_add_affine(X1, Y1, Z1, x2, y2):
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    H = (U2 - X1) % P
    r = (S2 - Y1) % P
    if H == 0:
        if r == 0:
            return _jac_double(X1, Y1, Z1)
        return (1, 1, 0)
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (r * r - HHH - 2 * V) % P
    Y3 = (r * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def _jac_to_affine(X, Y, Z):
    if Z == 0:
        return None
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g1_mul(pt, k: int):
    """[k]P, k reduced mod q (G1Projective * Scalar, g1.rs:754-774 double-and-add MSB first)"""
    k %= Q
    if pt is None or k == 0:
        return None
    x, y = pt
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(*acc)
        if bit == "1":
            acc = _jac_add_affine(*acc, x, y)
    return _jac_to_affine(*acc)


def g1_sum(points):
    acc = None
    for p in points:
        acc = g1_add(acc, p)
    return acc


# ---- projective (X:Y:Z homogeneous) Montgomery limb interface: what Rust hands over ----
def g1_affine_to_proj_limbs(pt) -> list[int]:
    """G1Projective from G1Affine (g1.rs:468-476): (x, y, 1) or identity (0, 1, 0) (g1.rs:605-611);
    returns 18 u64 Montgomery limbs X|Y|Z."""
    if pt is None:
        return fp_to_mont(0) + fp_to_mont(1) + fp_to_mont(0)
    return fp_to_mont(pt[0]) + fp_to_mont(pt[1]) + fp_to_mont(1)


def g1_proj_limbs_to_affine(limbs):
    """G1Affine::from(&G1Projective) (g1.rs:49-63): x = X/Z, y = Y/Z; Z == 0 -> identity"""
    X = fp_from_mont(limbs[0:6])
    Y = fp_from_mont(limbs[6:12])
    Z = fp_from_mont(limbs[12:18])
    if Z == 0:
        return None
    zi = pow(Z, -1, P)
    return (X * zi % P, Y * zi % P)


def g1_scale_proj(pt, z: int) -> list[int]:
    """a non-normalised projective representative (x*z, y*z, z) of an affine point, as limbs"""
    if pt is None:
        return fp_to_mont(0) + fp_to_mont(z) + fp_to_mont(0)
    return fp_to_mont(pt[0] * z) + fp_to_mont(pt[1] * z) + fp_to_mont(z)


# ---- encodings ----
def fp_lexicographically_largest(y: int) -> bool:
    """fp.rs:273-298: y > (p-1)/2"""
    return y > (P - 1) // 2


def g1_to_compressed(pt) -> bytes:
    """G1Affine::to_compressed.  g1.rs:221-242"""
    if pt is None:
        b = bytearray(48)
        b[0] |= 0x80 | 0x40
        return bytes(b)
    b = bytearray(pt[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if fp_lexicographically_largest(pt[1]):
        b[0] |= 0x20
    return bytes(b)


def g1_from_compressed(b: bytes):
    """G1Affine::from_compressed_unchecked.  g1.rs:336-398: y = sqrt(x^3 + 4) with the sign bit"""
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = pow((x * x * x + 4) % P, (P + 1) // 4, P)   # p = 3 mod 4
    assert y * y % P == (x * x * x + 4) % P, "not on the curve"
    if fp_lexicographically_largest(y) != bool(b[0] & 0x20):
        y = P - y
    return (x, y)


def g1_to_uncompressed(pt) -> bytes:
    """G1Affine::to_uncompressed.  g1.rs:246-260"""
    if pt is None:
        b = bytearray(96)
        b[0] |= 0x40
        return bytes(b)
    return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")


def g1_from_uncompressed(b: bytes):
    """G1Affine::from_uncompressed_unchecked.  g1.rs:273-322"""
    assert len(b) == 96
    if b[0] & 0x40:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    y = int.from_bytes(b[48:], "big")
    return (x, y)


# --------------------------------------------------------------------------------------
# MSM: restatement of src/msm.rs
# --------------------------------------------------------------------------------------
def get_c_bit_chunk(scalar: int, chunk_index: int, chunk_size: int) -> int:
    """msm.rs:119-139: bits [i*c, (i+1)*c) of the canonical scalar, counted from the MSB of its
    256-bit big-endian string."""
    start = chunk_index * chunk_size
    end = start + chunk_size
    if end > 256:
        raise IndexError("slice out of range (msm.rs:133 panics)")
    return (scalar >> (256 - end)) & ((1 << chunk_size) - 1)


def c_bit_msm(points, digits, c: int):
    """msm.rs:23-49: 2^c-1 buckets, scatter-add (zip truncates), running-sum reduction."""
    nb = (1 << c) - 1
    buckets = [None] * nb
    for pt, d in zip(points, digits):
        if d != 0:
            buckets[d - 1] = g1_add(buckets[d - 1], pt)
    acc = None
    res = None
    for b in reversed(buckets):
        acc = g1_add(acc, b)
        res = g1_add(res, acc)
    return res


def bucket_msm(points, scalars, b: int = 256, c: int = 4):
    """BucketMSM::bucket_msm (msm.rs:76-118).  points: affine tuples / None; scalars: canonical
    ints.  k = b // c MSB-first windows; windows are taken over *all* scalars (msm.rs:91-98) but
    the scatter zips with points (msm.rs:29), i.e. only min(len) pairs contribute.  Returns the
    affine image of the result.  If k == 0 the reference indexes t_points[0] and panics."""
    k = b // c
    if k == 0:
        raise IndexError("t_points[0] out of range (msm.rs:105 panics)")
    t_points = []
    for i in range(k):
        digits = [get_c_bit_chunk(s, i, c) for s in scalars]
        t_points.append(c_bit_msm(points, digits, c))
    res = t_points[0]
    for j in range(1, k):
        for _ in range(c):
            res = g1_double(res)
        res = g1_add(res, t_points[j])
    return res


def msm_naive(points, scalars):
    """Σ s_i P_i over min(len) pairs: the value bucket_msm(.., 256, c | 256) must equal."""
    acc = None
    for pt, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(pt, s))
    return acc


def generate_srs_points(powers: int, tau: int):
    """Setup::generate_srs powers_of_x (setup.rs:12-31): [tau^i]G for i < powers."""
    out = []
    t = 1
    for _ in range(powers):
        out.append(g1_mul(G1_GEN, t))
        t = t * tau % Q
    return out


# --------------------------------------------------------------------------------------
# NTT: restatement of src/utils.rs
# --------------------------------------------------------------------------------------
def is_power_of_two(n: int) -> bool:
    """utils.rs:82-84"""
    return n != 0 and (n & (n - 1)) == 0


def root_of_unity(group_order: int) -> int:
    """utils.rs:39-43: ROOT_OF_UNITY^(2^32/group_order) (integer division, as the reference)"""
    return pow(ROOT_OF_UNITY, (1 << 32) // group_order, Q)


def roots_of_unity(group_order: int) -> list[int]:
    """utils.rs:45-52"""
    g = root_of_unity(group_order)
    res = [1]
    for _ in range(1, group_order):
        res.append(res[-1] * g % Q)
    return res


def find_next_power_of_two(n: int, m: int) -> int:
    """utils.rs:54-61"""
    power = 1
    target = n + m + 1
    while power < target:
        power <<= 1
    return power


def ntt_381(elements: list[int]) -> list[int]:
    """utils.rs:63-81: naive DFT, out[x] = Σ_y in[y]·ROOT^(x·y·2^32/n); natural order."""
    n = len(elements)
    assert is_power_of_two(n)
    w = pow(ROOT_OF_UNITY, (1 << 32) // n, Q)
    pw = [pow(w, i, Q) for i in range(n)]
    return [sum(e * pw[(x * y) % n] for y, e in enumerate(elements)) % Q for x in range(n)]


def i_ntt_381(elements: list[int]) -> list[int]:
    """utils.rs:106-129: naive inverse DFT with ROOT_OF_UNITY_INV, scaled by n^-1."""
    n = len(elements)
    assert is_power_of_two(n)
    w = pow(ROOT_OF_UNITY_INV, (1 << 32) // n, Q)
    pw = [pow(w, i, Q) for i in range(n)]
    ninv = pow(n, -1, Q)
    return [sum(e * pw[(x * y) % n] for y, e in enumerate(elements)) * ninv % Q for x in range(n)]


def ntt_fast(elements: list[int], inverse: bool = False, coset_shift: int | None = None) -> list[int]:
    """O(n log n) radix-2 evaluation of the same DFT as ntt_381 / i_ntt_381 (equality with the
    naive definition is itself a test).  coset_shift g: forward evaluates on g·ω^i; inverse
    interpolates from g·ω^i (undoes the forward coset transform)."""
    n = len(elements)
    assert is_power_of_two(n)
    a = [e % Q for e in elements]
    if coset_shift is not None and not inverse:
        g = 1
        for i in range(n):
            a[i] = a[i] * g % Q
            g = g * coset_shift % Q
    logn = n.bit_length() - 1
    # bit reversal
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    root = ROOT_OF_UNITY_INV if inverse else ROOT_OF_UNITY
    for s in range(1, logn + 1):
        m = 1 << s
        wm = pow(root, (1 << 32) >> s, Q)
        half = m >> 1
        tw = [1] * half
        for i in range(1, half):
            tw[i] = tw[i - 1] * wm % Q
        for k in range(0, n, m):
            for i in range(half):
                t = a[k + i + half] * tw[i] % Q
                u = a[k + i]
                a[k + i] = (u + t) % Q
                a[k + i + half] = (u - t) % Q
    if inverse:
        ninv = pow(n, -1, Q)
        a = [x * ninv % Q for x in a]
        if coset_shift is not None:
            gi = pow(coset_shift, -1, Q)
            g = 1
            for i in range(n):
                a[i] = a[i] * g % Q
                g = g * gi % Q
    return a


# --------------------------------------------------------------------------------------
# Polynomial: restatement of src/polynomial.rs (the operations on the hot path + neighbours)
# --------------------------------------------------------------------------------------
MONOMIAL = "Monomial"
LAGRANGE = "Lagrange"


class Polynomial:
    """polynomial.rs:14-17"""

    def __init__(self, values, basis=MONOMIAL):
        self.values = [v % Q for v in values]
        self.basis = basis

    def __eq__(self, other):
        return self.basis == other.basis and self.values == other.values

    def __repr__(self):
        return f"Polynomial({self.basis}, {[hex(v) for v in self.values]})"

    def coeffs_evaluate(self, x: int) -> int:
        """polynomial.rs:34-45"""
        assert self.basis == MONOMIAL
        return sum(c * pow(x, i, Q) for i, c in enumerate(self.values)) % Q

    def ntt(self):
        """polynomial.rs:47-51"""
        assert self.basis == MONOMIAL
        return Polynomial(ntt_381(self.values), LAGRANGE)

    def i_ntt(self):
        """polynomial.rs:52-55"""
        assert self.basis == LAGRANGE
        return Polynomial(i_ntt_381(self.values), MONOMIAL)

    def __add__(self, rhs):
        """polynomial.rs:57-117"""
        if isinstance(rhs, int):
            if self.basis == MONOMIAL:
                v = list(self.values)
                v[0] = (v[0] + rhs) % Q
                return Polynomial(v, MONOMIAL)
            return Polynomial([(x + rhs) % Q for x in self.values], LAGRANGE)
        assert self.basis == rhs.basis
        if self.basis == LAGRANGE:
            assert len(self.values) == len(rhs.values)
            return Polynomial([(a + b) % Q for a, b in zip(self.values, rhs.values)], LAGRANGE)
        n = max(len(self.values), len(rhs.values))
        a = self.values + [0] * (n - len(self.values))
        b = rhs.values + [0] * (n - len(rhs.values))
        return Polynomial([(x + y) % Q for x, y in zip(a, b)], MONOMIAL)

    def __sub__(self, rhs):
        """polynomial.rs:119-174 (Monomial branch; Sub<Scalar> touches coefficient 0 only)"""
        if isinstance(rhs, int):
            assert self.basis == MONOMIAL
            v = list(self.values)
            v[0] = (v[0] - rhs) % Q
            return Polynomial(v, MONOMIAL)
        assert self.basis == rhs.basis
        if self.basis == LAGRANGE:
            assert len(self.values) == len(rhs.values)
            return Polynomial([(a - b) % Q for a, b in zip(self.values, rhs.values)], LAGRANGE)
        n = max(len(self.values), len(rhs.values))
        a = self.values + [0] * (n - len(self.values))
        b = rhs.values + [0] * (n - len(rhs.values))
        return Polynomial([(x - y) % Q for x, y in zip(a, b)], MONOMIAL)

    def scale(self, s: int):
        """Mul<Scalar>.  polynomial.rs:176-187"""
        return Polynomial([v * s % Q for v in self.values], self.basis)

    def __mul__(self, rhs):
        """impl Mul for Polynomial (polynomial.rs:189-276).  Monomial x Monomial: evaluate both on
        the domain of size D = find_next_power_of_two(deg a, deg b), pointwise product, i_ntt_381,
        truncate to la+lb-1 coefficients (trailing zeros kept)."""
        if isinstance(rhs, int):
            return self.scale(rhs)
        assert self.basis == rhs.basis
        if self.basis == LAGRANGE:
            raise NotImplementedError("todo!() in the reference (polynomial.rs:274-276)")
        la, lb = len(self.values), len(rhs.values)
        d = find_next_power_of_two(la - 1, lb - 1)
        ea = ntt_fast(self.values + [0] * (d - la))
        eb = ntt_fast(rhs.values + [0] * (d - lb))
        prod = [x * y % Q for x, y in zip(ea, eb)]
        coeffs = ntt_fast(prod, inverse=True)
        return Polynomial(coeffs[: la + lb - 1], MONOMIAL)


# --------------------------------------------------------------------------------------
# deterministic input generator shared by tests / bench: SplitMix64
# --------------------------------------------------------------------------------------
def splitmix64_stream(seed: int, count: int) -> list[int]:
    x = seed & MASK64
    out = []
    for _ in range(count):
        x = (x + 0x9E3779B97F4A7C15) & MASK64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        z ^= z >> 31
        out.append(z)
    return out


def random_fr(seed: int, n: int) -> list[int]:
    """n uniform Fr elements: 8 u64 per scalar -> 512-bit LE -> mod q (from_bytes_wide semantics)"""
    s = splitmix64_stream(seed, 8 * n)
    out = []
    for i in range(n):
        v = 0
        for j in range(8):
            v |= s[8 * i + j] << (64 * j)
        out.append(v % Q)
    return out
