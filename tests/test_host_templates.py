"""CPU check of the CUDA field / curve templates (csrc/ff.cuh, csrc/ec.cuh).

The templates compile for the host with the PTX carry flag emulated in C, so the exact limb
schedule the kernels run (interleaved even/odd Montgomery rows, XYZZ formulas with degenerate
cases) is checked against the Python oracle without a GPU.  The device build of the same
templates is checked again by the -m gpu parity tests.
"""
import os
import random
import subprocess

import pytest

from oracle import bls12_381 as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "ff_host_check.cpp")
EXE = os.path.join(ROOT, "build", "ff_host_check")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if (not os.path.exists(EXE)) or os.path.getmtime(EXE) < max(
            os.path.getmtime(SRC),
            os.path.getmtime(os.path.join(ROOT, "baby-plonk-rust_b200", "csrc", "ff.cuh")),
            os.path.getmtime(os.path.join(ROOT, "baby-plonk-rust_b200", "csrc", "ec.cuh"))):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-x", "c++", SRC, "-o", EXE])
    return EXE


def run(exe, lines):
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
    return out.stdout.strip().split("\n")


def hx(v):
    return "%x" % v


def test_field_ops(exe):
    rng = random.Random(1)
    lines, exp = [], []
    for name, mod, R in (("fr", O.Q, O.FR_R), ("fp", O.P, O.FP_R)):
        Rinv = pow(R, -1, mod)
        vals = [0, 1, 2, mod - 1, mod - 2, R, (1 << 32) - 1, (1 << 64) - 1, mod >> 1]
        vals += [rng.randrange(mod) for _ in range(40)]
        for a in vals:
            for b in rng.sample(vals, 6) + [a, mod - 1 - a if a != mod - 1 else 0]:
                lines.append(f"{name} mul {hx(a)} {hx(b)}"); exp.append(a * b * Rinv % mod)
                lines.append(f"{name} mulcc {hx(a)} {hx(b)}"); exp.append(a * b * Rinv % mod)
                # lazy addition / subtraction: operands anywhere in [0, 2 mod), result reduced once
                for da, db in ((0, 0), (mod, 0), (0, mod), (mod, mod)):
                    lines.append(f"{name} addlazy {hx(a + da)} {hx(b + db)}"); exp.append((a + b) % mod)
                    lines.append(f"{name} sublazy {hx(a + da)} {hx(b + db)}"); exp.append((a - b) % mod)
                if name == "fr":   # lazy product: canonical twiddle FIRST, value in [0, 2q) second (see mul_cc)
                    lines.append(f"{name} mullazy {hx(b)} {hx(a + mod)}"); exp.append(a * b * Rinv % mod)
                if name == "fp":   # lazy product (no final subtraction) on operands up to 2p - 1, then one reduction
                    lines.append(f"{name} mullazy {hx(a + mod)} {hx(b + mod)}"); exp.append(a * b * Rinv % mod)
                lines.append(f"{name} add {hx(a)} {hx(b)}"); exp.append((a + b) % mod)
                lines.append(f"{name} sub {hx(a)} {hx(b)}"); exp.append((a - b) % mod)
            lines.append(f"{name} sqr {hx(a)}"); exp.append(a * a * Rinv % mod)
            lines.append(f"{name} neg {hx(a)}"); exp.append((-a) % mod)
            lines.append(f"{name} from_mont {hx(a)}"); exp.append(a * Rinv % mod)
            lines.append(f"{name} to_mont {hx(a)}"); exp.append(a * R % mod)
        # Montgomery inverse: inv(aR) = a^-1 R  ->  on raw value v: v^-1 R^2  (division-step inversion)
        inv_vals = vals + [rng.randrange(mod) for _ in range(400)] + [1 << k for k in range(0, mod.bit_length() - 1, 7)] \
            + [mod - (1 << k) for k in range(0, mod.bit_length() - 1, 11)] + [(mod + 1) // 2, (mod - 1) // 2, 3, mod - 3]
        for a in inv_vals:
            lines.append(f"{name} inv {hx(a)}")
            exp.append(pow(a, -1, mod) * R * R % mod if a else 0)
    got = run(exe, lines)
    assert len(got) == len(exp)
    for l, g, e in zip(lines, got, exp):
        assert int(g, 16) == e, l


def test_fp_reference_kats(exe):
    import json
    kats = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))
    L = lambda xs: O.limbs_to_int([int(x, 16) for x in xs])
    t = kats["fp"]["test_multiplication"]
    got = run(exe, [f"fp mul {hx(L(t['a']))} {hx(L(t['b']))}"])
    assert int(got[0], 16) == L(t["a_times_b"])
    t = kats["fp"]["test_squaring"]
    assert int(run(exe, [f"fp sqr {hx(L(t['a']))}"])[0], 16) == L(t["a_squared"])
    t = kats["fp"]["test_inversion"]
    assert int(run(exe, [f"fp inv {hx(L(t['a']))}"])[0], 16) == L(t["a_inv"])


def M(v):
    return hx(v * O.FP_R % O.P)


def xyzz_of(pt, z):
    """XYZZ Montgomery hex coordinates of an affine point with Z-scale z (identity: zeros)"""
    if pt is None:
        return "0 0 0 0"
    zz, zzz = z * z % O.P, z * z * z % O.P
    return f"{M(pt[0] * zz)} {M(pt[1] * zzz)} {M(zz)} {M(zzz)}"


def parse_xyzz(line):
    X, Y, ZZ, ZZZ = (int(t, 16) * O.FP_RINV % O.P for t in line.split())
    if ZZ == 0:
        return None
    assert pow(ZZ, 3, O.P) == ZZZ * ZZZ % O.P
    return (X * pow(ZZ, -1, O.P) % O.P, Y * pow(ZZZ, -1, O.P) % O.P)


def test_curve_ops(exe):
    rng = random.Random(2)
    G = O.G1_GEN
    pts = [None] + [O.g1_mul(G, k) for k in (1, 2, 3, 5, 1000, O.Q - 1, O.Q - 2, O.Q - 5, 12345678901234567890)]
    lines, exp = [], []
    for a in pts:
        za = rng.randrange(1, O.P)
        lines.append(f"g1 dbl {xyzz_of(a, za)}"); exp.append(O.g1_double(a))
        for b in pts:
            zb = rng.randrange(1, O.P)
            lines.append(f"g1 add {xyzz_of(a, za)} {xyzz_of(b, zb)}"); exp.append(O.g1_add(a, b))
            bx, by = (b if b is not None else (0, 0))
            lines.append(f"g1 madd {xyzz_of(a, za)} {M(bx)} {M(by)}"); exp.append(O.g1_add(a, b))
    got = run(exe, lines)
    assert len(got) == len(exp)
    for l, g, e in zip(lines, got, exp):
        assert parse_xyzz(g) == e, l
    # conversions to affine
    lines, exp = [], []
    for a in pts:
        za = rng.randrange(1, O.P)
        lines.append(f"g1 to_affine {xyzz_of(a, za)}"); exp.append(a)
        if a is None:
            lines.append(f"g1 proj_to_affine 0 {M(za)} 0")
        else:
            lines.append(f"g1 proj_to_affine {M(a[0] * za)} {M(a[1] * za)} {M(za)}")
        exp.append(a)
    got = run(exe, lines)
    for l, g, e in zip(lines, got, exp):
        x, y = (int(t, 16) * O.FP_RINV % O.P for t in g.split())
        assert ((x, y) if (x, y) != (0, 0) else None) == e, l


def test_batched_affine_pair_addition(exe):
    """affine_add_prepare / affine_add_finish (the per-pair halves of the batched-affine bucket additions) on every
    case: generic, doubling, opposite points, identity operands"""
    G = O.G1_GEN
    pts = [None] + [O.g1_mul(G, k) for k in (1, 2, 3, 7, 1000, O.Q - 1, O.Q - 2, O.Q - 7, 98765432109876543210)]
    kinds = {"copy_p": 0, "copy_q": 1, "add": 2, "dbl": 3, "inf": 4}
    lines, exp = [], []
    for a in pts:
        for b in pts:
            ax, ay = a if a is not None else (0, 0)
            bx, by = b if b is not None else (0, 0)
            lines.append(f"g1 aff_add {M(ax)} {M(ay)} {M(bx)} {M(by)}")
            kind = "copy_p" if b is None else "copy_q" if a is None else "add" if a[0] != b[0] else \
                "dbl" if a[1] == b[1] else "inf"
            exp.append((kinds[kind], O.g1_add(a, b)))
    got = run(exe, lines)
    for l, g, (k, e) in zip(lines, got, exp):
        t = g.split()
        x, y = (int(v, 16) * O.FP_RINV % O.P for v in t[1:])
        assert int(t[0]) == k, l
        assert ((x, y) if (x, y) != (0, 0) else None) == e, l


def test_affine_pair_to_xyzz(exe):
    """xyzz_from_affine_sum (first level of the bucket-reduction tree: both operands affine, 4M + 2S)"""
    G = O.G1_GEN
    pts = [O.g1_mul(G, k) for k in (1, 2, 3, 7, 1000, O.Q - 2, 98765432109876543210)]
    lines, exp = [], []
    for a in pts:
        for b in pts:
            if a[0] == b[0]:
                continue
            lines.append(f"g1 aff_sum {M(a[0])} {M(a[1])} {M(b[0])} {M(b[1])}")
            exp.append(O.g1_add(a, b))
    for l, g, e in zip(lines, run(exe, lines), exp):
        assert parse_xyzz(g) == e, l


def test_inversion_terminates_on_non_canonical_limbs(exe):
    """limbs >= the modulus (p, 2p, all ones) must not hang the division-step loop (the binary Euclid did)"""
    lines = [f"fp inv {hx(O.P)}", f"fp inv {hx(2 * O.P)}", f"fp inv {hx((1 << 384) - 1)}",
             f"fr inv {hx(O.Q)}", f"fr inv {hx(2 * O.Q)}", f"fr inv {hx((1 << 256) - 1)}"]
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=20)
    assert out.returncode == 0 and len(out.stdout.strip().split("\n")) == len(lines)


def test_dedicated_fp_squaring(exe):
    """sqr_cc (off-diagonal products doubled + reduction-only rows) against big-int squaring: random values and
    carry-heavy limb patterns (all-ones limbs, single bits, p - small)"""
    rng = random.Random(7)
    mod, R = O.P, O.FP_R
    Rinv = pow(R, -1, mod)
    vals = [rng.randrange(mod) for _ in range(1500)]
    vals += [mod - k for k in range(1, 40)] + [k for k in range(40)]
    vals += [(1 << b) % mod for b in range(0, 381, 7)] + [((1 << b) - 1) % mod for b in range(1, 381, 5)]
    for _ in range(300):   # limbs drawn from {0, 1, 0xffffffff, 0x80000000, random}
        v = 0
        for k in range(12):
            limb = rng.choice([0, 1, 0xFFFFFFFF, 0x80000000, 0xFFFFFFFE, rng.getrandbits(32)])
            v |= limb << (32 * k)
        vals.append(v % mod)
    got = run(exe, [f"fp sqr {hx(a)}" for a in vals])
    for a, g in zip(vals, got):
        assert int(g, 16) == a * a * Rinv % mod, hx(a)
