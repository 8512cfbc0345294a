"""Product-side Fiat-Shamir transcript (baby-plonk-rust_b200/transcript.py) against merlin's published
conformance vector and against the oracle's independent restatement (oracle/plonk.py)."""
import importlib

from oracle import bls12_381 as O
from oracle import plonk as P
from tests._bpk import bpk  # noqa: F401  (puts the repository root on sys.path)

T = importlib.import_module("baby-plonk-rust_b200.transcript")


def test_keccak_matches_oracle_permutation():
    state = bytearray(range(200))
    lanes = [int.from_bytes(state[8 * i:8 * i + 8], "little") for i in range(25)]
    P.keccak_f1600(state)
    out = T.keccak_f1600(lanes)
    assert b"".join(v.to_bytes(8, "little") for v in out) == bytes(state)


def test_merlin_conformance_vector():
    """merlin's own test `equivalence_simple` (transcript.rs of merlin 3.0.0)"""
    t = T.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == \
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_long_messages_cross_the_rate_boundary():
    a, b = T.Transcript(b"x"), P.MerlinTranscript(b"x")
    for k in (0, 1, 165, 166, 167, 500):
        msg = bytes((i * 7 + k) & 0xFF for i in range(k))
        a.append_message(b"m", msg)
        b.append_message(b"m", msg)
        assert a.challenge_bytes(b"c", 200) == b.challenge_bytes(b"c", 200)


def test_plonk_transcript_schedule_matches_oracle():
    """src/transcript.rs label schedule: same challenges for the same appended points / scalars"""
    a, b = T.PlonkTranscript(), P.PlonkTranscript()
    pts = [O.g1_mul(O.G1_GEN, k) for k in (5, 77, 123456789)]
    for lab, pt in zip((b"a_1", b"b_1", b"c_1"), pts):
        a.append_point(lab, O.g1_to_compressed(pt))
        b.append_point(lab, pt)
    for lab in (b"beta", b"gamma", b"z_1"):
        assert a.get_and_append_challenge(lab) == b.get_and_append_challenge(lab)
    a.append_scalar(b"a_eval", 12345)
    b.append_scalar(b"a_eval", 12345)
    assert a.get_and_append_challenge(b"nu") == b.get_and_append_challenge(b"nu")


def test_native_keccak_matches_the_python_permutation():
    """bpk_keccak_f1600 (libbpk.so, host code) is what the transcript calls; the plain-Python permutation is its check"""
    import ctypes
    import random

    fn = T._native_keccak()
    assert fn is not None, "libbpk.so must be built (python __graft_entry__.py build)"
    rng = random.Random(5)
    for _ in range(20):
        state = bytearray(rng.getrandbits(8) for _ in range(200))
        lanes = [int.from_bytes(state[8 * i:8 * i + 8], "little") for i in range(25)]
        fn((ctypes.c_uint8 * 200).from_buffer(state))
        assert bytes(state) == b"".join(v.to_bytes(8, "little") for v in T.keccak_f1600(lanes))
