"""Pins for oracle/plonk.py (the callers of the hot path: prover rounds, transcript, pre-processing).

The reference pins none of this by bytes (random blinding, no transcript test); the pins available are
merlin's published conformance vector, SURVEY.md 8c's survey-time model values for the reference's own
test program (tests/verify_proof_test.rs, n = 8, tau = 101) with blinding 1..11, and self-verification."""
import pytest

from oracle import bls12_381 as O
from oracle import plonk as P


def test_merlin_conformance_vector():
    t = P.MerlinTranscript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == \
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_div_semantics_and_quirk():
    # (x^2 - 1) / (x - 1) = x + 1
    assert P.p_div([O.Q - 1, 0, 1], [O.Q - 1, 1]) == [1, 1]
    # remainder is dropped: (x^2 + 1) / (x - 1) = x + 1 (rem 2)
    assert P.p_div([1, 0, 1], [O.Q - 1, 1]) == [1, 1]
    # the reference's quirk (polynomial.rs:376): interior zero quotient coefficients vanish
    assert P.p_div([O.Q - 1, 0, 0, 0, 1], [O.Q - 1, 0, 1]) == [1, 1]      # true quotient is x^2 + 1 = [1, 0, 1]
    # trailing zeros are stripped from both operands
    assert P.p_div([O.Q - 1, 0, 1, 0, 0], [O.Q - 1, 1, 0]) == [1, 1]


def test_reference_test_program_preprocessing():
    prog, wit, pub = P.reference_test_circuit()
    ql, qr, qm, qo, qc = prog.selectors()
    m1 = O.Q - 1
    assert ql == [1, 0, 0, 0, 0, 0, 0, 0]
    assert qr == [0, m1, 0, 0, 0, 0, 0, 0]
    assert qm == [0, m1, m1, 0, 0, 0, 0, 0]
    assert qo == [0, 1, 1, 0, 0, 0, 0, 0]
    assert qc == [0] * 8
    assert prog.public_vars() == ["e"]
    s1, s2, s3 = prog.sigmas()
    # sigma is a permutation of the 3n cell labels {w^i, 2 w^i, 3 w^i}
    roots = O.roots_of_unity(8)
    labels = sorted([r * k % O.Q for r in roots for k in (1, 2, 3)])
    assert sorted(s1 + s2 + s3) == labels


def test_reference_test_program_proof_kat():
    """n = 8 circuit, tau = 101, blinding b1..b11 = 1..11"""
    prog, wit, pub = P.reference_test_circuit()
    srs = O.generate_srs_points(14, 101)  # Setup::generate_srs(8 + 6, tau)  verify_proof_test.rs:16
    be = P.OracleBackend(srs)
    tr = {}
    proof = P.prove(prog, wit, list(range(1, 12)), be, trace=tr)
    assert tr["beta"] >> 224 == 0x2C7978C1 and tr["beta"] & 0xFFFFFFFF == 0x3A280AE7
    assert tr["gamma"] >> 224 == 0x3C24F86B and tr["gamma"] & 0xFFFFFFFF == 0xA54F3AA3
    assert tr["alpha"] >> 224 == 0x5CDDF463 and tr["alpha"] & 0xFFFFFFFF == 0x3DECB41A
    assert tr["zeta"] >> 224 == 0x3E269367 and tr["zeta"] & 0xFFFFFFFF == 0x5EA53A33
    assert tr["nu"] >> 224 == 0x4C4B0F3B and tr["nu"] & 0xFFFFFFFF == 0xBF089026
    assert tr["mu"] >> 224 == 0x677CE1C1 and tr["mu"] & 0xFFFFFFFF == 0x8E674AC2
    assert O.g1_to_compressed(proof.a_1).hex() == \
        "80af34b1403d584b5730c2e4a87e9b2e7a328bbb67c90afcac06dafc36f515b7579d0dd82cfa53f05856cb3e10fd3cd5"
    assert O.g1_to_compressed(proof.z_1).hex().startswith("951e3f27e8db86ea")
    assert O.g1_to_compressed(proof.t_hi_1).hex().startswith("9000b080a760858e")
    assert O.g1_to_compressed(proof.w_zeta_1).hex().startswith("a44251547e81b86d")
    assert proof.a_bar >> 224 == 0x517151CE and proof.a_bar & 0xFFFFFFFF == 0xEFEFF9D2
    assert proof.z_omega_bar >> 224 == 0x1CAAAB74 and proof.z_omega_bar & 0xFFFFFFFF == 0x00344E58
    assert len(proof.to_bytes()) == 624
    assert proof.sha256() == "479cc377c535fd831b5fcaf30af5c2756c535a3ddbc20589ab6759843e974967"
    # verifier.rs:80-192 accepts it (trapdoor form), and rejects tampered proofs / wrong public input
    assert P.verify(prog, proof, pub, 101, be.commit)
    assert not P.verify(prog, proof, [81], 101, be.commit)
    proof.a_bar = (proof.a_bar + 1) % O.Q
    assert not P.verify(prog, proof, pub, 101, be.commit)


@pytest.mark.parametrize("n,used", [(16, 9), (32, 32)])
def test_synthetic_circuit_proves_and_verifies(n, used):
    prog, wit, pub = P.synthetic_circuit(n, used, seed=n)
    srs = O.generate_srs_points(n + 6, 101)
    be = P.OracleBackend(srs, reference_msm=False)
    blinding = O.random_fr(42, 11)  # SplitMix64(seed 42) -> from_bytes_wide, SURVEY 8d C1
    proof = P.prove(prog, wit, blinding, be)
    assert P.verify(prog, proof, pub, 101, be.commit)
    # reference MSM semantics (64 x 4-bit windows) give the same commitments
    be2 = P.OracleBackend(srs, reference_msm=True)
    assert be2.commit(O.random_fr(1, n + 6)) == be.commit(O.random_fr(1, n + 6))
