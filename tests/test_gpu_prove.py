"""config[0] of BASELINE.json end to end: the reference's prover algebra (oracle/plonk.py restates
src/prover.rs) run once on the CPU oracle and once with the three hot-path surfaces -- Setup::commit,
i_ntt_381, impl Mul for Polynomial -- served by the GPU library through its C ABI.  The two proofs must be
identical byte for byte (624 bytes: 9 compressed commitments + 6 evaluations), and for the reference's own
test program with blinding 1..11 equal the SHA-256 recorded in SURVEY.md 8c."""
import numpy as np
import pytest

from oracle import bls12_381 as O
from oracle import plonk as P
from tests._bpk import bpk

pytestmark = pytest.mark.gpu


class GpuBackend:
    """the drop-in: every commitment, inverse NTT and polynomial product goes through libbpk.so"""

    def __init__(self, ctx, powers, tau, precompute=False):
        self.ctx = ctx
        self.setup = bpk.Setup.generate_srs(powers, tau, ctx)     # Setup::generate_srs (setup.rs:12-31)
        if precompute:
            self.setup.precompute(0)
        self.calls = {"commit": 0, "i_ntt": 0, "mul": 0}

    def i_ntt(self, values):
        self.calls["i_ntt"] += 1
        return bpk.scalars_to_ints(bpk.i_ntt_381(bpk.scalars_from_ints(values), self.ctx))

    def mul(self, a, b):
        self.calls["mul"] += 1
        pa = bpk.Polynomial.from_ints(a, ctx=self.ctx)
        pb = bpk.Polynomial.from_ints(b, ctx=self.ctx)
        return (pa * pb).to_ints()

    def commit(self, coeffs):
        self.calls["commit"] += 1
        return bpk.point_to_affine(self.setup.commit(bpk.Polynomial.from_ints(coeffs)))


@pytest.fixture(scope="module")
def ctx():
    c = bpk.Context(0)
    yield c
    c.close()


def test_reference_test_program_proof_bytes(ctx):
    """tests/verify_proof_test.rs:13-50: setup(8 + 6, tau = 101), prove, verify"""
    prog, wit, pub = P.reference_test_circuit()
    gpu = GpuBackend(ctx, 14, 101)
    srs = [bpk.point_to_affine(p) for p in gpu.setup.powers_of_x()]
    assert srs == O.generate_srs_points(14, 101)
    cpu = P.OracleBackend(srs)
    blinding = list(range(1, 12))
    proof_gpu = P.prove(prog, wit, blinding, gpu)
    proof_cpu = P.prove(prog, wit, blinding, cpu)
    assert proof_gpu.to_bytes() == proof_cpu.to_bytes()
    assert proof_gpu.sha256() == "479cc377c535fd831b5fcaf30af5c2756c535a3ddbc20589ab6759843e974967"
    # 9 commitments and 16 polynomial products per proof; 15 distinct inverse NTTs (the reference
    # repeats 8 of them in round 5: 23 calls)
    assert gpu.calls == {"commit": 9, "i_ntt": 15, "mul": 16}
    # the verifier (CPU in the reference) accepts it; its 8 pre-processing commitments also use the drop-in
    assert P.verify(prog, proof_gpu, pub, 101, gpu.commit)
    assert not P.verify(prog, proof_gpu, [81], 101, gpu.commit)


@pytest.mark.parametrize("n,used,seed", [(16, 10, 7), (64, 64, 8), (256, 200, 9)])
def test_synthetic_circuit_proofs_are_byte_identical(ctx, n, used, seed):
    prog, wit, pub = P.synthetic_circuit(n, used, seed=seed)
    gpu = GpuBackend(ctx, n + 6, 101, precompute=(n >= 64))
    srs = [bpk.point_to_affine(p) for p in gpu.setup.powers_of_x()]
    cpu = P.OracleBackend(srs, reference_msm=False)
    blinding = O.random_fr(42, 11)
    proof_gpu = P.prove(prog, wit, blinding, gpu)
    proof_cpu = P.prove(prog, wit, blinding, cpu)
    assert proof_gpu.to_bytes() == proof_cpu.to_bytes()
    assert P.verify(prog, proof_gpu, pub, 101, gpu.commit)


def test_prove_at_scale_self_verifies(ctx):
    """n = 2^12 gates: too slow for the oracle backend, so the GPU-backed proof is checked by the verifier
    equation (trapdoor form) -- any wrong commitment, NTT output or product coefficient breaks it"""
    n = 1 << 12
    prog, wit, pub = P.synthetic_circuit(n, n - 5, seed=3)
    gpu = GpuBackend(ctx, n + 6, 101, precompute=True)
    proof = P.prove(prog, wit, O.random_fr(43, 11), gpu)
    assert P.verify(prog, proof, pub, 101, gpu.commit)
    proof.s1_bar = (proof.s1_bar + 1) % O.Q
    assert not P.verify(prog, proof, pub, 101, gpu.commit)
