"""The reference's own hot-path unit tests, restated in C++ on the host mirror of the Rust API
(baby-plonk-rust_b200/host/baby_plonk.hpp -> C ABI -> CUDA): tests/cpp/reference_style_tests.cpp."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _exe(name="reference_style_tests"):
    import __graft_entry__ as ge
    return ge.build_cpp_host_tests(name)


@pytest.mark.gpu
def test_reference_style_cpp_tests_pass_on_gpu():
    r = subprocess.run([_exe()], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASSED" in r.stdout
    for name in ("test_generate_srs", "test_monomial_commit", "test_ntt_round_trip", "test_polynomial_mul",
                 "test_root_of_unity", "test_bucket_msm"):
        assert name + " ... ok" in r.stdout


def test_cpp_host_layer_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_exe()], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stdout


def test_cpp_transcript_and_encodings_match_oracle():
    """host/plonk_prover.hpp on the CPU: merlin's conformance vector, a challenge of the PLONK label schedule and the
    compressed encodings, against oracle/plonk.py"""
    from oracle import bls12_381 as O
    from oracle import plonk as P

    r = subprocess.run([_exe("transcript_check")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.split()
    assert lines[0] == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    t = P.PlonkTranscript()
    t.append_point(b"a_1", O.G1_GEN)
    t.append_scalar(b"a_eval", 12345)
    assert lines[1] == O.fr_to_bytes(t.get_and_append_challenge(b"beta")).hex()
    assert lines[2] == O.g1_to_compressed(O.G1_GEN).hex()
    assert lines[3] == O.g1_to_compressed(O.g1_mul(O.G1_GEN, O.Q - 5)).hex()


def _write_instance(path, prog, wit, pub, blinding, powers, tau, cache, reps):
    import importlib
    import struct

    bpk = importlib.import_module("baby-plonk-rust_b200")
    from oracle import bls12_381 as O

    n = prog.n
    pad = n - len(prog.gates)
    cols = list(prog.selectors()) + list(prog.sigmas())
    for k in range(3):
        cols.append([wit[g.wires[k]] % O.Q if g.wires[k] is not None else 0 for g in prog.gates] + [0] * pad)
    with open(path, "wb") as f:
        f.write(struct.pack("<5Q", n, len(pub), powers, 1 if cache else 0, reps))
        f.write(bpk.scalars_from_ints([tau]).tobytes())
        for c in cols:
            f.write(bpk.scalars_from_ints(c).tobytes())
        f.write(bpk.scalars_from_ints(pub).tobytes())
        f.write(bpk.scalars_from_ints(blinding).tobytes())


@pytest.mark.gpu
def test_cpp_device_prover_proofs_are_byte_identical(tmp_path):
    """the C++ device-resident prover (host/plonk_prover.hpp over the C ABI only) against oracle/plonk.py::prove:
    the reference's test program with SURVEY 8c's SHA-256, and a 64-row circuit proved twice from cached
    pre-processed polynomials"""
    import hashlib

    from oracle import bls12_381 as O
    from oracle import plonk as P

    prog, wit, pub = P.reference_test_circuit()
    inst = str(tmp_path / "kat.bin")
    _write_instance(inst, prog, wit, pub, list(range(1, 12)), 14, 101, False, 1)
    r = subprocess.run([_exe("device_prover_main"), inst], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    proof = bytes.fromhex(r.stdout.split()[0])
    assert hashlib.sha256(proof).hexdigest() == "479cc377c535fd831b5fcaf30af5c2756c535a3ddbc20589ab6759843e974967"

    prog, wit, pub = P.synthetic_circuit(64, 50, seed=11)
    blinding = O.random_fr(44, 11)
    inst = str(tmp_path / "syn.bin")
    _write_instance(inst, prog, wit, pub, blinding, 70, 101, True, 2)
    r = subprocess.run([_exe("device_prover_main"), inst], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    want = P.prove(prog, wit, blinding, P.OracleBackend(O.generate_srs_points(70, 101), reference_msm=False)).to_bytes()
    got = r.stdout.split()
    assert len(got) == 2 and bytes.fromhex(got[0]) == want and bytes.fromhex(got[1]) == want
    # a broken copy constraint trips the reference's assertion
    wit = dict(wit)
    wit["x3"] = (wit["x3"] + 1) % O.Q
    _write_instance(inst, prog, wit, pub, blinding, 70, 101, False, 1)
    r = subprocess.run([_exe("device_prover_main"), inst], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "PANIC" in r.stdout
