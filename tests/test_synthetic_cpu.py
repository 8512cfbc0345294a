"""baby-plonk-rust_b200/synthetic.py (benchmark circuit columns) against the oracle's restatement of the
reference's pre-processing (src/program.rs:51-147) on the same gate list."""
import importlib

import pytest

from oracle import bls12_381 as O
from oracle import plonk as P
from tests._bpk import bpk

S = importlib.import_module("baby-plonk-rust_b200.synthetic")


@pytest.mark.parametrize("n,gates", [(8, 3), (16, 16), (64, 41)])
def test_chain_circuit_columns_match_program_preprocessing(n, gates):
    circ = S.chain_circuit(n, gates, seed=5, native=False)
    A, B, C = circ["ints"]["wires"]
    wit = {"x0": A[1], "out": circ["public_inputs"][0]}
    rows = [P.Gate.public_input("out")]
    m = gates - 1
    for k in range(1, m + 1):
        left = "x0" if k == 1 else "c%d" % (k - 1)
        res = "out" if k == m else "c%d" % k
        wit["y%d" % k] = B[k]
        wit[res] = C[k]
        rows.append(P.Gate.mul(res, left, "y%d" % k) if k & 1 else P.Gate.add(res, left, "y%d" % k))
    prog = P.Program(rows, n)
    assert prog.selectors() == circ["ints"]["selectors"]
    assert prog.sigmas() == circ["ints"]["sigmas"]
    assert prog.public_vars() == ["out"]
    # the witness satisfies every gate
    ql, qr, qm, qo, qc = circ["ints"]["selectors"]
    for i in range(n):
        pi = -A[0] if i == 0 else 0
        assert (ql[i] * A[i] + qr[i] * B[i] + qm[i] * A[i] * B[i] + qo[i] * C[i] + qc[i] + pi) % O.Q == 0
    assert bpk.scalars_to_ints(circ["wires"][2]) == C
    assert (S.mont_array(A) == bpk.scalars_from_ints(A)).all()


@pytest.mark.parametrize("n,gates,seed", [(2, 2, 1), (8, 3, 5), (16, 16, 2), (64, 41, 7), (1024, 1021, 2), (4096, 100, 1999)])
def test_native_generator_equals_the_interpreted_one(n, gates, seed):
    """bpk_synthetic_chain_circuit (host C++, what bench.py and the full-size tests use) == the interpreted generator
    that the test above pins to the oracle's program pre-processing"""
    import numpy as np
    a = S.chain_circuit(n, gates, seed=seed, native=True)
    b = S.chain_circuit(n, gates, seed=seed, native=False)
    assert a["public_inputs"] == b["public_inputs"]
    for key in ("selectors", "sigmas", "wires"):
        for x, y in zip(a[key], b[key]):
            assert np.array_equal(x, y), key
