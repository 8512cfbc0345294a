"""Synthetic circuits for benchmarking the device prover at sizes where no circuit source exists
(BASELINE.json configs[3]: 2^20 gates).

One family, pre-processed the way src/program.rs:51-147 would (selector columns, copy-constraint
permutation as sigma columns), built directly as columns so that 2^20 rows take seconds:

    row 0        "out public"                 L = 1,  wires (out, -, -)
    row k >= 1   "c_k <== c_{k-1} * y_k"      M = -1, O = 1           (k odd)
                 "c_k <== c_{k-1} + y_k"      L = -1, R = -1, O = 1   (k even)
    c_0 = x0, c_m = out; every y_k is a fresh variable; rows m+1 .. n-1 are empty.

Copy constraints: c_k sits in (O, k) and (L, k+1); out sits in (L, 0) and (O, m); all unused cells form one
cycle in row-major order, exactly as program.rs links the cells of the `None` variable.
"""
from __future__ import annotations

import numpy as np

from . import FR_MODULUS, _FR_R, roots_of_unity

Q = FR_MODULUS


def mont_array(values) -> np.ndarray:
    """canonical ints -> uint64[n, 4] Montgomery limbs, via one bytes join (fast path for 2^20 rows)"""
    raw = b"".join(((v % Q) * _FR_R % Q).to_bytes(32, "little") for v in values)
    return np.frombuffer(raw, dtype=np.uint64).reshape(-1, 4).copy()


def chain_circuit(n: int, gates: int, seed: int = 1, native: bool = True):
    """returns dict(selectors=[QL,QR,QM,QO,QC], sigmas=[S1,S2,S3], wires=[A,B,C] (all uint64[n,4] Montgomery),
    public_inputs=[out]).  native: built by the library's host routine (bpk_synthetic_chain_circuit: seconds at 2^24
    rows); otherwise by the interpreted generator below, which also returns ints=dict of the same columns as Python
    ints for cross-checks."""
    if native:
        import ctypes
        from . import load_library
        lib = load_library()
        cols = [np.empty((n, 4), dtype=np.uint64) for _ in range(11)]
        ptrs = (ctypes.c_void_p * 11)(*[c.ctypes.data for c in cols])
        pub = np.zeros(4, dtype=np.uint64)
        if lib.bpk_synthetic_chain_circuit(n, gates, seed, ptrs, pub.ctypes.data) != 0:
            raise ValueError("bad circuit shape")
        out = int.from_bytes(pub.tobytes(), "little")
        return dict(selectors=cols[0:5], sigmas=cols[5:8], wires=cols[8:11], public_inputs=[out])
    m = gates - 1                      # arithmetic rows 1..m
    assert 2 <= gates <= n and n & (n - 1) == 0
    roots = roots_of_unity(n)
    state = (seed * 0x9E3779B97F4A7C15 + 1) & 0xFFFFFFFFFFFFFFFF
    A, B, C = [0] * n, [0] * n, [0] * n
    ql, qr, qm, qo, qc = ([0] * n for _ in range(5))
    c_prev = 3 + seed % 1000
    for k in range(1, m + 1):
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        y = (state >> 20) + 2
        A[k], B[k] = c_prev, y
        if k & 1:
            c_prev = c_prev * y % Q
            qm[k], qo[k] = Q - 1, 1
        else:
            c_prev = (c_prev + y) % Q
            ql[k], qr[k], qo[k] = Q - 1, Q - 1, 1
        C[k] = c_prev
    out = c_prev
    A[0] = out
    ql[0] = 1
    # sigma: start from the identity labels (col + 1) * w^row, then link the cycles
    s = [list(roots), [2 * r % Q for r in roots], [3 * r % Q for r in roots]]
    for k in range(1, m):              # c_k: (O, k) <-> (L, k + 1)
        s[0][k + 1] = 3 * roots[k] % Q
        s[2][k] = roots[k + 1]
    s[0][0] = 3 * roots[m] % Q         # out: (L, 0) <-> (O, m)
    s[2][m] = roots[0]
    unused = [(1, 0), (2, 0)] + [(col, row) for row in range(m + 1, n) for col in range(3)]
    for i, (col, row) in enumerate(unused):
        ncol, nrow = unused[(i + 1) % len(unused)]
        s[ncol][nrow] = roots[row] * (col + 1) % Q
    ints = dict(selectors=[ql, qr, qm, qo, qc], sigmas=s, wires=[A, B, C])
    return dict(selectors=[mont_array(c) for c in ints["selectors"]], sigmas=[mont_array(c) for c in s],
                wires=[mont_array(c) for c in (A, B, C)], public_inputs=[out], ints=ints)
