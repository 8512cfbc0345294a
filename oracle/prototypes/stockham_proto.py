"""Index-logic prototype of the GPU NTT schedule (test infrastructure, not shipped).

Validates, on the CPU oracle field, the exact pass structure csrc/ntt.cu uses:
multi-pass Stockham autosort (natural order in/out) where each pass does R-point sub-FFTs
(bit-reversed placement + radix-2 DIT stages, natural output) on tiles of C consecutive columns,
inter-pass twiddles w_{Ns*R}^{(j mod Ns) r} taken from a two-level table of the 2^L-th root.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bls12_381 as O

Q = O.Q
L = 28
LO_BITS = 13


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def plan(logn):
    if logn <= 10:
        return [logn] if logn > 0 else []
    if logn <= 20:
        a = (logn + 1) // 2
        return [a, logn - a]
    a = (logn + 2) // 3
    b = (logn - a + 1) // 2
    return [a, b, logn - a - b]


def ntt_gpu_schedule(x, inverse=False, small_L=None):
    n = len(x)
    logn = n.bit_length() - 1
    Lx = small_L or L
    root = O.ROOT_OF_UNITY_INV if inverse else O.ROOT_OF_UNITY
    omega = pow(root, 1 << (32 - Lx), Q)            # primitive 2^L-th root
    lo = [pow(omega, i, Q) for i in range(1 << min(LO_BITS, Lx))]
    hi = [pow(omega, i << LO_BITS, Q) for i in range(1 << max(Lx - LO_BITS, 0))]

    def tw(E):  # omega^E, E < 2^L
        return lo[E & ((1 << LO_BITS) - 1)] * hi[E >> LO_BITS] % Q

    data = list(x)
    logNs = 0
    for logR in plan(logn):
        R = 1 << logR
        Ns = 1 << logNs
        out = [0] * n
        stride_in = n >> logR
        shift = Lx - (logNs + logR)                  # omega_{Ns R} = omega^(2^shift)
        twR = [tw(i << (Lx - logR)) for i in range(R // 2)]
        for j in range(n >> logR):
            k = j & (Ns - 1)
            sm = [0] * R
            for r in range(R):
                v = data[j + r * stride_in]
                if logNs:
                    v = v * tw((k * r) << shift) % Q
                sm[bitrev(r, logR)] = v
            for s in range(1, logR + 1):
                half = 1 << (s - 1)
                for b in range(R // 2):
                    i = ((b >> (s - 1)) << s) | (b & (half - 1))
                    t = sm[i + half] * twR[(b & (half - 1)) << (logR - s)] % Q
                    u = sm[i]
                    sm[i] = (u + t) % Q
                    sm[i + half] = (u - t) % Q
            base = ((j >> logNs) << (logNs + logR)) + k
            for r in range(R):
                out[base + (r << logNs)] = sm[r]
        data = out
        logNs += logR
    if inverse:
        ninv = pow(n, -1, Q)
        data = [v * ninv % Q for v in data]
    return data


if __name__ == "__main__":
    import random
    for logn in [0, 1, 2, 3, 5, 8, 10, 11, 12, 13]:
        n = 1 << logn
        x = O.random_fr(logn + 1, n)
        assert ntt_gpu_schedule(x) == O.ntt_fast(x), logn
        assert ntt_gpu_schedule(x, inverse=True) == O.ntt_fast(x, inverse=True), logn
        print("ok", logn, plan(logn))
    # 3-pass structure exercised at small size by shrinking the plan
    def plan3(logn):
        a = (logn + 2) // 3
        b = (logn - a + 1) // 2
        return [a, b, logn - a - b]
    _plan = plan
    plan = plan3
    for logn in [3, 6, 7, 9]:
        n = 1 << logn
        x = O.random_fr(100 + logn, n)
        assert ntt_gpu_schedule(x) == O.ntt_fast(x), logn
        print("ok 3-pass", logn, plan3(logn))
