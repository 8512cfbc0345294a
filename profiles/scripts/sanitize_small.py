"""Small end-to-end exercise of every kernel family (all NTT pass kernels, both bucket reductions, lanes, prover); sizes kept tiny so it can run under a sanitizer where one is available (compute-sanitizer is closed on this pool)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
bpk = importlib.import_module("baby-plonk-rust_b200")
prover_mod = importlib.import_module("baby-plonk-rust_b200.prover")
synthetic = importlib.import_module("baby-plonk-rust_b200.synthetic")
ctx = bpk.Context(0)
rng = np.random.default_rng(3)


def scalars(n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a


# NTT: all three pass kernels, forward / inverse / coset, batch
for kernel in (1, 2, 3):
    ctx.set_option("ntt.kernel", kernel)
    for logn in (6, 11, 13):
        x = scalars(2 << logn).reshape(2, 1 << logn, 4)
        y = bpk.ntt_381(x, ctx)
        assert np.array_equal(bpk.i_ntt_381(y, ctx), x)
        assert np.array_equal(bpk.coset_intt(bpk.coset_ntt(x, 7, ctx), 7, ctx), x)
ctx.set_option("ntt.kernel", 0)
# MSM: own windows, precomputed levels, batch on lanes, host path, both reductions
n = 3000
setup = bpk.Setup.generate_srs(n, 101, ctx)
sc = scalars(n)
ref = setup.commit_scalars(sc)
ctx.set_option("msm.affine_levels", 3)
assert np.array_equal(setup.commit_scalars(sc), ref)
ctx.set_option("msm.affine_levels", -1)
setup.precompute(0)
assert np.array_equal(setup.commit_scalars(sc), ref)
import ctypes
d = [torch.from_numpy(sc.view(np.int64)).cuda() for _ in range(3)]
ptrs = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in d])
firsts = (ctypes.c_size_t * 3)(0, 0, 0)
lens = (ctypes.c_size_t * 3)(n, n, n)
out = torch.zeros((3, 18), dtype=torch.int64, device="cuda")
ctx.check(ctx.lib.bpk_msm_g1_dev_batch(ctx.handle, setup.handle, 3, ptrs, firsts, lens, 1, out.data_ptr()))
assert all(np.array_equal(out[i].cpu().numpy().view(np.uint64), ref) for i in range(3))
# device prover on a 64-row circuit
circ = synthetic.chain_circuit(64, 40, seed=4)
s2 = bpk.Setup.generate_srs(72, 101, ctx)
proof = prover_mod.DeviceProver(s2, 64, circ["selectors"], circ["sigmas"]).prove(circ["wires"], circ["public_inputs"], list(range(1, 12)))
print("sanitize_small ok", proof.sha256()[:16])
