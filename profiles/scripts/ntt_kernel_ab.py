"""A/B of the two NTT pass kernels (ntt.kernel 1 = one radix-2 stage per barrier, 0 = register-blocked radix-8 steps)
over tile sizes / pass plans; every configuration must give the same output."""
import importlib, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.Context(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(1)
res = []
for logn, batch in ((22, 1), (24, 1), (20, 3), (16, 1)):
    n = 1 << logn
    a = rng.integers(0, 1 << 64, size=(n * batch, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1)
    x = torch.from_numpy(a.view(np.int64).reshape(-1)).cuda(); y = torch.empty_like(x)
    ref = None
    for kern, tile, maxr in ((1, 10, 10), (0, 10, 10), (0, 10, 8), (0, 11, 11), (0, 11, 10), (0, 11, 8), (0, 10, 9), (0, 11, 9), (0, 9, 9), (0, 10, 6), (0, 11, 6), (3, 10, 8), (3, 10, 10), (3, 9, 8), (3, 9, 9), (3, 10, 6), (3, 8, 8)):
        ctx.set_option("ntt.kernel", kern); ctx.set_option("ntt.tile_log2", tile); ctx.set_option("ntt.max_radix_log2", maxr)
        ts = []
        for it in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, x.data_ptr(), y.data_ptr(), n, batch, 0, None))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out = y.cpu().numpy()
        if ref is None: ref = out
        ok = bool(np.array_equal(ref, out))
        print("logn=%d x%d kernel=%d tile=%d maxR=%d: min %.3f ms  same=%s" % (logn, batch, kern, tile, maxr, min(ts[2:]), ok), flush=True)
        res.append({"logn": logn, "batch": batch, "kernel": kern, "tile": tile, "maxr": maxr, "ms": min(ts[2:]), "same": ok})
json.dump(res, open("gpurun_out/ntt_kernel_ab.json", "w"))
