"""Sweep NTT plan parameters (tile size, max radix, threads) at n = 2^22 / 2^24 on the GPU box."""
import importlib, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.Context(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(1)
res = []
for logn in (22, 24, 20):
    n = 1 << logn
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1)
    x = torch.from_numpy(a.view(np.int64).reshape(-1)).cuda(); y = torch.empty_like(x)
    ref = None
    for tile, maxr, thr in ((11, 10, 0), (11, 10, 512), (11, 8, 0), (12, 11, 1024), (12, 11, 512), (12, 12, 1024), (12, 10, 1024), (10, 10, 256), (10, 8, 256), (12, 8, 1024)):
        ctx.set_option("ntt.tile_log2", tile); ctx.set_option("ntt.max_radix_log2", maxr); ctx.set_option("ntt.threads", thr)
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, x.data_ptr(), y.data_ptr(), n, 1, 0, None))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out = y.cpu().numpy()
        if ref is None: ref = out
        ok = bool(np.array_equal(ref, out))
        print("logn=%d tile=%d maxR=%d threads=%d: min %.3f ms  same=%s" % (logn, tile, maxr, thr, min(ts[2:]), ok), flush=True)
        res.append({"logn": logn, "tile": tile, "maxr": maxr, "threads": thr, "ms": min(ts[2:]), "same": ok})
json.dump(res, open("gpurun_out/ntt_sweep.json", "w"))
