"""Round 2: sweep of the NTT plan tunables (tile size, largest radix) at 2^20 / 2^22 / 2^24 on one GPU."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.Context(0)
rng = np.random.default_rng(2022)
res = []
for logn in (20, 22, 24):
    n = 1 << logn
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    x = torch.from_numpy(a.view(np.int64)).cuda()
    y = torch.empty_like(x)
    ref = None
    for tile in (9, 10, 11):
        for maxr in (0, 6, 7, 8, 9):
            if maxr > tile:
                continue
            ctx.set_option("ntt.tile_log2", tile)
            ctx.set_option("ntt.max_radix_log2", maxr)
            ts = []
            for it in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, x.data_ptr(), y.data_ptr(), n, 1, 0, None))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            got = y.cpu().numpy().tobytes()
            ref = ref or got
            assert got == ref
            res.append({"logn": logn, "tile": tile, "max_radix": maxr, "ms": min(ts[2:])})
            print(res[-1], flush=True)
json.dump(res, open(sys.argv[1], "w"), indent=1)
