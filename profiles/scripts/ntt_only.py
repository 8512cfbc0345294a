"""One forward NTT of 2^22 (3 passes) repeated a few times: the command profiled under ncu."""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.Context(0)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << logn
rng = np.random.default_rng(1)
a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1)
x = torch.from_numpy(a.view(np.int64).reshape(-1)).cuda(); y = torch.empty_like(x)
for it in range(4):
    ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, x.data_ptr(), y.data_ptr(), n, 1, 0, None))
torch.cuda.synchronize()
print("ok")
