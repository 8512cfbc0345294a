"""One forward Fr NTT of 2^logn (device-resident), a few repetitions: the command the NTT ncu captures run.
   python profiles/scripts/ntt_only.py [logn]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.Context(0)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << logn
rng = np.random.default_rng(2022)
a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
a[:, 3] &= np.uint64((1 << 62) - 1)
x = torch.from_numpy(a.view(np.int64)).cuda()
y = torch.empty_like(x)
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, x.data_ptr(), y.data_ptr(), n, 1, 0, None))
    e1.record()
    torch.cuda.synchronize()
    print("ntt 2^%d: %.4f ms" % (logn, e0.elapsed_time(e1)))
