"""Time the precomputed-level MSM for every window size c at n = 2^16 / 2^18 / 2^20 (calibrates the auto choice)."""
import importlib, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.Context(0)
rng = np.random.default_rng(5)
res = {}
for logn in (int(a) for a in (sys.argv[1:] or ["16", "18", "20"])):
    n = 1 << logn
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1)
    d_sc = torch.from_numpy(a.view(np.int64).reshape(-1)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    ref = None
    for c in range(7, 23):
        if ((256 + c - 1) // c) * n * 96 > 40e9:
            continue
        s = pkg.Setup.generate_srs(n, 101, ctx).precompute(c)
        ts = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.check(ctx.lib.bpk_msm_g1_dev(ctx.handle, s.handle, 0, d_sc.data_ptr(), n, 1, d_out.data_ptr()))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out = d_out.cpu().numpy().copy()
        if ref is None: ref = out
        assert np.array_equal(ref, out)
        print("logn=%d c=%d W=%d: %.3f ms" % (logn, c, (256 + c - 1) // c, min(ts[2:])), flush=True)
        res["%d/%d" % (logn, c)] = min(ts[2:])
        s.free()
json.dump(res, open("gpurun_out/msm_window_sweep.json", "w"))
