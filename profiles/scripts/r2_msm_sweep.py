"""Round 2: sweep of the MSM plan tunables on one GPU (run under gpurun).
   python profiles/scripts/r2_msm_sweep.py <out.json> [logn ...]
For every size: precomputed window c (around the library default) x msm.min_pairs (which tree levels stay affine);
device-resident MSM, best of 5 after 2 warm-ups, L2 flushed between runs."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.Context(0)
lib, h = ctx.lib, ctx.handle
out_path = sys.argv[1]
logns = [int(x) for x in sys.argv[2:]] or [16, 18, 20, 22]
rng = np.random.default_rng(5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = []
WINDOWS = {16: [8, 10, 12, 13, 14], 18: [12, 14, 16, 17, 18], 20: [16, 18, 19, 20], 21: [19, 20, 21], 22: [18, 19, 20, 21],
           24: [21, 22]}
if os.environ.get("SWEEP_WINDOWS"):      # e.g. SWEEP_WINDOWS="16:12,13;20:19,20"
    WINDOWS = {int(k): [int(x) for x in v.split(",")] for k, v in (kv.split(":") for kv in os.environ["SWEEP_WINDOWS"].split(";"))}
MIN_PAIRS = [1 << int(x) for x in os.environ.get("SWEEP_MIN_PAIRS", "16,18,20,21,22,23,30").split(",")]
for logn in logns:
    n = 1 << logn
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    ref = None
    for c in WINDOWS.get(logn, [0]):
        setup = pkg.Setup.generate_srs(n, 101, ctx).precompute(c)
        for mp in MIN_PAIRS:
            ctx.set_option("msm.min_pairs", mp)
            ts = []
            for it in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.check(lib.bpk_msm_g1_dev(h, setup.handle, 0, d.data_ptr(), n, 1, d_out.data_ptr()))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            got = d_out.cpu().numpy().tobytes()
            ref = ref or got
            assert got == ref, "result changed with the plan"
            st = ctx.msm_last_stats()
            res.append({"logn": logn, "c": c, "min_pairs": mp, "ms": min(ts[2:]), "levels": st["tree_levels"]})
            print(res[-1], flush=True)
        setup.free()
    ctx.set_option("msm.min_pairs", 1 << 18)
json.dump(res, open(out_path, "w"), indent=1)
