"""Stage times of bpk_msm_g1_points (the reference's signature: points and scalars from pageable host memory on every call,
no precomputed levels): python profiles/scripts/points_call_time.py LOGN"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
bpk = importlib.import_module("baby-plonk-rust_b200")
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << logn
ctx = bpk.Context(0)
setup = bpk.Setup.generate_srs(n, 101, ctx)
pts = np.array(setup.powers_of_x(0, n), copy=True)
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
sc[:, 3] &= np.uint64((1 << 62) - 1)
out = np.zeros(18, dtype=np.uint64)
for rep in range(3):
    ctx.profile_reset(); ctx.profile_enable(True)
    t0 = time.perf_counter()
    ctx.check(ctx.lib.bpk_msm_g1_points(ctx.handle, pts.ctypes.data, n, sc.ctypes.data, n, out.ctypes.data), "points")
    dt = time.perf_counter() - t0
    ctx.profile_enable(False)
    st = {k: round(ctx.profile_get(k)[0], 3) for k in ("msm.recode", "msm.sort", "msm.accumulate", "msm.merge", "msm.reduce", "msm.finalize")}
    print("2^%d: %.1f ms" % (logn, dt * 1e3), st, ctx.msm_last_stats(), flush=True)
