#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MSM / NTT hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--logn 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one KZG commitment: a BLS12-381 G1 MSM of 2^24 uniform Fr scalars against the synthetic
SRS P_i = [tau^i]G, tau = 101 (BASELINE.json configs[1]; at N GPUs the same 2^24 MSM sharded by index
range with an NCCL all-gather of the per-rank partial points, configs[4]).  The JSON line carries:

  value      ms per 2^24 MSM with scalars and SRS already resident in HBM (CUDA events, max over ranks)
  e2e        the same through the public API with the scalars in pinned HOST memory (H2D + D2H inside)
  roofline   the dominant kernel (msm.accumulate) against the measured IMAD peak of this GPU
  cpu_baseline / --impl reference   the reference's own algorithm (C restatement, oracle/ref_cpu.c)
             on the box's host cores, bounded sample, linear in N
  also       MSM @2^20 and Fr NTT @2^22 (the other two numbers of BASELINE.json's metric)

Nothing here reads /root/reference.  oracle/ is used only for the cpu_baseline / reference legs.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TAU = 101
METRIC = "BLS12-381 G1 MSM ms @2^24"


def ref_decomposition(logn):
    """SURVEY.md 8d fixed reference decomposition (c, W) for the algorithmic-work figure"""
    c = int(min(16, max(4, round(0.625 * logn + 1))))
    w = -(-255 // c)
    return c, w


def imad_alg_accumulate(n):
    """algorithmic lo/hi-counted IMADs of the bucket accumulation of an n-pair MSM:
    n * W * 5616 (one XYZZ += affine = 8M + 2S = 5616 IMAD)"""
    logn = max(1, int(np.ceil(np.log2(max(n, 2)))))
    c, w = ref_decomposition(logn)
    return float(n) * w * 5616.0


def gen_scalars(n, seed):
    """n uniform 254-bit Montgomery residues (all < q), numpy PCG64 -- synthetic Fr scalars"""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a


SCALAR_BLOCK = 1 << 20


def gen_scalars_range(lo, hi, seed=12345):
    """scalars [lo, hi) of the global synthetic workload: block b of 2^20 scalars comes from PCG64(seed + b),
    so every GPU count sees the same 2^24 scalars and the commitment must not depend on N"""
    out = np.empty((hi - lo, 4), dtype=np.uint64)
    b = lo // SCALAR_BLOCK
    while b * SCALAR_BLOCK < hi:
        blk = gen_scalars(SCALAR_BLOCK, seed + b)
        s = max(lo, b * SCALAR_BLOCK)
        e = min(hi, (b + 1) * SCALAR_BLOCK)
        out[s - lo:e - lo] = blk[s - b * SCALAR_BLOCK:e - b * SCALAR_BLOCK]
        b += 1
    return out


class ClockSampler:
    FIELDS = ("uuid,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid = str(uuid)
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in open(self.f.name):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8 or self.uuid.replace("GPU-", "") not in parts[0]:
                continue
            rows.append(parts)
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        out["samples"] = len(rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any(r[4 + i].lower().startswith("active") for r in rows):
                out["reasons"].append(nm)
        return out


def cpu_reference_msm(n_target, threads, sample_per_thread_log2=13, points_xyz=None, scalars=None):
    """Time the reference's algorithm (64 windows x 15 buckets, complete projective additions,
    per-window digit extraction: oracle/ref_cpu.c restating src/msm.rs) on a bounded sample and
    extrapolate linearly in N (the algorithm is exactly O(N)).  Returns (ms at n_target, info)."""
    from oracle import cref
    n_s = min(n_target, threads << sample_per_thread_log2)
    if points_xyz is None:
        points_xyz = cref.g1_iota(n_s)
    if scalars is None:
        scalars = gen_scalars(n_s, 999)
    t0 = time.perf_counter()
    cref.bucket_msm(points_xyz[:n_s], scalars[:n_s], 256, 4, threads=threads)
    dt = time.perf_counter() - t0
    ms = dt * 1e3 * (n_target / n_s)
    return ms, {"sample_pairs": int(n_s), "sample_seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = 1 << args.logn
    threads = os.cpu_count() or 1
    from oracle import cref
    n_s = min(n, threads << 15)   # ~4 s of work on every host thread per step
    pts = cref.g1_iota(n_s)
    sc = gen_scalars(n_s, 999)
    times = []
    for i in range(args.warmup + args.steps):
        ms, info = cpu_reference_msm(n, threads, points_xyz=pts, scalars=sc)
        if i >= args.warmup:
            times.append(ms)
    ms = float(np.mean(times))
    sample = ("reference algorithm (src/msm.rs bucket_msm(256,4): 64 windows x 15 buckets, complete projective adds), "
              "C restatement oracle/ref_cpu.c, %d pairs per step split over %d threads, x%.0f linear extrapolation to 2^%d"
              % (n_s, threads, n / n_s, args.logn))
    line = {
        "impl": "reference", "metric": METRIC if args.logn == 24 else "BLS12-381 G1 MSM ms @2^%d" % args.logn,
        "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": "configs[1]/[4]: standalone G1 MSM, 2^%d uniform Fr scalars, reference CPU algorithm on host cores" % args.logn},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--logn", type=int, default=24)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-prove", action="store_true", help="skip the device-resident PLONK prove line in `also`")
    ap.add_argument("--prove-logn", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-precompute", action="store_true", help="do not keep [2^(cw)]P_i levels next to the SRS")
    ap.add_argument("--pre-window", type=int, default=0)
    ap.add_argument("--fanin", type=int, default=0)
    ap.add_argument("--reduce", type=int, default=0, help="1: running-sum tree instead of the bit-plane reduction")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MSM / NTT path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg = importlib.import_module("baby-plonk-rust_b200")
    mg = importlib.import_module("baby-plonk-rust_b200.multi_gpu")
    ctx = pkg.Context(local_rank)
    if args.window:
        ctx.set_option("msm.window", args.window)
    if args.chunk:
        ctx.set_option("msm.chunk", args.chunk)
    if args.fanin:
        ctx.set_option("msm.fanin", args.fanin)
    if args.reduce:
        ctx.set_option("msm.reduce", args.reduce)
    n_total = 1 << args.logn
    com = mg.ShardedCommitter(pkg, ctx, n_total, TAU, rank, world,
                              precompute=None if args.no_precompute else args.pre_window)
    n_local = com.hi - com.lo

    # synthetic scalars: pinned host copy (for e2e) + device copy (for value)
    host = torch.empty(n_local * 4, dtype=torch.int64).pin_memory()
    host.numpy().view(np.uint64).reshape(n_local, 4)[:] = gen_scalars_range(com.lo, com.hi)
    h_scalars = host.numpy().view(np.uint64).reshape(n_local, 4)
    d_scalars = host.cuda(non_blocking=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident inputs --------------------------------------------------------
    for _ in range(args.warmup):
        com.commit_device(d_scalars)
    barrier()
    ctx.profile_reset()
    ctx.profile_enable(True)
    uuid = torch.cuda.get_device_properties(local_rank).uuid if hasattr(torch.cuda.get_device_properties(local_rank), "uuid") else ""
    sampler = ClockSampler(uuid) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = com.commit_device(d_scalars)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else {}
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = ctx.launch_count()
    plan = ctx.msm_last_plan()
    stages = {}
    for nm in ("msm.recode", "msm.sort", "msm.accumulate", "msm.merge", "msm.reduce", "msm.finalize", "g1.sum"):
        ms, cnt = ctx.profile_get(nm)
        stages[nm] = {"ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps}
    ctx.profile_enable(False)
    result_limbs = out.cpu().numpy().view(np.uint64).copy()

    # ---- e2e: host buffers through the public API ----------------------------------------------
    com.commit_host(h_scalars)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = com.commit_host(h_scalars)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    barrier()
    assert np.array_equal(r.reshape(-1), result_limbs.reshape(-1)), "e2e and device-resident results differ"

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    chain_rate, _ = ctx.imad_peak(1)   # dense IMAD.WIDE.U32.X carry chains (what the multiplier issues)
    fused_rate, _ = ctx.imad_peak(2)   # IMAD.WIDE.U32 with a 64-bit addend
    acc_ms = stages["msm.accumulate"]["ms_per_step"]
    alg = imad_alg_accumulate(n_local)
    achieved = alg / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else 0.0
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    nominal = 148 * 64 * sm_max * 1e6 / 1e12
    # a 32x32->64 IMAD.WIDE retires at 32 / clk / SM on sm_100 in every form (profiles/r1_imad_forms.md) and counts
    # as 2 lo/hi IMADs in SURVEY 8d's algorithmic figure, so the measured peak in those units is 2 x the probe rate
    peak = 2.0 * max(chain_rate, fused_rate) / 1e12
    roofline = {
        "bound": "imad", "kernel": "msm_accumulate_kernel", "achieved": achieved, "peak": peak, "unit": "TIMAD/s",
        "frac": achieved / peak if peak else None,
        "traffic": ncu_traffic(args, world),
        "algorithmic_imad_per_launch": alg, "kernel_ms": acc_ms,
        "peak_source": "measured on this GPU by bpk_imad_peak: register-only IMAD.WIDE.U32(.X) probe, 2 lo/hi IMADs per "
                       "wide op as in SURVEY 8d; nominal 148 SM x 64 lanes x f_max = %.2f TIMAD/s" % nominal,
        "frac_of_nominal": achieved / nominal,
        # what the kernel really executed (own window choice; precomputed levels need fewer windows than the fixed
        # reference decomposition, which is why `frac` can exceed the pipe's duty cycle)
        "executed_plan": plan,
        "executed_imad_per_launch": float(n_local) * plan["windows"] * 5616.0,
        "executed_frac": (float(n_local) * plan["windows"] * 5616.0) / (acc_ms * 1e-3) / 1e12 / peak if acc_ms > 0 and peak else None,
        "probe_chain_wide_imad_per_s": chain_rate, "probe_fused_acc_wide_imad_per_s": fused_rate,
        "hbm_algorithmic_gbs": (n_local * 16 * (8 + 96)) / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else None,
    }

    also = {}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_extras:
        also = extras(ctx, pkg, com, d_scalars, torch, args, peak)
    if not args.no_extras and not args.no_prove:
        prove = prove_extra(ctx, pkg, mg, torch, args, rank, world)   # every rank takes part (sharded commitments)
        if rank == 0:
            also["prove_s_2^%d_gates" % args.prove_logn] = prove
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n_s = min(n_total, threads << 15)   # ~4-8 s on every host thread: a bounded sample, extrapolated linearly
        pts = com.setup.powers_of_x(0, n_s)
        ms, info = cpu_reference_msm(n_total, threads, points_xyz=pts, scalars=h_scalars[:n_s])
        cpu_baseline = {"value": ms, "unit": "ms", "cores": threads, "kind": "port",
                        "sample": "reference algorithm src/msm.rs bucket_msm(256,4) restated in C (oracle/ref_cpu.c), first %d "
                                  "pairs of this workload over %d threads in %.1f s, x%.0f linear extrapolation"
                                  % (info["sample_pairs"], threads, info["sample_seconds"], n_total / info["sample_pairs"])}

    if rank == 0:
        line = {
            "metric": METRIC if args.logn == 24 else "BLS12-381 G1 MSM ms @2^%d" % args.logn,
            "value": ms_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "configs[1]: standalone G1 MSM, 2^%d uniform Fr scalars x synthetic SRS [tau^i]G (tau=101)"
                                   % args.logn + (", sharded over %d GPUs by index range + NCCL all-gather of partials (configs[4])" % world if world > 1 else ""),
                       "pairs": n_total, "pairs_per_gpu": n_local, "l2": "inputs larger than L2 (scalars %d MiB + SRS %d MiB per GPU)"
                       % (n_local * 32 >> 20, n_local * 96 >> 20), "parallelism": "index-range shards x%d" % world,
                       "srs": "resident in HBM" + ("" if args.no_precompute else " with precomputed window levels (bpk_srs_precompute)")},
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(n_local * 32), "d2h_bytes_per_step": 144},
            "gpu_launches": int(launches),
            "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "power_w_max", "samples")},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "stages_ms": {k: round(v["ms_per_step"], 4) for k, v in stages.items()},
            "also": also,
            "result_compressed": pkg.point_to_compressed(result_limbs).hex(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def extras(ctx, pkg, com, d_scalars, torch, args, peak_timad):
    """the other numbers of BASELINE.json's metric, same timing hygiene (device-resident, L2 flushed)"""
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts)), float(np.min(ts))

    lib, h = ctx.lib, ctx.handle
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    for logn in (16, 20):
        n = 1 << logn
        if n > com.hi - com.lo:
            continue
        # its own SRS of exactly n points (window levels sized for n), like a Setup of that size
        setup = pkg.Setup.generate_srs(n, TAU, ctx)
        if not args.no_precompute:
            setup.precompute(0)
        mean, best = timed(lambda: ctx.check(lib.bpk_msm_g1_dev(h, setup.handle, 0, d_scalars.data_ptr(), n, 1,
                                                                d_out.data_ptr())))
        stages = {}
        ctx.profile_reset()
        ctx.profile_enable(True)
        ctx.check(lib.bpk_msm_g1_dev(h, setup.handle, 0, d_scalars.data_ptr(), n, 1, d_out.data_ptr()))
        torch.cuda.synchronize()
        for nm in ("msm.recode", "msm.sort", "msm.accumulate", "msm.merge", "msm.reduce", "msm.finalize"):
            stages[nm] = round(ctx.profile_get(nm)[0], 4)
        ctx.profile_enable(False)
        out["msm_ms_2^%d" % logn] = {"mean": mean, "min": best, "plan": ctx.msm_last_plan(), "stages_ms": stages}
        setup.free()
    for logn, batch in ((22, 1), (20, 3), (24, 1)):
        n = 1 << logn
        x = torch.from_numpy(gen_scalars(n * batch, 2022).view(np.int64).reshape(-1)).cuda()
        y = torch.empty_like(x)
        ctx.profile_reset()
        mean, best = timed(lambda: ctx.check(lib.bpk_ntt_fr_dev(h, x.data_ptr(), y.data_ptr(), n, batch, 0, None)))
        key = "ntt_ms_2^%d%s" % (logn, "" if batch == 1 else "_x%d" % batch)
        out[key] = {"mean": mean, "min": best, "hbm_gbs_algorithmic": 64.0 * n * batch / (best * 1e-3) / 1e9,
                    "hbm_frac_of_measured": 64.0 * n * batch / (best * 1e-3) / 1e9 / measured_hbm_gbs(),
                    "imad_alg": (n / 2) * logn * 264 * batch,
                    # the binding roof of the NTT is the multiplier, not HBM
                    "imad_frac_of_measured": ((n / 2) * logn * 264 * batch) / (best * 1e-3) / 1e12 / peak_timad if peak_timad else None}
        del x, y
    return out


def prove_extra(ctx, pkg, mg, torch, args, rank, world):
    """BASELINE.json configs[3]: PLONK prove at 2^20 gates (SURVEY 8d C4 circuit family), device-resident
    prover with the nine commitments sharded over the ranks.  Wall clock per proof, witness columns starting
    in pinned host memory (H2D inside), proof bytes back on the host.  `prove_s` recomputes the circuit's
    pre-processed polynomials in every proof as the reference does; `prove_cached_s` keeps them in HBM."""
    import torch.distributed as dist

    prover_mod = importlib.import_module("baby-plonk-rust_b200.prover")
    synthetic = importlib.import_module("baby-plonk-rust_b200.synthetic")
    n = 1 << args.prove_logn
    t0 = time.time()
    circ = synthetic.chain_circuit(n, n - 3, seed=2)
    t_circ = time.time() - t0
    com = mg.ShardedCommitter(pkg, ctx, n + 8, TAU, rank, world, precompute=None if args.no_precompute else 0)
    blinding = [int.from_bytes(gen_scalars(11, 42)[i].tobytes(), "little") for i in range(11)]
    # the witness columns wait in pinned host memory, like the scalars of the e2e MSM figure
    wires = torch.from_numpy(np.stack(circ["wires"]).view(np.int64)).pin_memory()
    res = {"gates": n, "circuit": "chain: out public; c_k <== c_{k-1} * y_k / c_{k-1} + y_k alternating",
           "circuit_build_s": round(t_circ, 2), "n_gpus": world}
    sha = None
    for key, cache in (("prove_s", False), ("prove_cached_s", True)):
        prover = prover_mod.DeviceProver(com.setup, n, circ["selectors"], circ["sigmas"], cache_preprocessed=cache,
                                         committer=com)
        ts = []
        for it in range(4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.time()
            proof = prover.prove(wires, circ["public_inputs"], blinding)
            ts.append(time.time() - t0)
        if sha is None:
            sha = proof.sha256()
        assert proof.sha256() == sha
        t = torch.tensor(ts[1:], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[key] = {"mean": float(t.mean()), "min": float(t.min())}
        del prover
    res["proof_sha256"] = sha
    ctx.profile_reset()
    com.setup.free()
    return res


def ncu_traffic(args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full
    capture (profiles/ncu_traffic.json); only valid for the configuration that capture was taken on"""
    if world != 1 or args.logn != 24 or args.no_precompute or args.pre_window or args.window or args.chunk:
        return None
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_hbm_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    sys.exit(main())
