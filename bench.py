#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MSM / NTT hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--logn 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one KZG commitment: a BLS12-381 G1 MSM of 2^24 uniform Fr scalars against the synthetic
SRS P_i = [tau^i]G, tau = 101 (BASELINE.json configs[1]; at N GPUs the same 2^24 MSM sharded by index
range with an NCCL all-gather of the per-rank partial points, configs[4]).  Scalars are SURVEY 8d's:
SplitMix64(seed 12345), 8 u64 per scalar -> 512-bit little-endian -> mod q.  The JSON line carries:

  value      ms per 2^24 MSM with scalars and SRS already resident in HBM (CUDA events, max over ranks)
  e2e        the same through the public API with the scalars in pinned HOST memory (H2D + D2H inside);
             e2e_variants adds pageable host scalars and the reference-signature call (points on every call)
  verified   the timed result equals the closed form [sum_i s_i tau^i]G computed on the CPU (oracle Horner + one
             scalar multiplication), outside the timed region
  roofline   the dominant kernel against the measured IMAD peak of this GPU
  cpu_baseline / --impl reference   the reference's own algorithm (C restatement, oracle/ref_cpu.c) on the box's
             host cores: ONE sampling rule for both legs (cpu_reference_sample)
  also       MSM @2^16 / 2^20, Fr NTT @2^22 (+ batch, 2^24), PLONK prove @2^20 gates (proof_verified), each with
             its CPU baselines (reference algorithm, flagged extrapolations; optimised all-core CPU code)

Nothing here reads /root/reference.  oracle/ is used only as the checker of results (outside every timed region)
and for the cpu_baseline / reference legs.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TAU = 101
SCALAR_SEED = 12345
METRIC = "BLS12-381 G1 MSM ms @2^24"
Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FR_R = (1 << 256) % Q
# lo/hi-counted IMADs: Fp product 588, Fp square 456
IMAD_MADD_XYZZ = 8 * 588 + 2 * 456          # XYZZ += affine (madd-2008-s): 5616
IMAD_ADD_AFFINE = 5 * 588 + 1 * 456         # affine + affine with a shared inversion (Montgomery's trick): 3396


def ref_decomposition(logn):
    """SURVEY.md 8d fixed reference decomposition (c, W) for the algorithmic-work figure"""
    c = int(min(16, max(4, round(0.625 * logn + 1))))
    w = -(-255 // c)
    return c, w


def imad_alg_accumulate(n):
    """algorithmic lo/hi-counted IMADs of the bucket accumulation of an n-pair MSM:
    n * W * 5616 (one XYZZ += affine = 8M + 2S = 5616 IMAD) with the fixed (c, W) of ref_decomposition"""
    logn = max(1, int(np.ceil(np.log2(max(n, 2)))))
    c, w = ref_decomposition(logn)
    return float(n) * w * IMAD_MADD_XYZZ


# ---------------------------------------------------------------------------------------------------------
# synthetic scalars: SplitMix64 -> 64 bytes per scalar -> from_bytes_wide (scalar.rs:654-658)
# ---------------------------------------------------------------------------------------------------------
def splitmix64_words(first, count, seed=SCALAR_SEED):
    """outputs [first, first + count) of SplitMix64(seed), vectorised (uint64 arithmetic wraps)"""
    with np.errstate(over="ignore"):
        k = np.arange(first + 1, first + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def scalars_host_ints(lo, hi, seed=SCALAR_SEED):
    """scalars [lo, hi) of the workload as canonical Python ints (CPU-only legs; ~1 us per scalar)"""
    raw = splitmix64_words(8 * lo, 8 * (hi - lo), seed).tobytes()
    return [int.from_bytes(raw[64 * i:64 * i + 64], "little") % Q for i in range(hi - lo)]


def scalars_host_mont(lo, hi, seed=SCALAR_SEED):
    """the same as uint64[n, 4] Montgomery limbs, computed on the CPU"""
    raw = b"".join((v * FR_R % Q).to_bytes(32, "little") for v in scalars_host_ints(lo, hi, seed))
    return np.frombuffer(raw, dtype=np.uint64).reshape(-1, 4).copy()


def scalars_device_mont(ctx, pkg, torch, lo, hi, seed=SCALAR_SEED):
    """scalars [lo, hi) as an int64[n, 4] device tensor of Montgomery limbs.  The 512-bit value v = a0 + 2^248 a1 +
    2^496 a2 (a0, a1 < 2^248, a2 < 2^16) is reduced on the device with the library's own Fr vector ops:
    v R = a0 * R + a1 * (2^248 R) + a2 * (2^496 R) as three Montgomery products by constants (workload setup,
    not timed; checked against the CPU generator in tests/test_gpu_parity.py)."""
    n = hi - lo
    out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    consts = pkg.scalars_from_ints([FR_R, (1 << 248) * FR_R % Q, (1 << 496) * FR_R % Q])
    lib, h = ctx.lib, ctx.handle
    block = 1 << 20
    for s in range(0, n, block):
        m = min(block, n - s)
        by = splitmix64_words(8 * (lo + s), 8 * m, seed).view(np.uint8).reshape(m, 64)
        parts = np.zeros((3, m, 32), dtype=np.uint8)
        parts[0, :, :31] = by[:, 0:31]
        parts[1, :, :31] = by[:, 31:62]
        parts[2, :, :2] = by[:, 62:64]
        d = torch.from_numpy(parts.view(np.int64).reshape(3, m, 4)).cuda()
        o = out[s:s + m]
        ctx.check(lib.bpk_fr_vec_op(h, 3, d[0].data_ptr(), None, consts[0].ctypes.data, o.data_ptr(), m))
        ctx.check(lib.bpk_fr_vec_op(h, 4, o.data_ptr(), d[1].data_ptr(), consts[1].ctypes.data, o.data_ptr(), m))
        ctx.check(lib.bpk_fr_vec_op(h, 4, o.data_ptr(), d[2].data_ptr(), consts[2].ctypes.data, o.data_ptr(), m))
        torch.cuda.synchronize()
    return out


def gen_scalars(n, seed):
    """n uniform 254-bit Montgomery residues (all < q), numpy PCG64 -- NTT inputs and blinding of the extras"""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a


class ClockSampler:
    """SM clock, throttle reasons and power of one GPU DURING the timed region.  NVML in a thread of this process (a sample
    every 10 ms: the timed region of the default run is ~0.3 s), `nvidia-smi -lms 100` next to it as the fallback (its
    start-up alone can outlast the region)."""
    FIELDS = ("uuid,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, uuid):
        self.uuid = str(uuid)
        self.rows = []           # NVML samples: (sm_mhz, reasons bitmask, watts)
        self.sm_max = None
        self.halt = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            name = self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(name.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByUUID(name)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")

            def loop():
                while not self.halt.is_set():
                    try:
                        self.rows.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(reasons_fn(h)),
                                          pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
                    except Exception:
                        pass
                    self.halt.wait(0.01)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.halt.set()
            self.thread.join(timeout=2)
        smi = self._stop_smi()
        if self.rows:
            sm = sorted(r[0] for r in self.rows)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.sm_max, power_w_max=max(r[2] for r in self.rows),
                       samples=len(self.rows), source="nvml, 10 ms")
            seen = 0
            for r in self.rows:
                seen |= r[1]
            out["reasons"] = [nm for nm, bit in self.NVML_REASONS if seen & bit]
            return out
        if smi:
            out.update(smi)
        return out

    def _stop_smi(self):
        if self.p is None:
            return None
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
            return None
        out = {"source": "nvidia-smi -lms 100"}
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        out["samples"] = len(rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [nm for i, nm in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU legs (oracle/ as the thing TIMED: only here and in run_reference)
# ---------------------------------------------------------------------------------------------------------
SAMPLE_PAIRS_PER_THREAD_LOG2 = 16   # ~7-15 s of the reference algorithm on every host thread


def _par(fn, jobs, threads):
    """run fn(job) on `threads` Python threads (the C calls release the GIL)"""
    out = [None] * len(jobs)

    def work(t):
        for i in range(t, len(jobs), threads):
            out[i] = fn(jobs[i])

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out


def srs_points_cpu(n, threads):
    """[tau^i]G, i < n, as un-normalised G1Projective limbs, built on the host cores (workload setup for the CPU legs)"""
    from oracle import bls12_381 as O
    from oracle import cref
    chunks = max(1, min(threads, n // 64 or 1))
    bounds = [n * k // chunks for k in range(chunks + 1)]
    starts = [np.array(O.g1_scale_proj(O.g1_mul(O.G1_GEN, pow(TAU, bounds[k], Q)), 1), dtype=np.uint64)
              for k in range(chunks)]
    parts = _par(lambda k: cref.g1_powers_small(starts[k], TAU, bounds[k + 1] - bounds[k]), list(range(chunks)), threads)
    return np.concatenate(parts)


def cpu_reference_sample(n_target, threads):
    """THE sampling rule of both CPU legs (cpu_baseline of the GPU arm and --impl reference): the reference's
    algorithm (src/msm.rs bucket_msm(256, 4): 64 windows x 15 buckets, complete projective additions, per-window digit
    extraction; C restatement oracle/ref_cpu.c) on the FIRST threads * 2^15 pairs of the benchmark workload
    (same SRS points [tau^i]G, same SplitMix64 scalars), split over all host threads, extrapolated linearly in N
    (the algorithm is exactly O(N)); the sample's result is checked against the closed form."""
    from oracle import bls12_381 as O
    from oracle import cref
    n_s = min(n_target, threads << SAMPLE_PAIRS_PER_THREAD_LOG2)
    pts = srs_points_cpu(n_s, threads)
    ints = scalars_host_ints(0, n_s)
    sc = np.frombuffer(b"".join((v * FR_R % Q).to_bytes(32, "little") for v in ints), dtype=np.uint64).reshape(-1, 4)

    def run():
        t0 = time.perf_counter()
        out = cref.bucket_msm(pts, sc, 256, 4, threads=threads)
        return time.perf_counter() - t0, out

    def check(out):
        e = 0
        for v in reversed(ints):
            e = (e * TAU + v) % Q
        return O.g1_proj_limbs_to_affine([int(x) for x in out]) == O.g1_mul(O.G1_GEN, e)

    sample = ("reference algorithm (src/msm.rs bucket_msm(256,4): 64 windows x 15 buckets, complete projective adds), "
              "C restatement oracle/ref_cpu.c, first %d pairs of this workload split over %d threads, x%.0f linear "
              "extrapolation to %d pairs" % (n_s, threads, n_target / n_s, n_target))
    return run, check, n_s, sample


def workload_name(logn, world):
    """config.workload, shared by both arms so that the driver compares like with like"""
    return ("configs[1]: standalone G1 MSM, 2^%d uniform Fr scalars (SplitMix64 seed %d, from_bytes_wide) x "
            "synthetic SRS [tau^i]G (tau=%d)" % (logn, SCALAR_SEED, TAU)
            + (", sharded over %d GPUs by index range + NCCL all-gather of partials (configs[4])" % world if world > 1 else ""))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = 1 << args.logn
    threads = os.cpu_count() or 1
    run, check, n_s, sample = cpu_reference_sample(n, threads)
    times, ok = [], True
    for i in range(args.warmup + args.steps):
        dt, out = run()
        if i >= args.warmup:
            times.append(dt * 1e3 * (n / n_s))
        if i == 0:
            ok = check(out)
    ms = float(np.mean(times))
    line = {
        "impl": "reference", "metric": METRIC if args.logn == 24 else "BLS12-381 G1 MSM ms @2^%d" % args.logn,
        "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.logn, args.gpus), "pairs": n,
                   "arm": "reference CPU algorithm on the host cores (the GPUs are not used)"},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": threads, "kind": "port", "sample": sample,
                         "sample_seconds_per_step": float(np.mean(times)) * n_s / n / 1e3, "sample_verified": bool(ok)},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_ntt_baselines(logn, threads):
    """CPU baselines of one forward Fr NTT of 2^logn: (a) the reference's naive DFT (utils.rs:63-81: n^2 terms, one
    256-bit pow each) timed on a bounded sample of output rows at n = 2^12 on all threads and extrapolated with n^2
    -- FLAGGED: the reference cannot run this size (8 n^2 bytes of scratch); (b) optimised radix-2 on all cores"""
    from oracle import cref
    n_s = 1 << 12
    x = gen_scalars(n_s, 7)
    rows_per_thread = 4
    t0 = time.perf_counter()
    _par(lambda t: cref.ntt_381_rows(x, t * rows_per_thread, (t + 1) * rows_per_thread), list(range(threads)), threads)
    dt = time.perf_counter() - t0
    per_term = dt / (rows_per_thread * n_s)            # seconds per DFT term with `threads` rows in flight
    n = 1 << logn
    ref_ms = per_term * float(n) * float(n) * 1e3
    y = gen_scalars(n, 2022)
    cref.ntt_fast(y[:1 << 12], threads=threads)
    t0 = time.perf_counter()
    cref.ntt_fast(y, threads=threads)
    opt_ms = (time.perf_counter() - t0) * 1e3
    return {
        "reference_algorithm": {"value": ref_ms, "unit": "ms", "cores": threads, "kind": "port", "extrapolated": True,
                                "sample": "naive DFT of src/utils.rs:63-81 (oracle/ref_cpu.c): %d output rows of a 2^12 transform on "
                                          "%d threads in %.1f s, scaled by n^2 to 2^%d (the reference itself cannot allocate its n x n "
                                          "matrix at this size)" % (threads * rows_per_thread, threads, dt, logn)},
        "optimised_cpu": {"value": opt_ms, "unit": "ms", "cores": threads, "kind": "optimised radix-2 NTT (oracle/fast_cpu.c), "
                          "measured directly at 2^%d" % logn},
    }


def cpu_msm_optimised(points_xyz, scalars_mont, threads, n_target):
    """optimised all-core CPU MSM (signed-digit Pippenger, XYZZ buckets; oracle/fast_cpu.c) on the given pairs,
    linear extrapolation to n_target"""
    from oracle import cref
    n_s = points_xyz.shape[0]
    t0 = time.perf_counter()
    out = cref.msm_pippenger(points_xyz, scalars_mont[:n_s], threads=threads)
    dt = time.perf_counter() - t0
    return {"value": dt * 1e3 * n_target / n_s, "unit": "ms", "cores": threads,
            "kind": "optimised signed-digit Pippenger on all cores (oracle/fast_cpu.c)",
            "sample": "%d pairs in %.2f s, x%.0f linear extrapolation" % (n_s, dt, n_target / n_s)}, out


def cpu_prove_baseline():
    """the reference-algorithm prover (oracle/plonk.py rounds on CRefBackend: naive i_ntt_381, Mul by coeffs_evaluate +
    naive inverse DFT, bucket_msm(256,4) with complete additions -- the three surfaces in C, the O(n) glue in Python),
    single thread as the reference, at the sizes it can reach; n^2 growth, so 2^20 gates is out of reach by ~10^9"""
    from oracle import bls12_381 as O
    from oracle import plonk as P
    out = {"kind": "port", "cores": 1, "unit": "s", "sizes": {}}
    for n in (8, 16, 32):
        if n == 8:
            prog, wit, pub = P.reference_test_circuit()
        else:
            prog, wit, pub = P.synthetic_circuit(n, n - 3, seed=4)
        srs = O.generate_srs_points(n + 6, TAU)
        t0 = time.perf_counter()
        proof = P.prove(prog, wit, list(range(1, 12)), P.CRefBackend(srs))
        out["sizes"]["n=%d" % n] = round(time.perf_counter() - t0, 3)
        if n == 8:
            out["n8_proof_sha256"] = proof.sha256()
    t32 = out["sizes"]["n=32"]
    out["extrapolated_2^20_gates_s"] = t32 * (float(1 << 20) / 32) ** 2
    out["sample"] = ("Prover::prove (src/prover.rs:106-176) by the reference's own algorithms at n = 8 (tests/verify_proof_test.rs), "
                     "16, 32; extrapolated_2^20_gates_s scales n = 32 by n^2 (FLAGGED: transforms and products are O(n^2) "
                     "in the reference, SURVEY 6)")
    return out


# ---------------------------------------------------------------------------------------------------------
def closed_form_commitment(scalars_mont_u64, first=0):
    """[sum_i s_i tau^(first + i)]G as an affine point: oracle C Horner + one scalar multiplication"""
    from oracle import bls12_381 as O
    from oracle import cref
    tau_m = np.array(O.fr_to_mont(TAU), dtype=np.uint64)
    e = O.fr_from_mont([int(x) for x in cref.fr_horner(scalars_mont_u64, tau_m)])
    return O.g1_mul(O.G1_GEN, e * pow(TAU, first, Q) % Q)


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
    ap.add_argument("--no-verify", action="store_true", help="skip the CPU closed-form / verifier checks of the results")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-precompute", action="store_true", help="do not keep [2^(cw)]P_i levels next to the SRS")
    ap.add_argument("--pre-window", type=int, default=0)
    ap.add_argument("--set", action="append", default=[], help="library tunable key=value (bpk_set_option)")
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
    for kv in args.set:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    n_total = 1 << args.logn
    t0 = time.perf_counter()
    com = mg.ShardedCommitter(pkg, ctx, n_total, TAU, rank, world, precompute=None, n_global=n_total)
    torch.cuda.synchronize()
    srs_generate_ms = (time.perf_counter() - t0) * 1e3
    precompute_ms = 0.0
    if not args.no_precompute:
        t0 = time.perf_counter()
        com.precompute(args.pre_window)
        torch.cuda.synchronize()
        precompute_ms = (time.perf_counter() - t0) * 1e3
    n_local = com.hi - com.lo

    # synthetic scalars: device copy (for value) + pinned / pageable host copies (for e2e)
    d_scalars = scalars_device_mont(ctx, pkg, torch, com.lo, com.hi)
    host = torch.empty(n_local * 4, dtype=torch.int64).pin_memory()
    host.copy_(d_scalars.view(-1))
    torch.cuda.synchronize()
    h_scalars = host.numpy().view(np.uint64).reshape(n_local, 4)

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
    props = torch.cuda.get_device_properties(local_rank)
    uuid = props.uuid if hasattr(props, "uuid") else ""
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
    stats = ctx.msm_last_stats()      # of the device-resident MSM just timed (the e2e legs below run two slices)
    stages = {}
    for nm in STAGES:
        ms, cnt = ctx.profile_get(nm)
        if cnt:
            stages[nm] = {"ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps}
    ctx.profile_enable(False)
    result_limbs = out.cpu().numpy().view(np.uint64).copy()

    # ---- e2e: host buffers through the public API ----------------------------------------------
    def time_host(fn, steps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        torch.cuda.synchronize()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
        barrier()
        return ms, r

    e2e_ms, r = time_host(lambda: com.commit_host(h_scalars), args.steps)
    assert np.array_equal(r.reshape(-1), result_limbs.reshape(-1)), "e2e and device-resident results differ"
    e2e_variants = {}
    if not args.no_extras:
        pageable = np.array(h_scalars, copy=True)       # an ordinary heap allocation, like a Rust Vec<Scalar>
        ms, r = time_host(lambda: com.commit_host(pageable), max(1, min(args.steps, 3)))
        assert np.array_equal(r.reshape(-1), result_limbs.reshape(-1))
        e2e_variants["pageable_host_scalars"] = {"value": ms, "unit": "ms", "h2d_bytes_per_step": int(n_local * 32),
                                                 "note": "scalars in pageable host memory (what a Rust Vec is); SRS resident"}
        del pageable

    # ---- verified: closed form on the CPU, outside every timed region ----------------------------
    verified = None
    if rank == 0 and not args.no_verify:
        if world == 1:
            all_sc = h_scalars
        else:
            all_sc = scalars_device_mont(ctx, pkg, torch, 0, n_total).cpu().numpy().view(np.uint64)
        verified = bool(closed_form_commitment(all_sc) == pkg.point_to_affine(result_limbs))
        del all_sc

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    chain_rate, _ = ctx.imad_peak(1)   # dense IMAD.WIDE.U32.X carry chains (what the multiplier issues)
    fused_rate, _ = ctx.imad_peak(2)   # IMAD.WIDE.U32 with a 64-bit addend
    acc_ms = stages.get("msm.accumulate", {}).get("ms_per_step", 0.0)
    alg = imad_alg_accumulate(n_local)
    achieved = alg / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else 0.0
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    nominal = 148 * 64 * sm_max * 1e6 / 1e12
    # a 32x32->64 IMAD.WIDE retires at 32 / clk / SM on sm_100 in every form (profiles/r1_imad_forms.md) and counts
    # as 2 lo/hi IMADs in SURVEY 8d's algorithmic figure, so the measured peak in those units is 2 x the probe rate
    peak = 2.0 * max(chain_rate, fused_rate) / 1e12
    executed = stats["affine_adds"] * IMAD_ADD_AFFINE + stats["xyzz_adds_bound"] * IMAD_MADD_XYZZ
    traffic, traffic_source = ncu_traffic(args, world)
    roofline = {
        "bound": "imad", "kernel": DOMINANT_KERNEL, "achieved": achieved, "peak": peak, "unit": "TIMAD/s",
        "frac": achieved / peak if peak else None,
        "traffic": traffic, "traffic_source": traffic_source,
        "algorithmic_imad_per_launch": alg, "kernel_ms": acc_ms,
        "kernel_ms_note": "sum over the launches of the accumulate stage of one MSM (one launch per level of the pairwise "
                          "batched-affine tree + the XYZZ tail), CUDA events on the launching stream",
        "peak_source": "measured on this GPU by bpk_imad_peak: register-only IMAD.WIDE.U32(.X) probe, 2 lo/hi IMADs per "
                       "wide op as in SURVEY 8d; nominal 148 SM x 64 lanes x f_max = %.2f TIMAD/s" % nominal,
        "frac_of_nominal": achieved / nominal,
        # what the kernel really executed (own window choice, precomputed levels, batched-affine additions at
        # 5M + 1S instead of 8M + 2S -- which is why `frac` can exceed the pipe's duty cycle)
        "executed_plan": plan, "executed_stats": stats,
        "executed_imad_per_launch": float(executed),
        "executed_frac": executed / (acc_ms * 1e-3) / 1e12 / peak if acc_ms > 0 and peak else None,
        "probe_chain_wide_imad_per_s": chain_rate, "probe_fused_acc_wide_imad_per_s": fused_rate,
    }
    if rank == 0:
        write_imad_peak(chain_rate, fused_rate, clocks, props.name)

    also = {}
    cpu_baseline = None
    threads = os.cpu_count() or 1
    if rank == 0 and world == 1 and not args.no_extras:
        also = extras(ctx, pkg, com, d_scalars, h_scalars, torch, args, peak, threads)
        if args.logn <= 24:
            e2e_variants["reference_signature_points_per_call"] = points_per_call(ctx, pkg, com, h_scalars, n_local, result_limbs)
    if not args.no_extras and not args.no_prove:
        prove = prove_extra(ctx, pkg, mg, torch, args, rank, world)   # every rank takes part (sharded commitments)
        if rank == 0:
            if not args.no_cpu and world == 1:
                prove["cpu_baseline"] = cpu_prove_baseline()
            also["prove_s_2^%d_gates" % args.prove_logn] = prove
    if rank == 0 and world == 1 and not args.no_cpu:
        run, check, n_s, sample = cpu_reference_sample(n_total, threads)
        dt, ref_out = run()
        cpu_baseline = {"value": dt * 1e3 * n_total / n_s, "unit": "ms", "cores": threads, "kind": "port", "sample": sample,
                        "sample_seconds": dt, "sample_verified": bool(check(ref_out))}
        pts = com.setup.powers_of_x(0, min(n_local, 1 << 22))
        opt, opt_out = cpu_msm_optimised(pts, h_scalars, threads, n_total)
        cpu_baseline["optimised_cpu"] = opt
        del pts

    if rank == 0:
        line = {
            "metric": METRIC if args.logn == 24 else "BLS12-381 G1 MSM ms @2^%d" % args.logn,
            "value": ms_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "verified": verified,
            "config": {"workload": workload_name(args.logn, world),
                       "pairs": n_total, "pairs_per_gpu": n_local, "l2": "inputs larger than L2 (scalars %d MiB + SRS %d MiB per GPU)"
                       % (n_local * 32 >> 20, n_local * 96 >> 20), "parallelism": "index-range shards x%d" % world,
                       "srs": "resident in HBM" + ("" if args.no_precompute else " with precomputed window levels (bpk_srs_precompute)"),
                       "srs_generate_ms": round(srs_generate_ms, 1), "precompute_ms": round(precompute_ms, 1),
                       "srs_table_bytes": int(com.setup.table_bytes()),
                       "setup_note": "one-time per SRS, outside the timed region: a KZG committer keeps its SRS resident"},
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(n_local * 32), "d2h_bytes_per_step": 144,
                    "note": "scalars in pinned host memory, SRS resident"},
            "e2e_variants": e2e_variants,
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


STAGES = ("msm.recode", "msm.sort", "msm.accumulate", "msm.merge", "msm.reduce", "msm.finalize", "msm.join", "g1.sum")
DOMINANT_KERNEL = "msm_affine_level_kernel (+ msm_accumulate_kernel tail)"


def write_imad_peak(chain_rate, fused_rate, clocks, gpu_name):
    """record the measured multiplier peak next to MEASURED_PEAKS.json's numbers: written to
    gpurun_out/ on the GPU box; the copy under profiles/ is the one the repo's roofline figures cite"""
    rec = {"gpu_name": gpu_name, "how": "bpk_imad_peak (csrc/api.cu imad_probe_kernel): register-only probes, 16 CTAs x 128 "
           "threads per SM, 2^14 iterations; chain = two independent 12-limb IMAD.WIDE.U32.X carry chains (the shape the Fp "
           "multiplier issues), fused = 14 independent IMAD.WIDE.U32 with 64-bit addend",
           "wide_imad_per_s_chain": chain_rate, "wide_imad_per_s_fused": fused_rate,
           "wide_imad_per_clk_per_sm": max(chain_rate, fused_rate) / 148 / ((clocks.get("sm_mhz") or 1965.0) * 1e6),
           "timad_per_s_lo_hi_counting": 2.0 * max(chain_rate, fused_rate) / 1e12,
           "sm_mhz_during_bench": clocks.get("sm_mhz"), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        json.dump(rec, open(os.path.join(d, "imad_peak.json"), "w"), indent=1)
    except Exception:
        pass


def points_per_call(ctx, pkg, com, h_scalars, n, expect_limbs):
    """the reference's own signature, BucketMSM::bucket_msm(points, scalars, 256, 4): points AND scalars handed over
    on every call from pageable host memory (bpk_msm_g1_points: upload 144 B/point, convert to affine, no
    precomputed levels, free) -- the drop-in without a resident SRS"""
    pts = com.setup.powers_of_x(0, n)            # uint64[n, 18] normalised projective, pageable
    sc = np.array(h_scalars, copy=True)
    out = np.empty(18, dtype=np.uint64)
    lib, h = ctx.lib, ctx.handle
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.check(lib.bpk_msm_g1_points(h, pts.ctypes.data, n, sc.ctypes.data, n, out.ctypes.data), "bpk_msm_g1_points")
        times.append((time.perf_counter() - t0) * 1e3)
    ok = bool(np.array_equal(out, expect_limbs.reshape(-1)))
    return {"value": min(times), "unit": "ms", "h2d_bytes_per_step": int(n * (144 + 32)), "d2h_bytes_per_step": 144,
            "same_result": ok, "all_ms": [round(t, 1) for t in times],
            "note": "bpk_msm_g1_points: 3 calls, best; includes cudaMalloc/cudaFree of the point buffer and staging 2.9 GB "
                    "from pageable host memory, which varies 2x between boxes"}


def extras(ctx, pkg, com, d_scalars, h_scalars, torch, args, peak_timad, threads):
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
        for nm in STAGES:
            ms, cnt = ctx.profile_get(nm)
            if cnt:
                stages[nm] = round(ms, 4)
        ctx.profile_enable(False)
        rec = {"mean": mean, "min": best, "plan": ctx.msm_last_plan(), "stages_ms": stages}
        if not args.no_verify:
            rec["verified"] = bool(closed_form_commitment(h_scalars[:n]) ==
                                   pkg.point_to_affine(d_out.cpu().numpy().view(np.uint64)))
        if logn == 20 and not args.no_cpu:
            opt, _ = cpu_msm_optimised(setup.powers_of_x(0, n), h_scalars, threads, n)
            rec["cpu_optimised"] = opt
        out["msm_ms_2^%d" % logn] = rec
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
        if logn == 22 and not args.no_verify:
            # size-independent check: the transform of a delta at position 1 is the row of powers of w (spot-checked)
            # and inverse(forward(x)) == x, bit for bit
            z = torch.empty_like(x)
            ctx.check(lib.bpk_ntt_fr_dev(h, y.data_ptr(), z.data_ptr(), n, batch, 1, None))
            out[key]["round_trip_exact"] = bool(torch.equal(x, z))
            del z
        if logn == 22 and not args.no_cpu:
            out[key]["cpu_baseline"] = cpu_ntt_baselines(logn, threads)
        del x, y
    return out


def as_oracle_proof(proof):
    from oracle import bls12_381 as O
    from oracle import plonk as P
    kw = {k: O.g1_from_compressed(getattr(proof, k)) for k in proof.POINTS}
    kw.update({k: getattr(proof, k) for k in proof.SCALARS})
    return P.Proof(**kw)


def prove_extra(ctx, pkg, mg, torch, args, rank, world):
    """BASELINE.json configs[3]: PLONK prove at 2^20 gates (SURVEY 8d C4 circuit family), device-resident
    prover with its commitments and transforms dealt over the ranks.  Wall clock per proof, witness columns starting
    in pinned host memory (H2D inside), proof bytes back on the host.  `prove_s` recomputes the circuit's
    pre-processed polynomials in every proof as the reference does; `prove_cached_s` keeps them in HBM.
    proof_verified: rank 0 checks the LAST timed proof with the verifier equation (trapdoor form), the eight
    pre-processed commitments formed on the CPU in closed form (oracle/plonk.py::verify_columns)."""
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
    if rank == 0 and not args.no_verify:
        from oracle import plonk as P
        t0 = time.time()
        res["proof_verified"] = bool(P.verify_columns(n, circ["selectors"], circ["sigmas"], as_oracle_proof(proof),
                                                      circ["public_inputs"], TAU))
        res["verify_s"] = round(time.time() - t0, 2)
    ctx.profile_reset()
    com.setup.free()
    return res


def ncu_traffic(args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per MSM.  NOT measured in this run (ncu
    cannot run inside a timed bench): read from the committed ncu --set full capture, and only reported for the
    configuration that capture was taken on"""
    if world != 1 or args.logn != 24 or args.no_precompute or args.pre_window or args.window or args.chunk or args.set:
        return None, None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return float(rec["dram_bytes_per_launch"]), "profiles/ncu_traffic.json <- " + rec.get("source", "")
    except Exception:
        return None, None


def measured_hbm_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    sys.exit(main())
