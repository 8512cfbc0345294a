"""Time the device-resident prover on the synthetic chain circuit: python profiles/scripts/prove_time.py LOGN [REPS]"""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
bpk = importlib.import_module("baby-plonk-rust_b200")
prover_mod = importlib.import_module("baby-plonk-rust_b200.prover")
synthetic = importlib.import_module("baby-plonk-rust_b200.synthetic")

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << logn
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
cache = os.environ.get("PROVE_CACHE", "0") == "1"
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mg = importlib.import_module("baby-plonk-rust_b200.multi_gpu")
ctx = bpk.Context(local)
t0 = time.time()
circ = synthetic.chain_circuit(n, n - 3, seed=2)
t_circ = time.time() - t0
t0 = time.time()
com = mg.ShardedCommitter(bpk, ctx, n + 8, 101, rank, world, precompute=0)
setup = com.setup
ctx.synchronize()
t_setup = time.time() - t0
prover = prover_mod.DeviceProver(setup, n, circ["selectors"], circ["sigmas"], cache_preprocessed=cache, committer=com)
blinding = list(range(11, 22))
import numpy as np
import torch
wires = torch.from_numpy(np.stack(circ["wires"]).view(np.int64))
if os.environ.get("PROVE_PINNED", "1") == "1":
    wires = wires.pin_memory()
times = []
for r in range(reps + 1):
    if r == reps:
        ctx.profile_enable(True)
        ctx.profile_reset()
    t0 = time.time()
    proof = prover.prove(wires, circ["public_inputs"], blinding)
    times.append(time.time() - t0)
stages = {}
for name in ("msm.recode", "msm.sort", "msm.accumulate", "msm.merge", "msm.reduce", "msm.finalize", "ntt.pass", "ntt.coset_table",
             "plonk.grand_product", "plonk.quotient", "fr.vec_op", "fr.eval", "fr.div_linear"):
    try:
        ms, cnt = ctx.profile_get(name)
        if cnt:
            stages[name] = [round(ms, 3), cnt]
    except Exception:
        pass
if rank == 0:
  print(json.dumps({"logn": logn, "world": world, "cache": cache, "circuit_s": round(t_circ, 2), "srs_setup_s": round(t_setup, 2),
                  "prove_s": [round(t, 4) for t in times], "proof_sha256": proof.sha256(), "rounds_s": {k: round(v, 4) for k, v in prover.last_round_seconds.items()}, "stages_ms": stages,
                  "plan": ctx.msm_last_plan()}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
