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
ctx = bpk.Context(0)
t0 = time.time()
circ = synthetic.chain_circuit(n, n - 3, seed=2)
t_circ = time.time() - t0
t0 = time.time()
setup = bpk.Setup.generate_srs(n + 8, 101, ctx)
setup.precompute(0)
ctx.synchronize()
t_setup = time.time() - t0
prover = prover_mod.DeviceProver(setup, n, circ["selectors"], circ["sigmas"])
blinding = list(range(11, 22))
times = []
for r in range(reps + 1):
    if r == reps:
        ctx.profile_enable(True)
        ctx.profile_reset()
    t0 = time.time()
    proof = prover.prove(circ["wires"], circ["public_inputs"], blinding)
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
print(json.dumps({"logn": logn, "circuit_s": round(t_circ, 2), "srs_setup_s": round(t_setup, 2),
                  "prove_s": [round(t, 4) for t in times], "proof_sha256": proof.sha256(), "stages_ms": stages,
                  "plan": ctx.msm_last_plan()}))
