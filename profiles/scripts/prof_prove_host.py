import importlib, os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
bpk = importlib.import_module("baby-plonk-rust_b200")
prover_mod = importlib.import_module("baby-plonk-rust_b200.prover")
synthetic = importlib.import_module("baby-plonk-rust_b200.synthetic")
n = 1 << 20
ctx = bpk.Context(0)
circ = synthetic.chain_circuit(n, n - 3, seed=2)
setup = bpk.Setup.generate_srs(n + 8, 101, ctx).precompute(0)
prover = prover_mod.DeviceProver(setup, n, circ["selectors"], circ["sigmas"], cache_preprocessed=True)
wires = torch.from_numpy(np.stack(circ["wires"]).view(np.int64)).pin_memory()
bl = list(range(11, 22))
for _ in range(2): prover.prove(wires, circ["public_inputs"], bl)
pr = cProfile.Profile(); pr.enable()
for _ in range(3): prover.prove(wires, circ["public_inputs"], bl)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
