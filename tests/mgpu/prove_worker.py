"""torchrun worker: the device prover with commitments sharded over WORLD_SIZE GPUs (one process per GPU,
NCCL all-gather of one partial point per rank).  Rank 0 prints one JSON line with the proof digests."""
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from oracle import plonk as P   # circuit pre-processing for the KAT circuit only (test infrastructure)

bpk = importlib.import_module("baby-plonk-rust_b200")
mg = importlib.import_module("baby-plonk-rust_b200.multi_gpu")
prover_mod = importlib.import_module("baby-plonk-rust_b200.prover")
synthetic = importlib.import_module("baby-plonk-rust_b200.synthetic")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bpk.Context(local)
    out = {}
    # 1. the reference's own test program (n = 8, SRS of 14 powers split over the ranks), blinding 1..11
    prog, wit, pub = P.reference_test_circuit()
    pad = prog.n - len(prog.gates)
    wires = [bpk.scalars_from_ints([wit[g.wires[k]] if g.wires[k] is not None else 0 for g in prog.gates] + [0] * pad)
             for k in range(3)]
    com = mg.ShardedCommitter(bpk, ctx, 14, 101, rank, world, precompute=None)
    prover = prover_mod.DeviceProver(com.setup, prog.n, [bpk.scalars_from_ints(c) for c in prog.selectors()],
                                     [bpk.scalars_from_ints(c) for c in prog.sigmas()], committer=com)
    out["kat"] = prover.prove(wires, pub, list(range(1, 12))).sha256()
    # 2. chain circuit, 2^12 rows, precomputed SRS slices
    n = 1 << 12
    circ = synthetic.chain_circuit(n, n - 3, seed=2)
    com = mg.ShardedCommitter(bpk, ctx, n + 8, 101, rank, world, precompute=0)
    prover = prover_mod.DeviceProver(com.setup, n, circ["selectors"], circ["sigmas"], committer=com)
    out["chain12"] = prover.prove(circ["wires"], circ["public_inputs"], list(range(11, 22))).sha256()
    # the same with the per-circuit sub-coset evaluations cached across proofs (second proof served from the cache)
    cached = prover_mod.DeviceProver(com.setup, n, circ["selectors"], circ["sigmas"], committer=com, cache_preprocessed=True)
    cached.prove(circ["wires"], circ["public_inputs"], list(range(1, 12)))
    out["chain12_cached"] = cached.prove(circ["wires"], circ["public_inputs"], list(range(11, 22))).sha256()
    # every rank must hold the same proof
    digest = torch.tensor(list(hashlib.sha256(json.dumps(out, sort_keys=True).encode()).digest()[:8]),
                          dtype=torch.int64, device="cuda")
    gathered = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    out["ranks_agree"] = all(bool((g == digest).all()) for g in gathered)
    out["world"] = world
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
