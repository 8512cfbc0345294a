"""Multi-GPU device prover: commitments sharded by index range over 2 GPUs, NCCL gather of
the partial points; the proof must be the single-GPU proof.  Needs two visible GPUs (gpurun --gpus 2)."""
import importlib
import json
import os
import subprocess
import sys

import pytest

from tests._bpk import bpk, ROOT

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_sharded_prover_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu", "prove_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    got = json.loads(line[len("RESULT "):])
    assert got["world"] == 2 and got["ranks_agree"]
    assert got["kat"] == "479cc377c535fd831b5fcaf30af5c2756c535a3ddbc20589ab6759843e974967"
    # single-GPU proof of the same chain circuit
    prover_mod = importlib.import_module("baby-plonk-rust_b200.prover")
    synthetic = importlib.import_module("baby-plonk-rust_b200.synthetic")
    ctx = bpk.Context(0)
    n = 1 << 12
    circ = synthetic.chain_circuit(n, n - 3, seed=2)
    setup = bpk.Setup.generate_srs(n + 8, 101, ctx)
    prover = prover_mod.DeviceProver(setup, n, circ["selectors"], circ["sigmas"])
    assert prover.prove(circ["wires"], circ["public_inputs"], list(range(11, 22))).sha256() == got["chain12"]
    assert got["chain12_cached"] == got["chain12"]
    ctx.close()
