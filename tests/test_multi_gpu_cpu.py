"""Host-side logic of the sharded MSM (baby-plonk-rust_b200/multi_gpu.py) on CPU: world_size 2 over
gloo.  Each rank computes the partial sum of its index range with the CPU oracle (standing in for the
GPU kernel), the partials are all-gathered exactly as on NCCL, and every rank must end up with the
full MSM.  Checks shard bounds, gather order and that the result is replicated."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bls12_381 as O  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = importlib.import_module("baby-plonk-rust_b200.multi_gpu")
    pkg = importlib.import_module("baby-plonk-rust_b200")
    pts = O.generate_srs_points(n, 101)
    sc = O.random_fr(77, n)
    lo, hi = mg.shard_bounds(n, world, rank)
    partial = O.msm_naive(pts[lo:hi], sc[lo:hi])
    # hand the partial over as a non-normalised projective point, like the GPU kernel does
    limbs = np.array(O.g1_scale_proj(partial, 1000 + rank), dtype=np.uint64)
    t = torch.from_numpy(limbs.view(np.int64).copy())

    def sum_points(gathered):
        arr = gathered.numpy().view(np.uint64)
        assert arr.shape == (world, 18)
        # gather order == rank order
        assert [int(v) for v in arr[rank]] == [int(v) for v in limbs]
        return O.g1_sum([pkg.point_to_affine(row) for row in arr])

    total = mg.sharded_msm(t, sum_points)
    assert total == O.msm_naive(pts, sc)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array(O.g1_affine_to_proj_limbs(total), dtype=np.uint64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [13, 2])
def test_sharded_msm_gloo_world2(tmp_path, n):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "r0.npy")
    r1 = np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)


def test_shard_bounds_tile_the_range():
    mg = importlib.import_module("baby-plonk-rust_b200.multi_gpu")
    for n in (0, 1, 7, 8, (1 << 24) + 6):
        for world in (1, 2, 4, 8):
            spans = [mg.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mg.shard_bounds(10, 2, 2)


def test_slice_of_prefix_tiles_every_prefix():
    """sharded commitment of a replicated polynomial: the per-rank slices of the first `length` indices are
    disjoint, ordered and cover the prefix exactly"""
    mg = importlib.import_module("baby-plonk-rust_b200.multi_gpu")
    for n, world in ((14, 2), (1030, 4), (4104, 8), (5, 8)):
        bounds = [mg.shard_bounds(n, world, r) for r in range(world)]
        for length in (0, 1, n // 2, n - 1, n):
            covered = []
            for lo, hi in bounds:
                first, count = mg.slice_of_prefix(lo, hi, length)
                assert count >= 0 and first == lo and first + count <= hi
                covered.extend(range(first, first + count))
            assert covered == list(range(length))
