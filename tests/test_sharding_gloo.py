"""N>1 path on CPU: two gloo ranks shard the render exactly like bench.py shards it over GPUs (epochs for the
stochastic pass, rows for the Whitted pass), sum-all-reduce their buffers, and must reproduce the unsharded
render.  The CPU oracle stands in for the CUDA renderer here (no GPU in this container); the sharding host logic
(b200rt.sharding) and the collective are the things under test."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_binding as ob
    from b200rt.sharding import shard_range
    b = ob.b
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    world_scene = b.World.fixture()
    cam = b.fixture_camera()
    params = b.default_params(width=96, height=64, seed=11)
    # epochs: [g*E/G, (g+1)*E/G)
    e0, en = shard_range(5, rank, world)
    acc, _ = ob.render_distributed(world_scene.scene(), cam, params, e0, en, n_threads=2)
    t = torch.from_numpy(acc)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    # rows: disjoint bands into zero frames, then a sum all-reduce == gather
    r0, rn = shard_range(params.height, rank, world)
    p_rows = b.copy_params(params, row_begin=r0, row_count=rn)
    rgb, prim, _ = ob.render_whitted(world_scene.scene(), cam, p_rows, n_threads=2)
    rgb[:r0] = 0; rgb[r0 + rn:] = 0
    tr = torch.from_numpy(rgb)
    dist.all_reduce(tr, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, "acc.npy"), t.numpy())
        np.save(os.path.join(out_dir, "rgb.npy"), tr.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions(b200rt):
    from b200rt.sharding import all_shards, shard_range
    for total in (0, 1, 5, 256, 2160, 1081):
        for world in (1, 2, 3, 4, 8):
            parts = all_shards(total, world)
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1                      # contiguous, no gap, no overlap
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert shard_range(256, 3, 8) == (96, 32)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_ranks_reproduce_the_unsharded_render(b200rt, oracle, fixture_world, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=96, height=64, seed=11)
    full_acc, _ = oracle.render_distributed(fixture_world.scene(), cam, params, 0, 5)
    full_rgb, _, _ = oracle.render_whitted(fixture_world.scene(), cam, params)
    acc = np.load(tmp_path / "acc.npy")
    rgb = np.load(tmp_path / "rgb.npy")
    assert np.array_equal(acc[..., 3], full_acc[..., 3])                    # sample counts add exactly
    np.testing.assert_allclose(acc[..., :3], full_acc[..., :3], rtol=2e-6, atol=1e-7)   # fp32 sum order only
    assert np.array_equal(rgb.view(np.uint32), full_rgb.view(np.uint32))    # row bands: bitwise
