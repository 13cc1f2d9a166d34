"""CPU, world_size 2, gloo: the N>1 decomposition used by bench.py --gpus N (tile i -> rank i mod N, replicated
dataset, collectives only for the timing / tile counters).  Each rank renders its shard with the oracle (there is no
GPU here); the union must be every tile exactly once and the reduced job totals must be whole-job numbers."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import FixtureInputs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import oracle
    from osm_renderer_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = FixtureInputs()
    tiles, begins, areas = fx.batches["16"]
    t, b, a, idx = sharding.shard_batch(tiles, begins, areas, rank, world)
    imgs = oracle.draw_tiles(fx.bin, fx.table, t, b, a, fx.canvas_rgb, fx.use_caps_for_dashes)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx, imgs=np.stack(imgs))
    secs, total = sharding.reduce_job(dist, 1.0 + rank, len(idx))
    assert secs == float(world) and total == len(tiles)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_round_robin_covers_every_tile_once(tmp_path, fx):
    import oracle

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    tiles, begins, areas = fx.batches["16"]
    want = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes))
    seen = np.zeros(len(tiles), dtype=int)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        for i, img in zip(z["idx"], z["imgs"]):
            seen[i] += 1
            assert (img == want[i]).all()
    assert (seen == 1).all()


def test_weak_scaling_request_list_gives_every_rank_one_full_batch():
    from osm_renderer_b200 import sharding

    for world in (1, 2, 4, 8):
        req = sharding.weak_scaling_request_list(1024, world)
        for r in range(world):
            mine = req[sharding.shard_indices(len(req), r, world)]
            assert (mine == np.arange(1024)).all()
