"""world_size-2 run of the multi-GPU harness logic on CPU (gloo): frames are sharded by rank with no
data-path collective; the shards' results concatenate to the unsharded result; the timing reduction
is a MAX over ranks.  The per-rank compute stand-in is the CPU oracle (test infrastructure)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hgsfusion_b200 import sharding, synthetic
from oracle import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 5
    cfg = synthetic.CONFIGS["vod"]
    pts, offs = synthetic.make_batch("vod", B, 400, "clustered", seed0=21)
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    w = synthetic.make_pfn(13, 64)
    pfn = oracle.PfnParams(w.weight, w.gamma, w.beta, w.running_mean, w.running_var)
    lo, hi = sharding.frame_range(B, world, rank)
    sp, so = sharding.shard_points(pts, offs, lo, hi)
    res = oracle.points_to_bev(sp, so, geom, pfn, 8, 300, F=7, xcol=1)
    dist.barrier()
    t = sharding.reduce_max(1.0 + rank)          # slowest rank wins
    frames = sharding.reduce_sum(hi - lo)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), coords=res["voxel_coords"], feats=res["pillar_features"],
             canvas=res["spatial_features"], t=t, frames=frames, lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert all(float(p["t"]) == 2.0 for p in parts)           # max over ranks of (1, 2)
    assert all(float(p["frames"]) == 5.0 for p in parts)
    cfg = synthetic.CONFIGS["vod"]
    pts, offs = synthetic.make_batch("vod", 5, 400, "clustered", seed0=21)
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    w = synthetic.make_pfn(13, 64)
    pfn = oracle.PfnParams(w.weight, w.gamma, w.beta, w.running_mean, w.running_var)
    full = oracle.points_to_bev(pts, offs, geom, pfn, 8, 300, F=7, xcol=1)
    coords = []
    for p in parts:
        c = p["coords"].copy()
        c[:, 0] += int(p["lo"])
        coords.append(c)
    assert np.array_equal(np.concatenate(coords), full["voxel_coords"])
    assert np.array_equal(np.concatenate([p["feats"] for p in parts]), full["pillar_features"])
    assert np.array_equal(np.concatenate([p["canvas"] for p in parts]), full["spatial_features"])
