"""GPU parity of the data-processor options around the voxelizer: spconv 1.x overflow semantics, DOUBLE_FLIP and
use_lead_xyz = False (pcdet/datasets/processor/data_processor.py:16-61,116-130,158-178), against the oracle."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from hgsfusion_b200 import synthetic
from hgsfusion_b200.data_processor import TransformPointsToVoxels
from hgsfusion_b200.ops import PillarPath
from oracle import oracle
from util import bits_equal, device_pfn, oracle_pfn

pytestmark = pytest.mark.gpu


def oracle_batch(pts, offs, geom, P, mv, F, spconv1_break):
    vox, co, nu = [], [], []
    for b in range(len(offs) - 1):
        v, c, k = oracle.voxelize(pts[offs[b]:offs[b + 1]], geom, P, mv, F=F, xcol=1, spconv1_break=spconv1_break)
        vox.append(v); nu.append(k)
        co.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], axis=1))
    return np.concatenate(vox), np.concatenate(co), np.concatenate(nu)


@pytest.mark.parametrize("config,P,mv,n", [("vod", 32, 600, 6000), ("vod", 5, 1500, 9000), ("tj4d", 10, 300, 4000), ("stress", 40, 2000, 30000)])
def test_spconv1_break_semantics(cuda, config, P, mv, n):
    """max_voxels binding: spconv 1.x stops the frame's voxelization at the first refused pillar, spconv 2.x only refuses new pillars.
    Both against the oracle, bit for bit -- pillarize alone and the fused path (features + canvas)."""
    cfg = synthetic.CONFIGS[config]
    F = cfg["F"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    frames = [synthetic.make_frame(n, cfg["pc_range"], F, 60 + b, "clustered") for b in range(3)]
    frames[1] = frames[1][:mv // 2]                      # a frame that does NOT overflow sits between two that do
    pts, offs = synthetic.batch_points(frames)
    w = synthetic.make_pfn(F + 6, 64, 2)
    results = {}
    for ver in (1, 2):
        vox, co, nu = oracle_batch(pts, offs, geom, P, mv, F, spconv1_break=(ver == 1))
        path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, mv, F, spconv_version=ver)
        got = path.pillarize(torch.from_numpy(pts).to(cuda), 3).trim()
        assert got["num_pillars"] == vox.shape[0]
        assert np.array_equal(got["voxel_coords"].cpu().numpy(), co) and np.array_equal(got["voxel_num_points"].cpu().numpy(), nu)
        assert bits_equal(got["voxels"].cpu().numpy(), vox)
        res = path.points_to_bev(torch.from_numpy(pts).to(cuda), 3, device_pfn(w, cuda), want_voxels=True).trim()
        pf = oracle.pillar_vfe(vox, co, nu, geom, oracle_pfn(w))
        assert np.array_equal(res["voxel_num_points"].cpu().numpy(), nu) and bits_equal(res["voxels"].cpu().numpy(), vox)
        assert bits_equal(res["pillar_features"].cpu().numpy(), pf)
        canvas = oracle.pointpillar_scatter(pf, co, 3, 64, int(geom.grid[1]), int(geom.grid[0]))
        assert bits_equal(res["spatial_features"].cpu().numpy(), canvas)
        results[ver] = nu
    # the two semantics really differ on this input (same pillars, fewer points with the break)
    assert results[1].sum() < results[2].sum() and len(results[1]) == len(results[2])


def test_double_flip_and_use_lead_xyz(cuda):
    cfg = synthetic.CONFIGS["vod"]
    # a range symmetric in x and y so that the mirrored clouds stay (mostly) inside, as nuScenes-style DOUBLE_FLIP configs have
    rng = [-25.6, -25.6, -3, 25.6, 25.6, 2]
    geom = oracle.Geometry(rng, cfg["voxel_size"])
    f = synthetic.make_frame(5000, rng, 7, 5, "clustered")
    conf = SimpleNamespace(VOXEL_SIZE=cfg["voxel_size"], MAX_POINTS_PER_VOXEL=10, MAX_NUMBER_OF_VOXELS={"train": 900, "test": 1200},
                           DOUBLE_FLIP=True)
    proc = TransformPointsToVoxels(conf, rng, 7, mode="test")
    assert list(proc.grid_size) == [320, 320, 1]
    for lead in (True, False):
        d = proc(dict(points=torch.from_numpy(f).to(cuda), use_lead_xyz=lead))
        assert isinstance(d["voxels"], list) and len(d["voxels"]) == 4
        for i, (fx, fy) in enumerate([(1, 1), (1, -1), (-1, 1), (-1, -1)]):        # original, yflip, xflip, xyflip
            g = f.copy(); g[:, 0] *= fx; g[:, 1] *= fy
            v, c, k = oracle.voxelize(g, geom, 10, 1200, F=7, xcol=0)
            assert np.array_equal(d["voxel_coords"][i].cpu().numpy(), c) and np.array_equal(d["voxel_num_points"][i].cpu().numpy(), k)
            assert bits_equal(d["voxels"][i].cpu().numpy(), v if lead else v[..., 3:])
    # without DOUBLE_FLIP: plain tensors, train-mode voxel budget
    conf.DOUBLE_FLIP = False
    proc = TransformPointsToVoxels(conf, rng, 7, mode="train")
    d = proc(dict(points=torch.from_numpy(f).to(cuda)))
    v, c, k = oracle.voxelize(f, geom, 10, 900, F=7, xcol=0)
    assert torch.is_tensor(d["voxels"]) and bits_equal(d["voxels"].cpu().numpy(), v) and np.array_equal(d["voxel_coords"].cpu().numpy(), c)
