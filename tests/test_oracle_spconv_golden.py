"""The voxelizer against spconv's own outputs -- when tests/golden/spconv_v*.npz exist.  They are produced by
tests/golden/make_spconv_golden.py on a machine where spconv imports (it is absent from the build image, so until someone runs
that script the voxelizer's parity stays 'unpinned': hand cases + the independent Python twin + structural invariants)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "spconv_v*.npz")))


def _check(d, got):
    vox, coords, num = got
    assert np.array_equal(coords, d["coords"].astype(np.int32))
    assert np.array_equal(num, d["num_points"].astype(np.int32))
    assert np.array_equal(vox.view(np.uint32), np.ascontiguousarray(d["voxels"], dtype=np.float32).view(np.uint32))


@pytest.mark.skipif(not FIX, reason="no spconv fixtures: run tests/golden/make_spconv_golden.py where spconv is installed")
@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_voxelizer_matches_spconv(path):
    d = np.load(path)
    geom = oracle.Geometry(d["pc_range"], [float(v) for v in d["voxel_size"]])
    pts = d["points"]
    _check(d, oracle.voxelize(pts, geom, int(d["P"]), int(d["max_voxels"]), F=pts.shape[1], xcol=0,
                              spconv1_break=(int(d["spconv_major"]) == 1)))


@pytest.mark.gpu
@pytest.mark.skipif(not FIX, reason="no spconv fixtures: run tests/golden/make_spconv_golden.py where spconv is installed")
@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[:-4])
def test_gpu_voxelizer_matches_spconv(cuda, path):
    import torch
    from hgsfusion_b200.ops import PillarPath
    d = np.load(path)
    pts = d["points"]
    pp = PillarPath(d["pc_range"], [float(v) for v in d["voxel_size"]], int(d["P"]), int(d["max_voxels"]), pts.shape[1],
                    spconv_version=int(d["spconv_major"]))
    offs = torch.tensor([0, pts.shape[0]], dtype=torch.int32, device=cuda)
    r = pp.pillarize(torch.from_numpy(pts).to(cuda), 1, xyz_col=0, batch_col=-1, frame_offsets=offs).trim()
    _check(d, (r["voxels"].cpu().numpy(), r["voxel_coords"].cpu().numpy()[:, 1:], r["voxel_num_points"].cpu().numpy()))
