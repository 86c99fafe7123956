"""The voxelizer restatement (oracle.voxelize, C) against hand-computed known answers, against its
independent pure-Python twin, and against the structural invariants of SURVEY.md §8c.
spconv is not available in this environment: parity with the real library is unpinned."""
import json
import os

import numpy as np
import pytest

from hgsfusion_b200 import synthetic
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    d = json.load(open(os.path.join(HERE, "golden", "voxelize_cases.json")))
    geom = oracle.Geometry(d["pc_range"], d["voxel_size"])
    return [(c["name"], c, geom) for c in d["cases"]]


def _arr(x, shape):
    conv = lambda v: float(v)      # "nan" / "inf" strings
    a = np.array([[conv(v) for v in row] for row in x], dtype=np.float32) if len(x) else np.zeros((0,) + shape, np.float32)
    return a


@pytest.mark.parametrize("name,case,geom", _cases(), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("impl", ["c", "py"])
def test_hand_cases(name, case, geom, impl):
    pts = _arr(case["points"], (4,))
    fn = oracle.voxelize if impl == "c" else oracle.voxelize_py
    vox, coords, num = fn(pts, geom, case["P"], case["max_voxels"])
    exp_coords = np.asarray(case["coords"], dtype=np.int32).reshape(-1, 3)
    assert np.array_equal(coords, exp_coords)
    assert np.array_equal(num, np.asarray(case["num"], dtype=np.int32))
    exp_vox = np.asarray(case["voxels"], dtype=np.float32).reshape(-1, case["P"], 4)
    assert vox.shape == exp_vox.shape
    assert np.array_equal(vox.view(np.uint32), exp_vox.view(np.uint32))   # bit pattern: -0.0 stays -0.0


def test_grid_size_matches_reference_expression():
    # data_processor.py:135-136
    for cfg in synthetic.CONFIGS.values():
        rng = np.array(cfg["pc_range"], dtype=np.float32)
        ref = np.round((rng[3:6] - rng[0:3]) / np.array(cfg["voxel_size"])).astype(np.int64)
        assert np.array_equal(oracle.Geometry(cfg["pc_range"], cfg["voxel_size"]).grid, ref)
    assert list(oracle.Geometry(synthetic.CONFIGS["vod"]["pc_range"], synthetic.CONFIGS["vod"]["voxel_size"]).grid) == [320, 320, 1]
    assert list(oracle.Geometry(synthetic.CONFIGS["tj4d"]["pc_range"], synthetic.CONFIGS["tj4d"]["voxel_size"]).grid) == [432, 496, 1]
    assert list(oracle.Geometry(synthetic.CONFIGS["stress"]["pc_range"], synthetic.CONFIGS["stress"]["voxel_size"]).grid) == [512, 512, 1]


@pytest.mark.parametrize("config,n,P,mv,mode", [("vod", 600, 4, 40000, "clustered"), ("vod", 500, 2, 60, "uniform"),
                                                 ("tj4d", 400, 32, 40000, "clustered"), ("stress", 700, 3, 150, "clustered")])
def test_c_matches_python_twin(config, n, P, mv, mode):
    cfg = synthetic.CONFIGS[config]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    pts = synthetic.make_frame(n, cfg["pc_range"], cfg["F"], seed=7, mode=mode, oob_fraction=0.05)
    a = oracle.voxelize(pts, geom, P, mv)
    b = oracle.voxelize_py(pts, geom, P, mv)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_invariants():
    cfg = synthetic.CONFIGS["vod"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    P, mv = 3, 300
    pts = synthetic.make_frame(3000, cfg["pc_range"], cfg["F"], seed=1, mode="clustered", oob_fraction=0.05)
    vox, coords, num, pp = oracle.voxelize(pts, geom, P, mv, return_point_pillar=True)
    M = coords.shape[0]
    assert M == mv                                            # overflow happened
    assert len({tuple(c) for c in coords}) == M               # one pillar per cell
    assert num.min() >= 1 and num.max() <= P
    g = geom.grid
    assert (coords[:, 0] == 0).all() and (coords[:, 1] >= 0).all() and (coords[:, 1] < g[1]).all() and (coords[:, 2] < g[0]).all()
    # pillar order = order of the first point; slot order = input order; padding is zero
    first = [np.flatnonzero(pp == v)[0] for v in range(M)]
    assert first == sorted(first)
    for v in range(0, M, 17):
        idx = np.flatnonzero(pp == v)
        assert len(idx) == num[v] and (np.diff(idx) > 0).all()
        assert np.array_equal(vox[v, :num[v]], pts[idx])
        assert not vox[v, num[v]:].any()
    # the inclusive range pre-mask (common_utils.py:78-81) never changes the result of the voxelizer
    keep = oracle.mask_points_by_range(pts, geom.pc_range)
    v2, c2, n2 = oracle.voxelize(pts[keep], geom, P, mv)
    assert np.array_equal(c2, coords) and np.array_equal(n2, num) and np.array_equal(v2, vox)


def test_batch_matches_per_frame_and_is_thread_invariant():
    cfg = synthetic.CONFIGS["vod"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    pts, offs = synthetic.make_batch("vod", 3, 800, "clustered", seed0=5, oob_fraction=0.03)
    w = synthetic.make_pfn(13, 64, 1)
    pfn = oracle.PfnParams(w.weight, w.gamma, w.beta, w.running_mean, w.running_var)
    oracle.set_num_threads(1)
    a = oracle.points_to_bev(pts, offs, geom, pfn, 8, 500, F=7, xcol=1)
    oracle.set_num_threads(4)
    b = oracle.points_to_bev(pts, offs, geom, pfn, 8, 500, F=7, xcol=1)
    oracle.set_num_threads(1)
    for k in ("voxels", "voxel_coords", "voxel_num_points", "pillar_features", "spatial_features", "frame_pillars"):
        assert np.array_equal(a[k], b[k]), k
    m0 = 0
    for f in range(3):
        v, c, n = oracle.voxelize(pts[offs[f]:offs[f + 1]], geom, 8, 500, F=7, xcol=1)
        m = c.shape[0]
        assert a["frame_pillars"][f] == m
        assert np.array_equal(a["voxel_coords"][m0:m0 + m, 1:], c) and (a["voxel_coords"][m0:m0 + m, 0] == f).all()
        assert np.array_equal(a["voxels"][m0:m0 + m], v) and np.array_equal(a["voxel_num_points"][m0:m0 + m], n)
        m0 += m
    assert m0 == a["num_pillars"]


def test_spconv1_break_hand_case():
    """max_voxels = 2 on a 1 m grid: points land in cells A, B, A, C (refused: third pillar), A, B.
    spconv 2.x (`continue`): pillars A (3 points: #0, #2, #4) and B (2 points: #1, #5).
    spconv 1.x (`break` at point #3): pillars A (2 points: #0, #2) and B (1 point: #1) -- points #4 and #5 are never looked at."""
    geom = oracle.Geometry([0, 0, 0, 4, 4, 1], [1.0, 1.0, 1.0])
    pts = np.array([[0.5, 0.5, 0.5, 10], [1.5, 0.5, 0.5, 11], [0.6, 0.4, 0.5, 12], [2.5, 0.5, 0.5, 13], [0.7, 0.3, 0.5, 14],
                    [1.6, 0.4, 0.5, 15]], dtype=np.float32)
    for twin in (oracle.voxelize, oracle.voxelize_py):
        v, c, n = twin(pts, geom, 4, 2, F=4, xcol=0)
        assert c.tolist() == [[0, 0, 0], [0, 0, 1]] and n.tolist() == [3, 2]
        assert v[0, :3, 3].tolist() == [10, 12, 14] and v[1, :2, 3].tolist() == [11, 15]
        v, c, n = twin(pts, geom, 4, 2, F=4, xcol=0, spconv1_break=True)
        assert c.tolist() == [[0, 0, 0], [0, 0, 1]] and n.tolist() == [2, 1]
        assert v[0, :, 3].tolist() == [10, 12, 0, 0] and v[1, :, 3].tolist() == [11, 0, 0, 0]
