"""Edge cases and full-size checks of the fused GPU path: empty / ragged frames, heavy truncation
(more than 32 arrivals in a cell), max_voxels overflow, all points in one cell, hand-computed cases,
grids whose width is not a multiple of the tile or of 4 (no-TMA store path), determinism, and the
BASELINE.json shapes through size-independent properties plus a direct oracle comparison."""
import json
import os

import numpy as np
import pytest
import torch

from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from oracle import oracle
from util import bits_equal, device_pfn, geom_for, oracle_pfn

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def both(pts, offs, pc_range, voxel_size, P, mv, F, cuda, seed=0, want_voxels=True, frame_offsets=False):
    geom = oracle.Geometry(pc_range, voxel_size)
    w = synthetic.make_pfn(F + 6, 64, seed)
    ref = oracle.points_to_bev(pts, offs, geom, oracle_pfn(w), P, mv, F=F, xcol=1)
    path = PillarPath(np.asarray(pc_range, dtype=np.float32), voxel_size, P, mv, F)
    B = len(offs) - 1
    fo = torch.from_numpy(np.asarray(offs, dtype=np.int32)).to(cuda) if frame_offsets else None
    res = path.points_to_bev(torch.from_numpy(pts).to(cuda), B, device_pfn(w, cuda), want_voxels=want_voxels, frame_offsets=fo)
    got = res.trim()
    assert got["num_pillars"] == ref["num_pillars"]
    assert np.array_equal(res.num_pillars[1:].cpu().numpy(), ref["frame_pillars"])
    assert np.array_equal(got["voxel_coords"].cpu().numpy(), ref["voxel_coords"])
    assert np.array_equal(got["voxel_num_points"].cpu().numpy(), ref["voxel_num_points"])
    if want_voxels:
        assert bits_equal(got["voxels"].cpu().numpy(), ref["voxels"])
    assert bits_equal(got["pillar_features"].cpu().numpy(), ref["pillar_features"])
    assert bits_equal(got["spatial_features"].cpu().numpy(), ref["spatial_features"])
    return ref, got


def test_hand_computed_cases_through_the_gpu(cuda):
    d = json.load(open(os.path.join(HERE, "golden", "voxelize_cases.json")))
    for case in d["cases"]:
        rows = [[0.0] + [float(v) for v in r] for r in case["points"]]
        pts = np.asarray(rows, dtype=np.float32).reshape(-1, 5)
        path = PillarPath(np.asarray(d["pc_range"], dtype=np.float32), d["voxel_size"], case["P"], case["max_voxels"], 4)
        res = path.pillarize(torch.from_numpy(pts).to(cuda), 1).trim()
        assert np.array_equal(res["voxel_coords"].cpu().numpy()[:, 1:], np.asarray(case["coords"], dtype=np.int32).reshape(-1, 3)), case["name"]
        assert np.array_equal(res["voxel_num_points"].cpu().numpy(), np.asarray(case["num"], dtype=np.int32)), case["name"]
        exp = np.asarray(case["voxels"], dtype=np.float32).reshape(-1, case["P"], 4)
        assert bits_equal(res["voxels"].cpu().numpy(), exp), case["name"]


def test_empty_batch_and_empty_frames(cuda):
    cfg = synthetic.CONFIGS["vod"]
    empty = np.zeros((0, 8), dtype=np.float32)
    ref, got = both(empty, [0, 0, 0], cfg["pc_range"], cfg["voxel_size"], 32, 100, 7, cuda)
    assert got["num_pillars"] == 0 and not got["spatial_features"].any()
    # ragged: frames 0 and 2 empty, 1 and 3 populated; both ways of describing the frames
    f1 = synthetic.make_frame(700, cfg["pc_range"], 7, 1)
    f3 = synthetic.make_frame(300, cfg["pc_range"], 7, 2)
    pts, _ = synthetic.batch_points([f1, f3])
    pts[pts[:, 0] == 1, 0] = 3
    pts[pts[:, 0] == 0, 0] = 1
    offs = [0, 0, 700, 700, 1000, 1000]
    both(pts, offs, cfg["pc_range"], cfg["voxel_size"], 32, 40000, 7, cuda)
    both(pts, offs, cfg["pc_range"], cfg["voxel_size"], 32, 40000, 7, cuda, frame_offsets=True)


def test_all_points_outside_range(cuda):
    cfg = synthetic.CONFIGS["vod"]
    pts = synthetic.make_frame(500, cfg["pc_range"], 7, 3)
    pts[:, 0] += 1000.0
    bp, offs = synthetic.batch_points([pts])
    ref, got = both(bp, offs, cfg["pc_range"], cfg["voxel_size"], 8, 100, 7, cuda)
    assert got["num_pillars"] == 0


@pytest.mark.parametrize("P", [1, 5, 32])
def test_heavy_truncation_many_arrivals_per_cell(cuda, P):
    """Hundreds of points in a handful of cells: exercises the > 32 arrivals selection path."""
    rng = np.random.default_rng(5)
    pc_range, vs = [0, -4, -3, 8, 4, 2], [0.5, 0.5, 5]
    n = 6000
    xyz = np.stack([rng.uniform(0, 3.0, n), rng.uniform(-1.5, 1.5, n), rng.uniform(-3, 2, n)], axis=1)
    feats = rng.normal(size=(n, 4))
    f = np.concatenate([xyz, feats], axis=1).astype(np.float32)
    bp, offs = synthetic.batch_points([f, f[::-1].copy()])
    ref, got = both(bp, offs, pc_range, vs, P, 10000, 7, cuda)
    assert ref["voxel_num_points"].max() == P


def test_single_cell_gets_everything(cuda):
    pc_range, vs = [0, -4, -3, 8, 4, 2], [0.5, 0.5, 5]
    rng = np.random.default_rng(6)
    n = 5000
    f = np.concatenate([np.stack([rng.uniform(1.0, 1.49, n), rng.uniform(0.0, 0.49, n), rng.uniform(-3, 2, n)], 1),
                        rng.normal(size=(n, 4))], axis=1).astype(np.float32)
    bp, offs = synthetic.batch_points([f])
    ref, got = both(bp, offs, pc_range, vs, 32, 10, 7, cuda)
    assert got["num_pillars"] == 1 and int(got["voxel_num_points"][0]) == 32


def test_max_voxels_overflow_per_frame(cuda):
    cfg = synthetic.CONFIGS["vod"]
    pts, offs = synthetic.make_batch("vod", 3, 3000, "uniform", seed0=12)
    ref, got = both(pts, offs, cfg["pc_range"], cfg["voxel_size"], 4, 257, 7, cuda)
    assert list(ref["frame_pillars"]) == [257, 257, 257]


@pytest.mark.parametrize("nx_cells", [50, 45, 33])
def test_grid_width_not_multiple_of_tile_or_four(cuda, nx_cells):
    """nx % 32 != 0 exercises ragged tiles; nx % 4 != 0 the plain-store path (no TMA: row pitch not 16-byte)."""
    pc_range, vs = [0, -3, -3, float(nx_cells) * 0.2, 3, 2], [0.2, 0.2, 5]
    geom = oracle.Geometry(pc_range, vs)
    assert geom.grid[0] == nx_cells
    rng = np.random.default_rng(nx_cells)
    frames = []
    for b in range(2):
        n = 1500
        xyz = np.stack([rng.uniform(-0.5, nx_cells * 0.2 + 0.5, n), rng.uniform(-3.2, 3.2, n), rng.uniform(-3, 2, n)], 1)
        frames.append(np.concatenate([xyz, rng.normal(size=(n, 4))], 1).astype(np.float32))
    bp, offs = synthetic.batch_points(frames)
    both(bp, offs, pc_range, vs, 6, 40000, 7, cuda)


def test_deterministic_across_runs(cuda):
    cfg = synthetic.CONFIGS["vod"]
    pts, offs = synthetic.make_batch("vod", 4, 8000, "clustered", seed0=30)
    w = synthetic.make_pfn(13, 64, 0)
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, 7)
    d = torch.from_numpy(pts).to(cuda)
    pf = device_pfn(w, cuda)
    a = path.points_to_bev(d, 4, pf).trim()
    for _ in range(3):
        b = path.points_to_bev(d, 4, pf).trim()
        for k in ("voxel_coords", "voxel_num_points", "pillar_features", "spatial_features"):
            assert torch.equal(a[k], b[k]), k


def _properties(res, pts, B, P, C, ny, nx):
    """Size-independent properties of a correct result."""
    got = res.trim()
    M = got["num_pillars"]
    co = got["voxel_coords"].long()
    num = got["voxel_num_points"]
    feats, canvas = got["pillar_features"], got["spatial_features"]
    assert canvas.shape == (B, C, ny, nx)
    assert int(res.num_pillars[1:].sum()) == M
    assert (num >= 1).all() and (num <= P).all()
    assert (co[:, 0] >= 0).all() and (co[:, 0] < B).all() and (co[:, 1] == 0).all()
    assert (co[:, 2] >= 0).all() and (co[:, 2] < ny).all() and (co[:, 3] >= 0).all() and (co[:, 3] < nx).all()
    key = (co[:, 0] * ny + co[:, 2]) * nx + co[:, 3]
    assert torch.unique(key).numel() == M                      # one pillar per cell
    assert (co[1:, 0] >= co[:-1, 0]).all()                     # frames concatenated in order
    assert torch.equal(canvas[co[:, 0], :, co[:, 2], co[:, 3]], feats)   # canvas = scattered pillar_features ...
    assert int(torch.count_nonzero(canvas)) == int(torch.count_nonzero(feats))   # ... and zero elsewhere
    assert (feats >= 0).all()                                  # post-ReLU
    return got


@pytest.mark.parametrize("config,B,n,P,mode", [("vod", 16, 30000, 32, "clustered"), ("vod", 16, 30000, 32, "uniform"),
                                               ("tj4d", 16, 30000, 32, "clustered")])
def test_baseline_shapes_full_size(cuda, config, B, n, P, mode):
    """BASELINE.json configs 2 and 3 at full size: properties + direct comparison with the oracle."""
    cfg = synthetic.CONFIGS[config]
    F = cfg["F"]
    geom = geom_for(config)
    nx, ny = int(geom.grid[0]), int(geom.grid[1])
    pts, offs = synthetic.make_batch(config, B, n, mode)
    w = synthetic.make_pfn(F + 6, 64, 0)
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, 40000, F)
    res = path.points_to_bev(torch.from_numpy(pts).to(cuda), B, device_pfn(w, cuda))
    got = _properties(res, pts, B, P, 64, ny, nx)
    oracle.set_num_threads(8)
    ref = oracle.points_to_bev(pts, offs, geom, oracle_pfn(w), P, 40000, F=F, xcol=1, want_canvas=False, want_voxels=False)
    oracle.set_num_threads(1)
    assert np.array_equal(got["voxel_coords"].cpu().numpy(), ref["voxel_coords"])
    assert np.array_equal(got["voxel_num_points"].cpu().numpy(), ref["voxel_num_points"])
    assert bits_equal(got["pillar_features"].cpu().numpy(), ref["pillar_features"])


def test_density_stress_shape(cuda):
    """BASELINE.json config 4 (200k points/frame, 0.1 m pillars, 512x512), one GPU's share of frames."""
    config, B, n, P = "stress", 4, 200000, 32
    cfg = synthetic.CONFIGS[config]
    geom = geom_for(config)
    pts, offs = synthetic.make_batch(config, B, n, "clustered")
    w = synthetic.make_pfn(13, 64, 0)
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, 150000, 7)
    res = path.points_to_bev(torch.from_numpy(pts).to(cuda), B, device_pfn(w, cuda))
    got = _properties(res, pts, B, P, 64, 512, 512)
    oracle.set_num_threads(8)
    ref = oracle.points_to_bev(pts, offs, geom, oracle_pfn(w), P, 150000, F=7, xcol=1, want_canvas=False, want_voxels=False)
    oracle.set_num_threads(1)
    assert np.array_equal(got["voxel_coords"].cpu().numpy(), ref["voxel_coords"])
    assert np.array_equal(got["voxel_num_points"].cpu().numpy(), ref["voxel_num_points"])
    assert bits_equal(got["pillar_features"].cpu().numpy(), ref["pillar_features"])


@pytest.mark.parametrize("seed", range(16))
def test_random_configurations_bit_exact(cuda, seed):
    """Seeded sweep over geometry, limits and input pathologies; every output bit-compared with the oracle."""
    rng = np.random.default_rng(1000 + seed)
    F = int(rng.choice([4, 5, 6, 7, 8]))
    nx, ny = int(rng.integers(3, 200)), int(rng.integers(1, 120))
    if seed % 3 == 0:
        nx = int(rng.choice([32, 64, 96, 128]))                   # TMA store path (nx % 4 == 0) with whole tiles
    vs = [float(rng.choice([0.1, 0.16, 0.25, 0.5])), float(rng.choice([0.1, 0.16, 0.4])), 4.0]
    x0, y0 = float(rng.uniform(-20, 5)), float(rng.uniform(-30, 0))
    pc_range = [x0, y0, -3.0, x0 + nx * vs[0], y0 + ny * vs[1], 1.0]
    geom = oracle.Geometry(pc_range, vs)
    if geom.grid[2] != 1:
        pytest.skip("rounding of the synthetic range gave nz != 1")
    B = int(rng.integers(1, 6))
    P = int(rng.choice([1, 2, 3, 5, 10, 20, 32]))
    mv = int(rng.choice([1, 7, 50, 400, 40000]))
    counts = [int(rng.integers(0, 2500)) if rng.random() > 0.15 else 0 for _ in range(B)]
    rows = []
    for b, nb in enumerate(counts):
        p = np.empty((nb, 1 + F), dtype=np.float32)
        p[:, 0] = b
        lo, hi = np.array(pc_range[:3]), np.array(pc_range[3:])
        p[:, 1:4] = (lo + (hi - lo) * (rng.random((nb, 3)) * 1.1 - 0.05)).astype(np.float32)   # some outside
        k = nb // 2                                                                         # half the points in a few cells
        if k:
            hot = lo[:2] + (hi - lo)[:2] * rng.random((3, 2))
            p[:k, 1:3] = (hot[rng.integers(0, 3, k)] + rng.normal(0, 0.2, (k, 2))).astype(np.float32)
        p[:, 4:] = rng.standard_normal((nb, F - 3)).astype(np.float32)
        if nb > 10:
            p[rng.integers(0, nb, 3), 1] = [np.nan, np.inf, -np.inf]        # rejected by the range test
            p[rng.integers(0, nb)] = p[rng.integers(0, nb)]                 # an exact duplicate point
            p[rng.integers(0, nb), 1:4] = [pc_range[0], pc_range[1], pc_range[2]]   # exactly on the lower corner: inside
            p[rng.integers(0, nb), 1] = pc_range[3]                         # exactly on the upper bound: outside
            p[:, 0] = b
        rows.append(p[rng.permutation(nb)])
    pts = np.concatenate(rows) if rows else np.zeros((0, 1 + F), np.float32)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    both(pts, offs, pc_range, vs, P, mv, F, cuda, seed=seed, want_voxels=bool(seed & 1), frame_offsets=bool(seed & 2))


def test_pfn_requested_without_feature_or_canvas_outputs(cuda):
    """hgsf_points_to_bev with pillar_features == NULL and spatial_features == NULL (the header allows any output to be
    NULL): the pillar-major kernel must not store through the null feature pointer; coords / counts still come out."""
    cfg = synthetic.CONFIGS["vod"]
    pts, offs = synthetic.make_batch("vod", 2, 3000, "clustered", seed0=5)
    w = synthetic.make_pfn(13, 64, 0)
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, 7)
    dpts = torch.from_numpy(pts).to(cuda)
    res = path.points_to_bev(dpts, 2, device_pfn(w, cuda), want_features=False, want_canvas=False)
    torch.cuda.synchronize()
    got = res.trim()
    assert res.pillar_features is None and res.spatial_features is None
    ref = oracle.points_to_bev(pts, offs, oracle.Geometry(cfg["pc_range"], cfg["voxel_size"]), oracle_pfn(w), 32, 40000, F=7, xcol=1)
    assert got["num_pillars"] == ref["num_pillars"]
    assert np.array_equal(got["voxel_coords"].cpu().numpy(), ref["voxel_coords"])
    assert np.array_equal(got["voxel_num_points"].cpu().numpy(), ref["voxel_num_points"])
    # features only (no canvas): the pillar-major kernel with the PFN
    res = path.points_to_bev(dpts, 2, device_pfn(w, cuda), want_features=True, want_canvas=False)
    assert bits_equal(res.trim()["pillar_features"].cpu().numpy(), ref["pillar_features"])
    # canvas only (no pillar_features rows)
    res = path.points_to_bev(dpts, 2, device_pfn(w, cuda), want_features=False, want_canvas=True)
    assert bits_equal(res.trim()["spatial_features"].cpu().numpy(), ref["spatial_features"])


@pytest.mark.parametrize("P", [33, 42, 100, 128])
def test_more_than_32_points_per_pillar(cuda, P):
    """MAX_POINTS_PER_VOXEL above 32 (KITTI PointPillars uses 32, the VoD radar configs up to 100): pillars with 1 .. several
    hundred arrivals, so that untruncated (cnt <= P), truncated (cnt > P) and multi-round (cnt > 32) pillars all occur; P = 42 also
    exercises the tail of torch's 4-way interleaved slot sum (slots beyond 4*floor(P/4))."""
    cfg = synthetic.CONFIGS["vod"]
    rng = np.random.default_rng(P)
    frames = []
    for b in range(2):
        f = synthetic.make_frame(5000, cfg["pc_range"], 7, 40 + b, "clustered")
        # three very dense spots: up to a few hundred points per 0.16 m cell
        for k in range(3):
            c = rng.uniform([5, -20], [45, 20])
            f[k * 900:(k + 1) * 900, :2] = (c + rng.normal(0, 0.12, size=(900, 2))).astype(np.float32)
        frames.append(f)
    pts, offs = synthetic.batch_points(frames)
    ref, got = both(pts, offs, cfg["pc_range"], cfg["voxel_size"], P, 40000, 7, cuda)
    n = ref["voxel_num_points"]
    assert n.max() == P and (n > 32).any() and ((n > 1) & (n < 32)).any()
    # pillarize alone (no PFN: the generic-F kernel) with the padded voxels tensor
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, 40000, 7)
    r = path.pillarize(torch.from_numpy(pts).to(cuda), 2).trim()
    assert bits_equal(r["voxels"].cpu().numpy(), ref["voxels"])
    assert np.array_equal(r["voxel_num_points"].cpu().numpy(), n)


@pytest.mark.parametrize("heavy_pts,chunk", [("0", "1"), ("1", "2"), ("3", "1"), ("1000000", "2")])
def test_tile_hand_out_variants_do_not_change_results(cuda, monkeypatch, heavy_pts, chunk):
    """The consumer's scheduling knobs -- which tiles k_front lists as heavy (handed out first, stepped over by the moving
    window) and how many tiles a ticket covers -- must not change a bit: thresholds that list every occupied tile, almost all,
    some, none; one and two tiles per ticket; clustered points with a binding max_voxels."""
    monkeypatch.setenv("HGSF_HEAVY_PTS", heavy_pts)
    monkeypatch.setenv("HGSF_TILE_CHUNK", chunk)
    cfg = synthetic.CONFIGS["vod"]
    pts, offs = synthetic.make_batch("vod", 3, 4000, "clustered", seed0=21)
    ref, got = both(pts, offs, cfg["pc_range"], cfg["voxel_size"], 4, 900, 7, cuda)
    assert ref["num_pillars"] == 2700 and ref["voxel_num_points"].max() == 4       # both limits bind
