"""CPU checks of the Path B restatement (oracle/pathb_oracle.py): hand-computed cases and invariants."""
import numpy as np

from oracle import pathb_oracle as pb


def test_reference_debug_fixture_four_points():
    # the commented fixture in pcdet/ops/pillar_ops/pillar_modules.py:60-72: (x, y) = (1,20), (1,40), (40,20), (40,40)
    xyz = np.array([[1, 20, 0], [1, 40, 0], [40, 20, 0], [40, 40, 0]], dtype=np.float32)
    r = pb.gen_indice_pairs_flat(xyz, [4], 0.16, 320, 320)
    # cells (y, x) = (125, 6), (250, 6), (125, 250), (250, 250); raster order sorts by y then x
    assert r["pillars"].tolist() == [[0, 125, 6], [0, 125, 250], [0, 250, 6], [0, 250, 250]]
    assert r["indice_pairs"].ravel().tolist() == [0, 2, 1, 3]
    assert r["point_set_indices"].tolist() == [0, 1, 2, 3] and r["pillar_set_indices"].tolist() == [0, 2, 1, 3]
    assert r["M"] == 4 and r["L"] == 4


def test_truncation_toward_zero_and_bounds():
    s = 0.5
    xyz = np.array([[-0.4, 0.1, 0], [-0.5, 0.1, 0], [0.0, -0.2, 0], [3.99, 1.99, 0], [4.0, 1.0, 0], [1.0, 2.0, 0],
                    [np.nan, 0.3, 0]], dtype=np.float32)
    r = pb.gen_indice_pairs_flat(xyz, [7], s, 4, 8)          # H = 4 (y < 2.0), W = 8 (x < 4.0)
    # int(-0.8) = 0 -> inside; int(-1.0) = -1 -> outside; y = -0.2 -> int(-0.4) = 0 inside; upper bounds exclusive;
    # NaN converts to 0 (cvt.rzi) -> inside cell 0
    assert r["indice_pairs"].ravel().tolist() == [0, -1, 0, 1, -1, -1, 0]
    assert r["pillars"].tolist() == [[0, 0, 0], [0, 3, 7]]


def test_frames_and_surplus_points():
    # 5 points, counts (2, 1): the reference assigns every point past the counted ones to the LAST frame
    xyz = np.array([[0.1, 0.1, 0]] * 5, dtype=np.float32)
    r = pb.gen_indice_pairs_flat(xyz, [2, 1], 1.0, 2, 2)
    assert r["pillars"].tolist() == [[0, 0, 0], [1, 0, 0]]
    assert r["indice_pairs"].ravel().tolist() == [0, 0, 1, 1, 1]


def test_scatter_max_and_arg_rules():
    rng = np.random.default_rng(0)
    C, L, M = 5, 200, 17
    src = rng.normal(size=(C, L)).astype(np.float32)
    idx = rng.integers(0, M - 2, size=L).astype(np.int32)      # pillars M-2, M-1 stay empty
    out = pb.scatter_max(src, idx, M)
    assert (out >= 0).all() and (out[:, M - 2:] == 0).all()
    for m in range(M - 2):
        sel = src[:, idx == m]
        exp = np.maximum(sel.max(axis=1), 0) if sel.size else np.zeros(C, np.float32)
        assert np.array_equal(out[:, m], exp)
    # a valid arg: the arg-max where it is positive
    arg = np.full((C, M), -1, dtype=np.int32)
    for c in range(C):
        for p in range(L):
            if abs(src[c, p] - out[c, idx[p]]) < 1e-5:
                arg[c, idx[p]] = c * L + p
    assert pb.check_arg(arg, src, idx, out)
    bad = arg.copy()
    bad[0, 0] = -1 if arg[0, 0] >= 0 else 0
    assert not pb.check_arg(bad, src, idx, out)


def test_gather_and_grad():
    rng = np.random.default_rng(1)
    f = rng.normal(size=(30, 7)).astype(np.float32)
    idx = rng.integers(0, 30, size=50).astype(np.int32)
    g = pb.gather_feature(f, idx)
    assert np.array_equal(g, f[idx])
    go = rng.normal(size=(50, 7)).astype(np.float32)
    gi = pb.gather_feature_grad(idx, go, 30)
    for n in range(30):
        assert np.allclose(gi[n], go[idx == n].sum(axis=0), atol=1e-5)


# ---- split_encode: pinned against the reference's DynamicPillarFeatureNet.forward (tests/golden/make_golden.py) ----
import glob
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPLIT_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "split_*.npz")))


def test_split_fixtures_present():
    assert len(SPLIT_FIXTURES) == 6


@pytest.mark.parametrize("path", SPLIT_FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_split_encode_matches_reference_fixture(path):
    d = np.load(path)
    dataset, Fin, num_input, virtual, encoding, order = d["meta"]
    xyz, cnt, feat = pb.split_encode(d["points"], d["pc_range"], encoding, dataset, int(num_input), bool(int(virtual)))
    assert np.array_equal(cnt, d["xyz_batch_cnt"]) and cnt.dtype == np.int32
    assert np.array_equal(xyz.view(np.uint32), d["xyz"].view(np.uint32))
    assert np.array_equal(feat.view(np.uint32), d["pt_features"].view(np.uint32))


def test_split_encode_hand_case():
    # Fin = 17: xyz, 12 features, flags.  (1,1) raw radar point -> real block; (0,1) virtual -> virtual block; (0,0) -> virtual block
    row = lambda b, f0, f1: [b, 10, -5, 1] + list(range(100, 112)) + [f0, f1]
    pts = np.array([row(0, 1, 1), row(0, 0, 1), row(1, 0, 0), row(0.5, 1, 1)], dtype=np.float32)
    xyz, cnt, feat = pb.split_encode(pts, [0, -25.6, -3, 51.2, 25.6, 2], "split", "vod", 29)
    assert cnt.tolist() == [2, 1]                                  # the row with frame index 0.5 matches no mask
    assert np.allclose(xyz[0], [10, 20.6, 4])
    assert feat.shape == (3, 29)
    assert feat[0, 3:15].tolist() == list(range(100, 112)) and not feat[0, 15:27].any() and feat[0, 27:].tolist() == [1, 1]
    assert feat[1, 15:27].tolist() == list(range(100, 112)) and not feat[1, 3:15].any() and feat[1, 27:].tolist() == [0, 1]
    assert feat[2, 15:27].tolist() == list(range(100, 112)) and feat[2, 27:].tolist() == [0, 0]


# ---- pillar-list consumer: the oracle's submanifold 3x3 convolution against torch's dense conv2d ----

def _random_pillars(rng, B, H, W, M_per):
    pillars = []
    bev = np.full((B, H, W), -1, dtype=np.int32)
    for b in range(B):
        cells = np.sort(rng.choice(H * W, size=min(M_per, H * W), replace=False))       # raster order like the reader
        for c in cells:
            bev[b, c // W, c % W] = len(pillars)
            pillars.append((b, c // W, c % W))
    return np.asarray(pillars, dtype=np.int32).reshape(-1, 3), bev


@pytest.mark.parametrize("Cin,Cout,H,W", [(32, 32, 12, 9), (8, 16, 5, 5), (32, 64, 3, 20)])
def test_subm_conv_oracle_matches_dense_conv2d_on_the_active_set(Cin, Cout, H, W):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(Cin + Cout)
    pillars, bev = _random_pillars(rng, 2, H, W, H * W // 3)
    M = pillars.shape[0]
    feats = rng.normal(size=(M, Cin)).astype(np.float32)
    w = rng.normal(size=(Cout, 3, 3, Cin)).astype(np.float32) * 0.1
    bias = rng.normal(size=Cout).astype(np.float32)
    nbr = pb.subm_neighbors(bev, pillars)
    assert (nbr[:, 4] == np.arange(M)).all()                      # the centre tap is the pillar itself
    got = pb.subm_conv3x3(feats, nbr, w, bias=bias)
    dense = torch.zeros(2, Cin, H, W)
    dense[pillars[:, 0], :, pillars[:, 1], pillars[:, 2]] = torch.from_numpy(feats)
    ref = F.conv2d(dense, torch.from_numpy(w).permute(0, 3, 1, 2).contiguous(), torch.from_numpy(bias), padding=1)
    ref = ref[pillars[:, 0], :, pillars[:, 1], pillars[:, 2]].numpy()
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_subm_neighbors_edges_and_frames():
    bev = np.full((2, 2, 3), -1, dtype=np.int32)
    pillars = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 2], [1, 0, 0]], dtype=np.int32)
    for i, (b, y, x) in enumerate(pillars):
        bev[b, y, x] = i
    nbr = pb.subm_neighbors(bev, pillars)
    assert nbr[0].tolist() == [-1, -1, -1, -1, 0, 1, -1, -1, -1]
    assert nbr[1].tolist() == [-1, -1, -1, 0, 1, -1, -1, -1, 2]
    assert nbr[2].tolist() == [1, -1, -1, -1, 2, -1, -1, -1, -1]
    assert nbr[3].tolist() == [-1, -1, -1, -1, 3, -1, -1, -1, -1]    # other frames are not neighbours


@pytest.mark.parametrize("H,W", [(12, 9), (7, 7), (2, 1), (16, 16)])
def test_sparse_conv_s2_oracle_matches_dense_conv2d(H, W):
    """Active set = where a stride-2 3x3 window holds a pillar (max_pool2d of the occupancy); features = dense conv2d there."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(H * 31 + W)
    pillars, bev = _random_pillars(rng, 2, H, W, max(1, H * W // 4))
    M = pillars.shape[0]
    feats = rng.normal(size=(M, 32)).astype(np.float32)
    w = (rng.normal(size=(64, 3, 3, 32)) * 0.1).astype(np.float32)
    out_pillars, out_bev, nbr = pb.sparse_conv_s2_indices(bev, pillars)
    occ = torch.from_numpy((bev >= 0).astype(np.float32))[:, None]
    act = F.max_pool2d(occ, 3, 2, 1)[:, 0].numpy() > 0
    assert np.array_equal(out_pillars, np.argwhere(act).astype(np.int32))
    assert np.array_equal(out_bev >= 0, act)
    got = pb.subm_conv3x3(feats, nbr, w)
    dense = torch.zeros(2, 32, H, W)
    dense[pillars[:, 0], :, pillars[:, 1], pillars[:, 2]] = torch.from_numpy(feats)
    ref = F.conv2d(dense, torch.from_numpy(w).permute(0, 3, 1, 2).contiguous(), None, stride=2, padding=1)
    ref = ref[out_pillars[:, 0], :, out_pillars[:, 1], out_pillars[:, 2]].numpy()
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
