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
