"""CPU: the numpy training restatement (oracle/train_oracle.py) against the fixtures produced by the REFERENCE's own
PillarVFE in train mode under torch autograd (tests/golden/make_golden.py --only-train)."""
import glob
import os

import numpy as np
import pytest

from hgsfusion_b200 import synthetic
from oracle import train_oracle as to

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "train_*.npz")))


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())


def total_cotangent(d):
    g = d["grad_out"].astype(np.float64).copy()
    g[d["grad_canvas_idx"], 8:24] += d["grad_canvas_vals"]          # the canvas cotangent gathered at the pillar cells
    return g


def test_fixtures_present():
    assert len(FIXTURES) == 5


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_train_forward_backward_match_reference_autograd(path):
    d = np.load(path)
    cfgname, P, ua, wd, _ = d["meta"]
    cfg = synthetic.CONFIGS[str(cfgname)]
    feats = to.decorate(d["voxels"], d["voxel_coords"], d["voxel_num_points"], cfg["pc_range"], cfg["voxel_size"],
                        bool(int(ua)), bool(int(wd)))
    out, cache = to.pfn_train_forward(feats, d["weight"], d["gamma"], d["beta"])
    assert rel(out, d["pillar_features"]) < 2e-6                     # fp32 reference vs float64 restatement
    dW, dg, db = to.pfn_backward(cache, d["gamma"], total_cotangent(d))
    assert rel(dW, d["grad_weight"]) < 1e-5 and rel(dg, d["grad_gamma"]) < 1e-5 and rel(db, d["grad_beta"]) < 1e-5
    rm, rv = to.running_update(d["running_mean"], d["running_var"], cache["mean"], cache["var"], cache["N"])
    assert rel(rm, d["running_mean_after"]) < 1e-6 and rel(rv, d["running_var_after"]) < 1e-6


def test_padded_rows_count_in_the_statistics():
    # one pillar with 1 of 4 slots filled: mean of x over the 4 rows is x0/4, not x0
    vox = np.zeros((1, 4, 4), np.float32)
    vox[0, 0] = [1.0, 2.0, 0.5, 3.0]
    feats = to.decorate(vox, np.array([[0, 0, 12, 6]]), np.array([1]), [0, 0, -3, 8, 8, 2], [0.16, 0.16, 5])
    W = np.ones((2, 10), np.float32)
    out, cache = to.pfn_train_forward(feats, W, np.ones(2), np.zeros(2))
    x0 = feats[0, 0].astype(np.float64).sum()
    assert np.allclose(cache["mean"], x0 / 4) and cache["N"] == 4
    assert np.allclose(cache["var"], (x0 ** 2) / 4 - (x0 / 4) ** 2)
