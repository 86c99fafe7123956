"""Pins the CPU oracle to the REFERENCE: tests/golden/vfe_*.npz hold inputs and outputs of the
reference's own PillarVFE + PointPillarScatter (imported from /root/reference by
tests/golden/make_golden.py, torch CPU, eval mode).  The oracle must reproduce them bit for bit."""
import glob
import os

import numpy as np
import pytest

from hgsfusion_b200 import synthetic
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "vfe_*.npz")))


def load(path):
    d = np.load(path)
    config, P, mv, use_abs, with_dist = d["meta"][:5]
    cfg = synthetic.CONFIGS[str(config)]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    pfn = oracle.PfnParams(d["weight"], d["gamma"], d["beta"], d["running_mean"], d["running_var"])
    return d, cfg, geom, pfn, int(P), int(mv), bool(int(use_abs)), bool(int(with_dist))


def test_fixtures_present():
    assert len(FIXTURES) >= 8


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[4:-4] for p in FIXTURES])
def test_oracle_reproduces_reference_bit_for_bit(path):
    d, cfg, geom, pfn, P, mv, use_abs, with_dist = load(path)
    # torch evaluates 1/sqrt(var+eps) through MKL VML's sqrt, which is not correctly rounded; with
    # torch's own vector the oracle must match the reference on EVERY bit
    got = oracle.pillar_vfe(d["voxels"], d["voxel_coords"], d["voxel_num_points"], geom, pfn,
                            use_absolute_xyz=use_abs, with_distance=with_dist, invstd_override=d["ref_invstd"])
    ref = d["pillar_features"]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # with the IEEE invstd (what the CUDA kernels compute) only the channels whose invstd differs may move,
    # and by no more than the 1e-5 relative bar allows by a wide margin
    ieee = oracle.pillar_vfe(d["voxels"], d["voxel_coords"], d["voxel_num_points"], geom, pfn,
                             use_absolute_xyz=use_abs, with_distance=with_dist)
    my_invstd = (np.float32(1) / np.sqrt(d["running_var"] + np.float32(1e-3))).astype(np.float32)
    moved = np.flatnonzero(my_invstd != d["ref_invstd"])
    diff_cols = np.unique(np.argwhere(ieee.view(np.uint32) != ref.view(np.uint32))[:, 1])
    assert set(diff_cols) <= set(moved)
    assert len(moved) <= 3
    assert np.abs(ieee - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[4:-4] for p in FIXTURES])
def test_scatter_reproduces_reference(path):
    d, cfg, geom, pfn, P, mv, use_abs, with_dist = load(path)
    B, C, ny, nx = (int(v) for v in d["canvas_shape"])
    canvas = oracle.pointpillar_scatter(d["pillar_features"], d["voxel_coords"], B, C, ny, nx)
    nz = np.argwhere(np.abs(canvas).sum(axis=1) != 0).astype(np.int32)
    assert np.array_equal(nz, d["canvas_nonzero_byx"])
    chk = np.asarray([canvas.astype(np.float64).sum(), np.abs(canvas).astype(np.float64).sum()])
    assert np.array_equal(chk, d["canvas_checksum"])
    co = d["voxel_coords"]
    assert np.array_equal(canvas[co[:, 0], :, co[:, 2], co[:, 3]], d["pillar_features"])


@pytest.mark.parametrize("path", FIXTURES[:3], ids=[os.path.basename(p)[4:-4] for p in FIXTURES[:3]])
def test_fixture_voxels_come_from_the_oracle_voxelizer(path):
    """The fixtures' voxels/coords/num were produced by oracle.voxelize from the stored points: re-derive them."""
    d, cfg, geom, pfn, P, mv, use_abs, with_dist = load(path)
    offs = d["frame_offsets"]
    res = oracle.points_to_bev(d["points"], offs, geom, pfn, P, mv, F=cfg["F"], xcol=1, use_absolute_xyz=use_abs,
                               with_distance=with_dist, want_canvas=False)
    assert np.array_equal(res["voxel_coords"], d["voxel_coords"])
    assert np.array_equal(res["voxel_num_points"], d["voxel_num_points"])
    assert np.array_equal(res["voxels"], d["voxels"])


RADAR7 = sorted(glob.glob(os.path.join(HERE, "golden", "radar7_*.npz")))


def expand_radar7_weight(w, selected):
    """[C, len(sel)+6] -> [C, 13]: zero columns for the unselected raw features (modules.Radar7PillarVFE._expanded)."""
    full = np.zeros((w.shape[0], 13), dtype=np.float32)
    k = len(selected)
    full[:, selected] = w[:, :k]
    full[:, 7:13] = w[:, k:k + 6]
    return full


@pytest.mark.parametrize("path", RADAR7, ids=[os.path.basename(p)[:-4] for p in RADAR7])
def test_radar7_equals_zero_column_pillar_vfe(path):
    """Radar7PillarVFE (pillar_vfe.py:125-271) == PillarVFE arithmetic with zero weight columns, bit for bit."""
    d = np.load(path)
    cfg = synthetic.CONFIGS["vod"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    pfn = oracle.PfnParams(expand_radar7_weight(d["weight"], d["selected_indexes"]), d["gamma"], d["beta"],
                           d["running_mean"], d["running_var"])
    got = oracle.pillar_vfe(d["voxels_after"], d["voxel_coords"], d["voxel_num_points"], geom, pfn,
                            invstd_override=d["ref_invstd"])
    assert np.array_equal(got.view(np.uint32), d["pillar_features"].view(np.uint32))
    noz = any(str(f) == "USE_ELEVATION=0" for f in d["flags"])
    assert (d["voxels_after"][:, :, 2] == 0).all() == noz          # z zeroed in place only without elevation
