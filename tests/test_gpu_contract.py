"""GPU parity of the batch_dict contract ops (PillarVFE, PointPillarScatter, pillarize) and of the
drop-in nn.Modules, against the committed golden fixtures of the REFERENCE and against the oracle."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from hgsfusion_b200 import modules, synthetic
from hgsfusion_b200.ops import PillarPath
from oracle import oracle
from util import bits_equal, device_pfn, geom_for, features_close, oracle_pfn

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "vfe_*.npz")))
RTOL = 1e-5


def load(path):
    d = np.load(path)
    config, P, mv, use_abs, with_dist = d["meta"][:5]
    return d, str(config), int(P), int(mv), bool(int(use_abs)), bool(int(with_dist))


def fixture_pfn(d):
    return SimpleNamespace(weight=d["weight"], gamma=d["gamma"], beta=d["beta"], running_mean=d["running_mean"],
                           running_var=d["running_var"], bias=None, eps=1e-3)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[4:-4] for p in FIXTURES])
@pytest.mark.parametrize("as_float", [True, False], ids=["f32coords", "i32coords"])
def test_pillar_vfe_against_reference_fixture(cuda, path, as_float):
    d, config, P, mv, use_abs, with_dist = load(path)
    cfg = synthetic.CONFIGS[config]
    path_ = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, mv, cfg["F"])
    vox = torch.from_numpy(d["voxels"]).to(cuda)
    co = torch.from_numpy(d["voxel_coords"]).to(cuda)
    nu = torch.from_numpy(d["voxel_num_points"]).to(cuda)
    if as_float:                       # load_data_to_gpu: everything arrives as float32 (pcdet/models/__init__.py:36)
        co, nu = co.float(), nu.float()
    got = path_.pillar_vfe(vox, co, nu, device_pfn(fixture_pfn(d), cuda, use_abs, with_dist)).cpu().numpy()
    ref = d["pillar_features"]
    assert features_close(got, ref, RTOL)
    # bit-identical to the oracle (which equals the reference except where torch's MKL sqrt is 1 ulp off)
    geom = geom_for(config)
    orc = oracle.pillar_vfe(d["voxels"], d["voxel_coords"], d["voxel_num_points"], geom, oracle_pfn(fixture_pfn(d)),
                            use_absolute_xyz=use_abs, with_distance=with_dist)
    assert bits_equal(got, orc)


@pytest.mark.parametrize("path", FIXTURES[:4], ids=[os.path.basename(p)[4:-4] for p in FIXTURES[:4]])
def test_scatter_against_reference_fixture(cuda, path):
    d, config, P, mv, use_abs, with_dist = load(path)
    cfg = synthetic.CONFIGS[config]
    B, C, ny, nx = (int(v) for v in d["canvas_shape"])
    path_ = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, mv, cfg["F"])
    feats = torch.from_numpy(d["pillar_features"]).to(cuda)
    for co in (torch.from_numpy(d["voxel_coords"]).to(cuda), torch.from_numpy(d["voxel_coords"]).to(cuda).float()):
        canvas = path_.pointpillar_scatter(feats, co, B).cpu().numpy()
        assert canvas.shape == (B, C, ny, nx)
        ref = oracle.pointpillar_scatter(d["pillar_features"], d["voxel_coords"], B, C, ny, nx)
        assert bits_equal(canvas, ref)
        nz = np.argwhere(np.abs(canvas).sum(axis=1) != 0).astype(np.int32)
        assert np.array_equal(nz, d["canvas_nonzero_byx"])
        chk = np.asarray([canvas.astype(np.float64).sum(), np.abs(canvas).astype(np.float64).sum()])
        assert np.array_equal(chk, d["canvas_checksum"])


def test_scatter_duplicates_last_row_wins_and_bad_batch_ignored(cuda):
    path_ = PillarPath(np.asarray([0, 0, 0, 8, 4, 1], dtype=np.float32), [1, 1, 1], 1, 1, 4)
    feats = torch.arange(5 * 64, dtype=torch.float32, device=cuda).view(5, 64) + 1
    coords = torch.tensor([[0, 0, 1, 2], [0, 0, 1, 2], [1, 0, 3, 7], [5, 0, 0, 0], [-1, 0, 0, 0]], dtype=torch.int32, device=cuda)
    canvas = path_.pointpillar_scatter(feats, coords, 2).cpu().numpy()
    assert np.array_equal(canvas[0, :, 1, 2], feats[1].cpu().numpy())
    assert np.array_equal(canvas[1, :, 3, 7], feats[2].cpu().numpy())
    assert np.count_nonzero(canvas) == 2 * 64


@pytest.mark.parametrize("config,B,n,P,mv", [("vod", 2, 2500, 32, 40000), ("tj4d", 2, 2500, 10, 900), ("stress", 1, 6000, 5, 40000)])
def test_pillarize_only(cuda, config, B, n, P, mv):
    """transform_points_to_voxels + collate on the device: voxels, voxel_coords, voxel_num_points."""
    cfg = synthetic.CONFIGS[config]
    F = cfg["F"]
    pts, offs = synthetic.make_batch(config, B, n, "clustered", seed0=9, oob_fraction=0.03)
    geom = geom_for(config)
    path_ = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, mv, F)
    res = path_.pillarize(torch.from_numpy(pts).to(cuda), B).trim()
    m0 = 0
    for b in range(B):
        v, c, k = oracle.voxelize(pts[offs[b]:offs[b + 1]], geom, P, mv, F=F, xcol=1)
        m = c.shape[0]
        assert np.array_equal(res["voxel_coords"][m0:m0 + m].cpu().numpy()[:, 1:], c)
        assert (res["voxel_coords"][m0:m0 + m, 0] == b).all()
        assert np.array_equal(res["voxel_num_points"][m0:m0 + m].cpu().numpy(), k)
        assert bits_equal(res["voxels"][m0:m0 + m].cpu().numpy(), v)
        m0 += m
    assert m0 == res["num_pillars"]


def _ns(**kw):
    base = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64])
    base.update(kw)
    return SimpleNamespace(**base)


def _load_weights(m, w):
    sd = m.state_dict()
    sd["pfn_layers.0.linear.weight"] = torch.from_numpy(w.weight)
    sd["pfn_layers.0.norm.weight"] = torch.from_numpy(w.gamma)
    sd["pfn_layers.0.norm.bias"] = torch.from_numpy(w.beta)
    sd["pfn_layers.0.norm.running_mean"] = torch.from_numpy(w.running_mean)
    sd["pfn_layers.0.norm.running_var"] = torch.from_numpy(w.running_var)
    m.load_state_dict(sd)


def test_modules_reproduce_the_batch_dict_contract(cuda):
    """PillarVFE -> PointPillarScatter through the modules == FusedPillarVFE -> passthrough == oracle."""
    config, B, n, P, mv = "vod", 3, 2000, 32, 40000
    cfg = synthetic.CONFIGS[config]
    rng = np.asarray(cfg["pc_range"], dtype=np.float32)
    pts, offs = synthetic.make_batch(config, B, n, "clustered", seed0=4, oob_fraction=0.02)
    w = synthetic.make_pfn(13, 64, 2)
    ref = oracle.points_to_bev(pts, offs, geom_for(config), oracle_pfn(w), P, mv, F=7, xcol=1)

    vfe = modules.PillarVFE(model_cfg=_ns(), num_point_features=7, voxel_size=cfg["voxel_size"], point_cloud_range=rng)
    _load_weights(vfe, w)
    vfe = vfe.to(cuda).eval()
    scatter = modules.PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=[320, 320, 1]).to(cuda).eval()
    bd = dict(voxels=torch.from_numpy(ref["voxels"]).to(cuda), voxel_coords=torch.from_numpy(ref["voxel_coords"]).to(cuda).float(),
              voxel_num_points=torch.from_numpy(ref["voxel_num_points"]).to(cuda).float(), batch_size=B)
    with torch.no_grad():
        bd = scatter(vfe(bd))
    assert bits_equal(bd["pillar_features"].cpu().numpy(), ref["pillar_features"])
    assert bits_equal(bd["spatial_features"].cpu().numpy(), ref["spatial_features"])

    fused = modules.FusedPillarVFE(model_cfg=_ns(MAX_POINTS_PER_VOXEL=P, MAX_NUMBER_OF_VOXELS={'train': 16000, 'test': mv},
                                                 RETURN_VOXELS=True),
                                   num_point_features=7, voxel_size=cfg["voxel_size"], point_cloud_range=rng)
    _load_weights(fused, w)
    fused = fused.to(cuda).eval()
    passthrough = modules.PillarScatterPassthrough(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=[320, 320, 1])
    with torch.no_grad():
        bd2 = passthrough(fused(dict(points=torch.from_numpy(pts).to(cuda), batch_size=B)))
    assert np.array_equal(bd2["voxel_coords"].cpu().numpy(), ref["voxel_coords"])
    assert np.array_equal(bd2["voxel_num_points"].cpu().numpy(), ref["voxel_num_points"])
    assert bits_equal(bd2["voxels"].cpu().numpy(), ref["voxels"])
    assert bits_equal(bd2["pillar_features"].cpu().numpy(), ref["pillar_features"])
    assert bits_equal(bd2["spatial_features"].cpu().numpy(), ref["spatial_features"])


def test_use_norm_false_and_relative_xyz_and_distance(cuda):
    config, B, n, P, mv = "tj4d", 2, 1500, 16, 40000
    cfg = synthetic.CONFIGS[config]
    F = cfg["F"]
    pts, offs = synthetic.make_batch(config, B, n, "clustered", seed0=8)
    for use_abs, with_dist, use_norm in ((False, False, True), (True, True, True), (True, False, False)):
        Cin = (F if use_abs else F - 3) + 6 + (1 if with_dist else 0)
        w = synthetic.make_pfn(Cin, 64, 5)
        if not use_norm:
            w = SimpleNamespace(weight=w.weight, gamma=None, beta=None, running_mean=None, running_var=None,
                                bias=w.beta, eps=1e-3)
        ref = oracle.points_to_bev(pts, offs, geom_for(config), oracle_pfn(w), P, mv, F=F, xcol=1,
                                   use_absolute_xyz=use_abs, with_distance=with_dist)
        path_ = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, mv, F)
        got = path_.points_to_bev(torch.from_numpy(pts).to(cuda), B, device_pfn(w, cuda, use_abs, with_dist)).trim()
        assert np.array_equal(got["voxel_coords"].cpu().numpy(), ref["voxel_coords"])
        assert bits_equal(got["pillar_features"].cpu().numpy(), ref["pillar_features"]), (use_abs, with_dist, use_norm)
        assert bits_equal(got["spatial_features"].cpu().numpy(), ref["spatial_features"])


RADAR7 = sorted(glob.glob(os.path.join(HERE, "golden", "radar7_*.npz")))


@pytest.mark.parametrize("path", RADAR7, ids=[os.path.basename(p)[:-4] for p in RADAR7])
def test_radar7_module_against_reference_fixture(cuda, path):
    d = np.load(path)
    cfg = synthetic.CONFIGS["vod"]
    flags = {str(f).split("=")[0]: bool(int(str(f).split("=")[1])) for f in d["flags"]}
    m = modules.Radar7PillarVFE(model_cfg=SimpleNamespace(USE_NORM=True, NUM_FILTERS=[64], **flags), num_point_features=7,
                                voxel_size=cfg["voxel_size"], point_cloud_range=np.asarray(cfg["pc_range"], dtype=np.float32))
    assert m.selected_indexes.tolist() == d["selected_indexes"].tolist()
    _load_weights(m, SimpleNamespace(weight=d["weight"], gamma=d["gamma"], beta=d["beta"], running_mean=d["running_mean"],
                                     running_var=d["running_var"]))
    m = m.to(cuda).eval()
    bd = dict(voxels=torch.from_numpy(d["voxels"].copy()).to(cuda), voxel_coords=torch.from_numpy(d["voxel_coords"]).to(cuda).float(),
              voxel_num_points=torch.from_numpy(d["voxel_num_points"]).to(cuda).float())
    with torch.no_grad():
        bd = m(bd)
    assert features_close(bd["pillar_features"].cpu().numpy(), d["pillar_features"], RTOL)
    assert np.array_equal(bd["voxels"].cpu().numpy(), d["voxels_after"])        # same in-place side effect on z
