"""GPU: training through the native modules (csrc/train_ops.cu) against the REFERENCE's own autograd results
(tests/golden/train_*.npz: PillarVFE in train mode + PointPillarScatter, torch CPU) and the numpy restatement.
Tolerances: forward 1e-5 (north_star's fp32 bound); gradients and running statistics 2e-5 of the largest entry
(fp32 reference autograd vs fp64-accumulated native sums)."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from hgsfusion_b200 import modules, synthetic
from oracle import oracle, train_oracle as to
from util import features_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "train_*.npz")))


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a.astype(np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())


def build(d, cuda, use_norm=True):
    cfgname, P, ua, wd, _ = d["meta"]
    cfg = synthetic.CONFIGS[str(cfgname)]
    model_cfg = SimpleNamespace(USE_NORM=use_norm, WITH_DISTANCE=bool(int(wd)), USE_ABSLOTE_XYZ=bool(int(ua)), NUM_FILTERS=[64])
    vfe = modules.PillarVFE(model_cfg=model_cfg, num_point_features=cfg["F"], voxel_size=list(cfg["voxel_size"]),
                            point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(cuda)
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    sc = modules.PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=np.asarray(geom.grid)).to(cuda)
    sd = vfe.state_dict()
    sd["pfn_layers.0.linear.weight"] = torch.from_numpy(d["weight"])
    if use_norm:
        sd["pfn_layers.0.norm.weight"] = torch.from_numpy(d["gamma"])
        sd["pfn_layers.0.norm.bias"] = torch.from_numpy(d["beta"])
        sd["pfn_layers.0.norm.running_mean"] = torch.from_numpy(d["running_mean"])
        sd["pfn_layers.0.norm.running_var"] = torch.from_numpy(d["running_var"])
    else:
        sd["pfn_layers.0.linear.bias"] = torch.from_numpy(d["beta"])
    vfe.load_state_dict(sd)
    bd = dict(voxels=torch.from_numpy(d["voxels"]).to(cuda), voxel_coords=torch.from_numpy(d["voxel_coords"]).float().to(cuda),
              voxel_num_points=torch.from_numpy(d["voxel_num_points"]).float().to(cuda))
    return vfe, sc, bd, cfg


def canvas_cotangent(d, shape, cuda):
    Rc = np.zeros(shape, dtype=np.float32)
    c, sel = d["voxel_coords"], d["grad_canvas_idx"]
    Rc[c[sel, 0], 8:24, c[sel, 2], c[sel, 3]] = d["grad_canvas_vals"]
    return torch.from_numpy(Rc).to(cuda)


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_train_step_matches_reference_autograd(cuda, path):
    d = np.load(path)
    vfe, sc, bd, cfg = build(d, cuda)
    vfe.train()
    bd = sc(vfe(bd))
    pf, canvas = bd["pillar_features"], bd["spatial_features"]
    assert pf.requires_grad and canvas.requires_grad
    assert features_close(pf.detach().cpu().numpy(), d["pillar_features"], rtol=1e-5)
    loss = (pf * torch.from_numpy(d["grad_out"]).to(cuda)).sum() + (canvas * canvas_cotangent(d, canvas.shape, cuda)).sum()
    loss.backward()
    lin, bn = vfe.pfn_layers[0].linear, vfe.pfn_layers[0].norm
    assert rel(lin.weight.grad, d["grad_weight"]) < 2e-5
    assert rel(bn.weight.grad, d["grad_gamma"]) < 2e-5 and rel(bn.bias.grad, d["grad_beta"]) < 2e-5
    assert rel(bn.running_mean, d["running_mean_after"]) < 2e-6 and rel(bn.running_var, d["running_var_after"]) < 2e-6
    assert int(bn.num_batches_tracked) == 1


def test_frozen_batchnorm_and_no_norm_gradients(cuda):
    d = np.load(FIXTURES[0])
    cfgname, P, ua, wd, _ = d["meta"]
    cfg = synthetic.CONFIGS[str(cfgname)]
    feats = to.decorate(d["voxels"], d["voxel_coords"], d["voxel_num_points"], cfg["pc_range"], cfg["voxel_size"]).astype(np.float64)
    g = d["grad_out"].astype(np.float64)
    # (1) eval mode with gradients enabled: BatchNorm on the running statistics, which are constants in the backward
    vfe, sc, bd, _ = build(d, cuda)
    vfe.eval()
    pf = vfe(bd)["pillar_features"]
    (pf * torch.from_numpy(d["grad_out"]).to(cuda)).sum().backward()
    x = feats @ d["weight"].astype(np.float64).T
    invstd = 1 / np.sqrt(d["running_var"].astype(np.float64) + 1e-3)
    xhat = (x - d["running_mean"]) * invstd
    y = xhat * d["gamma"] + d["beta"]
    z = np.maximum(y, 0)
    arg = z.argmax(axis=1)
    dz = np.zeros_like(y)
    np.put_along_axis(dz, arg[:, None, :], g[:, None, :], axis=1)
    dy = dz * (y > 0)
    lin, bn = vfe.pfn_layers[0].linear, vfe.pfn_layers[0].norm
    assert rel(lin.weight.grad, np.einsum("mpc,mpk->ck", dy * d["gamma"] * invstd, feats)) < 2e-5
    assert rel(bn.weight.grad, (dy * xhat).sum((0, 1))) < 2e-5 and rel(bn.bias.grad, dy.sum((0, 1))) < 2e-5
    assert rel(bn.running_mean, d["running_mean"]) == 0                      # untouched in eval mode
    # (2) USE_NORM False: Linear with bias, ReLU, max
    vfe, sc, bd, _ = build(d, cuda, use_norm=False)
    vfe.train()
    pf = vfe(bd)["pillar_features"]
    (pf * torch.from_numpy(d["grad_out"]).to(cuda)).sum().backward()
    y = feats @ d["weight"].astype(np.float64).T + d["beta"]
    z = np.maximum(y, 0)
    arg = z.argmax(axis=1)
    assert features_close(pf.detach().cpu().numpy(), np.take_along_axis(z, arg[:, None, :], axis=1)[:, 0, :].astype(np.float32), rtol=1e-5)
    dz = np.zeros_like(y)
    np.put_along_axis(dz, arg[:, None, :], g[:, None, :], axis=1)
    dy = dz * (y > 0)
    lin = vfe.pfn_layers[0].linear
    assert rel(lin.weight.grad, np.einsum("mpc,mpk->ck", dy, feats)) < 2e-5 and rel(lin.bias.grad, dy.sum((0, 1))) < 2e-5


def test_fused_module_trains_from_points(cuda):
    """FusedPillarVFE in train mode (points in) gives the gradients of PillarVFE + PointPillarScatter on the same pillars."""
    cfg = synthetic.CONFIGS["vod"]
    pts, offs = synthetic.make_batch("vod", 2, 1500, "clustered", seed0=5)
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    P = 10
    fused_cfg = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64],
                                MAX_POINTS_PER_VOXEL=P, MAX_NUMBER_OF_VOXELS={'train': 16000, 'test': 40000})
    fused = modules.FusedPillarVFE(model_cfg=fused_cfg, num_point_features=7, voxel_size=list(cfg["voxel_size"]),
                                   point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(cuda)
    w = synthetic.make_pfn(13, 64, seed=3)
    state = {"pfn_layers.0.linear.weight": torch.from_numpy(w.weight), "pfn_layers.0.norm.weight": torch.from_numpy(w.gamma),
             "pfn_layers.0.norm.bias": torch.from_numpy(w.beta), "pfn_layers.0.norm.running_mean": torch.from_numpy(w.running_mean),
             "pfn_layers.0.norm.running_var": torch.from_numpy(w.running_var)}
    fused.load_state_dict(state, strict=False)
    fused.train()
    bd = fused(dict(points=torch.from_numpy(pts).to(cuda), batch_size=2))
    M = bd["pillar_features"].shape[0]
    rng = np.random.default_rng(0)
    R = rng.standard_normal((M, 64)).astype(np.float32)
    Rc = rng.standard_normal(tuple(bd["spatial_features"].shape)).astype(np.float32)
    ((bd["pillar_features"] * torch.from_numpy(R).to(cuda)).sum() + (bd["spatial_features"] * torch.from_numpy(Rc).to(cuda)).sum()).backward()
    # the numpy restatement on the oracle's pillars
    vox, coords, num = [], [], []
    for b in range(2):
        v, c, k = oracle.voxelize(pts[offs[b]:offs[b + 1]], geom, P, 16000, F=7, xcol=1)
        vox.append(v); num.append(k)
        coords.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], axis=1))
    vox, coords, num = np.concatenate(vox), np.concatenate(coords), np.concatenate(num)
    assert M == vox.shape[0] and np.array_equal(bd["voxel_coords"].cpu().numpy(), coords)
    feats = to.decorate(vox, coords, num, cfg["pc_range"], cfg["voxel_size"])
    out, cache = to.pfn_train_forward(feats, w.weight, w.gamma, w.beta)
    assert features_close(bd["pillar_features"].detach().cpu().numpy(), out.astype(np.float32), rtol=1e-5)
    g = R.astype(np.float64) + to.scatter_backward(Rc, coords)
    dW, dg, db = to.pfn_backward(cache, w.gamma, g)
    lin, bn = fused.pfn_layers[0].linear, fused.pfn_layers[0].norm
    assert rel(lin.weight.grad, dW) < 2e-5 and rel(bn.weight.grad, dg) < 2e-5 and rel(bn.bias.grad, db) < 2e-5
    # and the eval path of the same module is the fused kernel again
    fused.eval()
    with torch.no_grad():
        ev = fused(dict(points=torch.from_numpy(pts).to(cuda), batch_size=2))
    assert not ev["pillar_features"].requires_grad and ev["spatial_features"].shape == bd["spatial_features"].shape


def test_radar7_trains_only_the_selected_columns(cuda):
    d = np.load(os.path.join(ROOT, "tests", "golden", "radar7_subset_noz.npz"))
    flags = {k: bool(int(v)) for k, v in (f.split("=") for f in d["flags"])}
    cfg = synthetic.CONFIGS["vod"]
    m = modules.Radar7PillarVFE(model_cfg=SimpleNamespace(USE_NORM=True, NUM_FILTERS=[64], **flags), num_point_features=7,
                                voxel_size=list(cfg["voxel_size"]), point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(cuda)
    m.train()
    bd = dict(voxels=torch.from_numpy(d["voxels"].copy()).to(cuda), voxel_coords=torch.from_numpy(d["voxel_coords"]).float().to(cuda),
              voxel_num_points=torch.from_numpy(d["voxel_num_points"]).float().to(cuda))
    pf = m(bd)["pillar_features"]
    pf.square().sum().backward()
    gw = m.pfn_layers[0].linear.weight.grad
    assert gw.shape == m.pfn_layers[0].linear.weight.shape and torch.isfinite(gw).all() and gw.abs().sum() > 0


@pytest.mark.parametrize("cfgname,npts,P", [("vod", 1500, 10), ("tj4d", 2500, 32), ("vod", 40000, 5)])
def test_fused_train_path_three_launches_and_no_readback(cuda, cfgname, npts, P):
    """Train mode inside the fused kernel's domain: forward = k_front + statistics pass + fused pass (3 launches), backward 3,
    capacity-sized outputs with the count on the device; same results as the contract-layout train path (pillarize ->
    batch statistics -> PFN -> scatter) on the same points."""
    cfg = synthetic.CONFIGS[cfgname]
    B = 2
    pts, offs = synthetic.make_batch(cfgname, B, npts, "clustered", seed0=11)
    F = cfg["F"]
    w = synthetic.make_pfn(F + 6, 64, seed=5)
    state = {"pfn_layers.0.linear.weight": torch.from_numpy(w.weight), "pfn_layers.0.norm.weight": torch.from_numpy(w.gamma),
             "pfn_layers.0.norm.bias": torch.from_numpy(w.beta), "pfn_layers.0.norm.running_mean": torch.from_numpy(w.running_mean),
             "pfn_layers.0.norm.running_var": torch.from_numpy(w.running_var)}

    def make(trim):
        mc = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64], MAX_POINTS_PER_VOXEL=P,
                             MAX_NUMBER_OF_VOXELS={'train': 16000, 'test': 40000}, TRIM=trim, RETURN_VOXELS=True)
        m = modules.FusedPillarVFE(model_cfg=mc, num_point_features=F, voxel_size=list(cfg["voxel_size"]),
                                   point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(cuda)
        m.load_state_dict(state, strict=False)
        return m.train()

    dpts = torch.from_numpy(pts).to(cuda)
    fused, ref = make(False), make(True)
    bd = fused(dict(points=dpts, batch_size=B))
    assert fused._path().last_launches == 3
    out_ref = ref._forward_train(dpts, B)                       # the contract-layout chain (trimmed, one sync)
    counts = bd["num_pillars"].cpu().numpy()
    M = int(counts[0])
    assert M == out_ref["pillar_features"].shape[0] and bd["pillar_features"].shape[0] >= M
    assert torch.equal(bd["voxel_coords"][:M], out_ref["voxel_coords"]) and torch.equal(bd["voxels"][:M], out_ref["voxels"])
    assert features_close(bd["pillar_features"][:M].detach().cpu().numpy(), out_ref["pillar_features"].detach().cpu().numpy(), rtol=1e-5)
    assert features_close(bd["spatial_features"].detach().cpu().numpy(), out_ref["spatial_features"].detach().cpu().numpy(), rtol=1e-5)
    bn_f, bn_r = fused.pfn_layers[0].norm, ref.pfn_layers[0].norm
    ref.pfn_layers[0].norm.num_batches_tracked += 0
    assert rel(bn_f.running_mean, bn_r.running_mean.cpu().numpy()) < 2e-6 and rel(bn_f.running_var, bn_r.running_var.cpu().numpy()) < 2e-6
    assert int(bn_f.num_batches_tracked) == 1
    rng = np.random.default_rng(1)
    R = torch.from_numpy(rng.standard_normal((M, 64)).astype(np.float32)).to(cuda)
    Rc = torch.from_numpy(rng.standard_normal(tuple(bd["spatial_features"].shape)).astype(np.float32)).to(cuda)
    ((bd["pillar_features"][:M] * R).sum() + (bd["spatial_features"] * Rc).sum()).backward()
    assert fused._path().last_launches == 3
    ((out_ref["pillar_features"] * R).sum() + (out_ref["spatial_features"] * Rc).sum()).backward()
    for a, b in ((fused.pfn_layers[0].linear.weight, ref.pfn_layers[0].linear.weight), (bn_f.weight, bn_r.weight), (bn_f.bias, bn_r.bias)):
        assert rel(a.grad, b.grad.cpu().numpy()) < 2e-5
    # only a canvas cotangent (what a detector's loss gives): the pillar_features one is absent
    fused.zero_grad(); ref.zero_grad()
    bd = fused(dict(points=dpts, batch_size=B))
    (bd["spatial_features"] * Rc).sum().backward()
    out_ref = ref._forward_train(dpts, B)
    (out_ref["spatial_features"] * Rc).sum().backward()
    assert rel(fused.pfn_layers[0].linear.weight.grad, ref.pfn_layers[0].linear.weight.grad.cpu().numpy()) < 2e-5
