"""GPU parity for the PFN variants beyond the fused kernel's single 64-channel layer: widths 32 / 128, MAX_POINTS_PER_VOXEL 100 and
stacked (two-layer) PFNs, against the REFERENCE's own outputs (tests/golden/pfnvar_*.npz) -- through the C ABI contract entry
points and through the module mirrors (PillarVFE in contract mode, FusedPillarVFE from raw points)."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from hgsfusion_b200 import modules, synthetic
from hgsfusion_b200.ops import PfnWeights, PillarPath
from oracle import oracle
from test_oracle_pfn_variants import load_layers

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "pfnvar_*.npz")))


def dev_layer(p, dev, use_abs=True, with_dist=False):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return PfnWeights(weight=t(p.weight), bn_weight=t(p.gamma), bn_bias=t(p.beta), running_mean=t(p.running_mean),
                      running_var=t(p.running_var), use_absolute_xyz=use_abs, with_distance=with_dist)


def close(got, ref):
    tol = 1e-5 * np.abs(ref) + 1e-6 * np.abs(ref).max()        # north_star's 1e-5 relative; the absolute term covers outputs near 0
    return bool((np.abs(got - ref) <= tol).all())


@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[7:-4])
def test_contract_entry_points_match_reference(cuda, path):
    d = np.load(path)
    config, P, use_abs, with_dist = str(d["meta"][0]), int(d["meta"][1]), bool(int(d["meta"][2])), bool(int(d["meta"][3]))
    cfg = synthetic.CONFIGS[config]
    layers = load_layers(d)
    pp = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, 40000, cfg["F"])
    vox, co, nu = (torch.from_numpy(d[k]).to(cuda) for k in ("voxels", "voxel_coords", "voxel_num_points"))
    if len(layers) == 1:
        got = pp.pillar_vfe(vox, co.float(), nu.float(), dev_layer(layers[0], cuda, use_abs, with_dist))
    else:
        got = pp.pillar_vfe_stacked(vox, co, nu, dev_layer(layers[0], cuda, use_abs, with_dist), dev_layer(layers[1], cuda))
    assert close(got.cpu().numpy(), d["pillar_features"])
    if len(layers) == 1:
        # bit-identical to the C oracle for every width
        geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
        orc = oracle.pillar_vfe(d["voxels"], d["voxel_coords"], d["voxel_num_points"], geom, layers[0], use_abs, with_dist)
        assert np.array_equal(got.cpu().numpy().view(np.uint32), orc.view(np.uint32))


@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[7:-4])
def test_modules_accept_every_pfn_variant(cuda, path):
    """PillarVFE (contract mode) loads the reference's state_dict names for any NUM_FILTERS of one or two layers and reproduces
    the reference's pillar_features; the module no longer refuses stacked PFNs, other widths or P > 32."""
    d = np.load(path)
    config, P, use_abs, with_dist = str(d["meta"][0]), int(d["meta"][1]), bool(int(d["meta"][2])), bool(int(d["meta"][3]))
    filters = [int(v) for v in str(d["meta"][4]).split(",")]
    cfg = synthetic.CONFIGS[config]
    mc = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=with_dist, USE_ABSLOTE_XYZ=use_abs, NUM_FILTERS=filters)
    m = modules.PillarVFE(model_cfg=mc, num_point_features=cfg["F"], voxel_size=cfg["voxel_size"],
                          point_cloud_range=np.asarray(cfg["pc_range"], dtype=np.float32)).to(cuda).eval()
    sd = {}
    for li in range(len(filters)):
        sd[f"pfn_layers.{li}.linear.weight"] = torch.from_numpy(d[f"l{li}_weight"])
        sd[f"pfn_layers.{li}.norm.weight"] = torch.from_numpy(d[f"l{li}_gamma"])
        sd[f"pfn_layers.{li}.norm.bias"] = torch.from_numpy(d[f"l{li}_beta"])
        sd[f"pfn_layers.{li}.norm.running_mean"] = torch.from_numpy(d[f"l{li}_running_mean"])
        sd[f"pfn_layers.{li}.norm.running_var"] = torch.from_numpy(d[f"l{li}_running_var"])
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    bd = dict(voxels=torch.from_numpy(d["voxels"]).to(cuda), voxel_coords=torch.from_numpy(d["voxel_coords"]).float().to(cuda),
              voxel_num_points=torch.from_numpy(d["voxel_num_points"]).float().to(cuda))
    with torch.no_grad():
        out = m(bd)["pillar_features"]
    assert close(out.cpu().numpy(), d["pillar_features"])


@pytest.mark.parametrize("filters", [[32], [128], [64, 64], [128, 64]])
def test_fused_module_from_points_for_other_widths_and_stacks(cuda, filters):
    """FusedPillarVFE outside the fused kernel set (not 64 channels, or stacked): pillarize -> PFN -> scatter, all native; the
    canvas holds the pillar features at the pillar cells and zeros elsewhere, and the features equal the numpy oracle."""
    cfg = synthetic.CONFIGS["vod"]
    B, n, P = 2, 3000, 32
    pts, offs = synthetic.make_batch("vod", B, n, "clustered", seed0=9, oob_fraction=0.02)
    mc = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters, MAX_POINTS_PER_VOXEL=P,
                         MAX_NUMBER_OF_VOXELS=40000, RETURN_VOXELS=True)
    torch.manual_seed(3)
    m = modules.FusedPillarVFE(model_cfg=mc, num_point_features=7, voxel_size=cfg["voxel_size"],
                               point_cloud_range=np.asarray(cfg["pc_range"], dtype=np.float32)).to(cuda).eval()
    with torch.no_grad():
        for layer in m.pfn_layers:
            layer.norm.running_mean.normal_(); layer.norm.running_var.uniform_(0.5, 2.0)
            layer.norm.weight.uniform_(0.5, 1.5); layer.norm.bias.normal_(0, 0.5)
        out = m(dict(points=torch.from_numpy(pts).to(cuda), batch_size=B))
    assert not m.fused_ok
    Cout = filters[-1]
    feats, coords, canvas = out["pillar_features"], out["voxel_coords"], out["spatial_features"]
    assert tuple(canvas.shape) == (B, Cout, 320, 320) and feats.shape[1] == Cout
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    layers = [oracle.PfnParams(l.linear.weight.detach().cpu().numpy(), l.norm.weight.detach().cpu().numpy(), l.norm.bias.detach().cpu().numpy(),
                               l.norm.running_mean.cpu().numpy(), l.norm.running_var.cpu().numpy()) for l in m.pfn_layers]
    ref = oracle.pillar_vfe_layers(out["voxels"].cpu().numpy(), coords.cpu().numpy(), out["voxel_num_points"].cpu().numpy(), geom, layers)
    assert close(feats.cpu().numpy(), ref)
    idx = coords.long()
    assert torch.equal(canvas[idx[:, 0], :, idx[:, 2], idx[:, 3]], feats)
    assert int((canvas != 0).any(dim=1).sum()) <= feats.shape[0]
