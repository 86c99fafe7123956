"""Generates tests/golden/pfnvar_*.npz by running the REFERENCE's own PillarVFE (imported by file path from /root/reference,
CPU, eval mode) on seeded inputs for the PFN variants beyond the shipped [64] layer: other widths (NUM_FILTERS [32], [128]),
stacked PFNs ([64, 64], [128, 128], [64, 32]: pillar_vfe.py:18-19,47-49,63-74) and MAX_POINTS_PER_VOXEL = 100.
Runs only in the build container; the fixtures are committed.

    python tests/golden/make_golden_pfn_variants.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from hgsfusion_b200 import synthetic  # noqa: E402
from oracle import oracle  # noqa: E402
from make_golden import load_reference  # noqa: E402

CASES = [
    # name, config, n points, P, NUM_FILTERS, use_abs, with_dist, dense spots
    ("c32",       "vod",  2500,  32, [32],       True,  False, 0),
    ("c128",      "tj4d", 2500,  10, [128],      True,  False, 0),
    ("p100",      "vod",  3000, 100, [64],       True,  False, 2),
    ("stack64",   "vod",  2500,  32, [64, 64],   True,  False, 0),
    ("stack128",  "tj4d", 2000,  10, [128, 128], True,  False, 0),
    ("stack6432", "vod",  2000,  20, [64, 32],   True,  True,  1),
]


def run(PillarVFE, name, config, n, P, filters, use_abs, with_dist, spots):
    cfg = synthetic.CONFIGS[config]
    F = cfg["F"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    rng = np.random.default_rng(len(name) * 7 + P)
    f = synthetic.make_frame(n, cfg["pc_range"], F, 77 + P, "clustered")
    for k in range(spots):                                   # very dense cells: pillars beyond P points
        c = rng.uniform([5, -15], [40, 15])
        f[k * 400:(k + 1) * 400, :2] = (c + rng.normal(0, 0.1, size=(400, 2))).astype(np.float32)
    vox, c3, num = oracle.voxelize(f, geom, P, 40000, F=F, xcol=0)
    coords = np.concatenate([np.zeros((c3.shape[0], 1), np.int32), c3], axis=1)
    model_cfg = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=with_dist, USE_ABSLOTE_XYZ=use_abs, NUM_FILTERS=list(filters))
    vfe = PillarVFE(model_cfg=model_cfg, num_point_features=F, voxel_size=list(cfg["voxel_size"]),
                    point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32))
    out = dict(voxels=vox, voxel_coords=coords, voxel_num_points=num,
               meta=np.asarray([config, str(P), str(int(use_abs)), str(int(with_dist)), ",".join(map(str, filters)), torch.__version__]))
    sd = vfe.state_dict()
    for li, layer in enumerate(vfe.pfn_layers):
        co, ci = layer.linear.weight.shape
        w = synthetic.make_pfn(ci, co, seed=100 * li + len(name))
        sd[f"pfn_layers.{li}.linear.weight"] = torch.from_numpy(w.weight)
        sd[f"pfn_layers.{li}.norm.weight"] = torch.from_numpy(w.gamma)
        sd[f"pfn_layers.{li}.norm.bias"] = torch.from_numpy(w.beta)
        sd[f"pfn_layers.{li}.norm.running_mean"] = torch.from_numpy(w.running_mean)
        sd[f"pfn_layers.{li}.norm.running_var"] = torch.from_numpy(w.running_var)
        for k, v in (("weight", w.weight), ("gamma", w.gamma), ("beta", w.beta), ("running_mean", w.running_mean),
                     ("running_var", w.running_var)):
            out[f"l{li}_{k}"] = v
    vfe.load_state_dict(sd)
    vfe.eval()
    bd = dict(voxels=torch.from_numpy(vox).float(), voxel_coords=torch.from_numpy(coords).float(),
              voxel_num_points=torch.from_numpy(num).float())
    with torch.no_grad():
        bd = vfe(bd)
    out["pillar_features"] = bd["pillar_features"].numpy()
    return out


def main():
    PillarVFE, _, _ = load_reference()
    torch.manual_seed(0)
    for case in CASES:
        data = run(PillarVFE, *case)
        path = os.path.join(HERE, f"pfnvar_{case[0]}.npz")
        np.savez_compressed(path, **data)
        print(f"{case[0]:10s} M={data['voxels'].shape[0]:5d} P={case[3]:3d} filters={case[4]} max cnt={int(data['voxel_num_points'].max())} "
              f"-> {os.path.basename(path)} ({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
