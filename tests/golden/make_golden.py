"""Generates tests/golden/vfe_*.npz by running the REFERENCE's own PillarVFE and
PointPillarScatter (imported by file path from /root/reference, CPU, eval mode) on seeded
synthetic inputs.  Runs only in the build container (the reference mount does not exist on
the GPU box); the .npz fixtures it writes are committed.

    python tests/golden/make_golden.py            # rewrite fixtures
    python tests/golden/make_golden.py --check    # also compare the C oracle against them
    python tests/golden/make_golden.py --only-split | --only-train   # just the Path B input-prep / the training fixtures

The voxelizer that feeds these fixtures is oracle.voxelize (spconv is not available, see
oracle/pillar_oracle.c header); the fixtures pin PillarVFE + PointPillarScatter, whose inputs
(voxels, voxel_coords, voxel_num_points) are stored alongside the outputs.
"""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from hgsfusion_b200 import synthetic  # noqa: E402
from oracle import oracle  # noqa: E402


def load_reference():
    """pillar_vfe.py has one relative import (.vfe_template); give it a synthetic parent package."""
    pkg = types.ModuleType("refvfe")
    pkg.__path__ = []
    sys.modules["refvfe"] = pkg

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("refvfe.vfe_template", f"{REF}/pcdet/models/backbones_3d/vfe/vfe_template.py")
    vfe = load("refvfe.pillar_vfe", f"{REF}/pcdet/models/backbones_3d/vfe/pillar_vfe.py")
    sc = load("refscatter", f"{REF}/pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py")
    return vfe.PillarVFE, sc.PointPillarScatter, vfe.Radar7PillarVFE


CASES = [
    # name, config, B, n/frame, P, max_voxels, mode, use_abs, with_dist
    ("vod_p32",        "vod",    2, 3000, 32, 40000, "clustered", True,  False),
    ("vod_p10",        "vod",    2, 3000, 10, 40000, "clustered", True,  False),
    ("vod_p5_trunc",   "vod",    1, 6000,  5,  1500, "clustered", True,  False),
    ("vod_uniform",    "vod",    2, 2000, 32, 40000, "uniform",   True,  False),
    ("tj4d_p32",       "tj4d",   2, 3000, 32, 40000, "clustered", True,  False),
    ("stress_p10",     "stress", 1, 8000, 10, 40000, "clustered", True,  False),
    ("vod_relxyz",     "vod",    1, 2000, 32, 40000, "clustered", False, False),
    ("vod_dist",       "vod",    1, 2000, 32, 40000, "clustered", True,  True),
]


def run_case(PillarVFE, PointPillarScatter, name, config, B, n, P, max_voxels, mode, use_abs, with_dist):
    cfg = synthetic.CONFIGS[config]
    F = cfg["F"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    grid = geom.grid
    pts, offs = synthetic.make_batch(config, B, n, mode, seed0=11, oob_fraction=0.02)
    vox, coords, num = [], [], []
    for b in range(B):
        v, c, k = oracle.voxelize(pts[offs[b]:offs[b + 1]], geom, P, max_voxels, F=F, xcol=1)
        vox.append(v)
        coords.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], axis=1))
        num.append(k)
    vox, coords, num = np.concatenate(vox), np.concatenate(coords), np.concatenate(num)

    Cin = (F if use_abs else F - 3) + 6 + (1 if with_dist else 0)
    w = synthetic.make_pfn(Cin, 64, seed=len(name))
    model_cfg = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=with_dist, USE_ABSLOTE_XYZ=use_abs, NUM_FILTERS=[64])
    # the reference is constructed exactly as detector3d_template.py:93-108 does:
    # voxel_size is the YAML list of python floats, point_cloud_range an np.float32 array
    vfe = PillarVFE(model_cfg=model_cfg, num_point_features=F, voxel_size=list(cfg["voxel_size"]),
                    point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32))
    sd = vfe.state_dict()
    sd["pfn_layers.0.linear.weight"] = torch.from_numpy(w.weight)
    sd["pfn_layers.0.norm.weight"] = torch.from_numpy(w.gamma)
    sd["pfn_layers.0.norm.bias"] = torch.from_numpy(w.beta)
    sd["pfn_layers.0.norm.running_mean"] = torch.from_numpy(w.running_mean)
    sd["pfn_layers.0.norm.running_var"] = torch.from_numpy(w.running_var)
    vfe.load_state_dict(sd)
    vfe.eval()
    scatter = PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=grid)
    # load_data_to_gpu converts everything to float32 (pcdet/models/__init__.py:36)
    bd = dict(voxels=torch.from_numpy(vox).float(), voxel_coords=torch.from_numpy(coords).float(),
              voxel_num_points=torch.from_numpy(num).float())
    with torch.no_grad():
        bd = vfe(bd)
        bd = scatter(bd)
    feats = bd["pillar_features"].numpy()
    canvas = bd["spatial_features"].numpy()
    # the canvas is pillar_features scattered (exact copy); store it sparsely: occupied cells only
    nz = np.argwhere(np.abs(canvas).sum(axis=1) != 0)          # (b, y, x)
    return dict(points=pts, frame_offsets=offs, voxels=vox, voxel_coords=coords, voxel_num_points=num,
                weight=w.weight, gamma=w.gamma, beta=w.beta, running_mean=w.running_mean,
                running_var=w.running_var, pillar_features=feats,
                # torch's own 1/sqrt(var+eps) (MKL VML sqrt: not always correctly rounded)
                ref_invstd=(1 / torch.sqrt(torch.from_numpy(w.running_var) + 1e-3)).numpy(),
                canvas_shape=np.asarray(canvas.shape), canvas_nonzero_byx=nz.astype(np.int32),
                canvas_checksum=np.asarray([canvas.astype(np.float64).sum(), np.abs(canvas).astype(np.float64).sum()]),
                meta=np.asarray([config, str(P), str(max_voxels), str(int(use_abs)), str(int(with_dist)),
                                 torch.__version__, np.__version__])), canvas


RADAR7_CASES = [
    # name, flags (USE_XYZ, USE_RCS, USE_VR, USE_VR_COMP, USE_TIME, USE_ELEVATION, USE_DISTANCE), P
    ("radar7_all",        dict(USE_XYZ=True, USE_RCS=True, USE_VR=True, USE_VR_COMP=True, USE_TIME=True, USE_ELEVATION=True, USE_DISTANCE=False), 10),
    ("radar7_subset_noz", dict(USE_XYZ=True, USE_RCS=True, USE_VR=False, USE_VR_COMP=True, USE_TIME=False, USE_ELEVATION=False, USE_DISTANCE=False), 10),
    ("radar7_p32_sparse", dict(USE_XYZ=True, USE_RCS=False, USE_VR=False, USE_VR_COMP=True, USE_TIME=True, USE_ELEVATION=True, USE_DISTANCE=False), 32),
    # USE_DISTANCE=True cannot be pinned: the reference itself crashes (it appends the range feature without
    # counting it in the Linear's input width, pillar_vfe.py:133,250-252)
]


def run_radar7(Radar7PillarVFE, name, flags, P):
    """Radar7PillarVFE (pillar_vfe.py:125-271): feature selection by flags, optional in-place zeroing of z."""
    import contextlib, io
    cfg = synthetic.CONFIGS["vod"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    pts, offs = synthetic.make_batch("vod", 1, 2500, "clustered", seed0=31, oob_fraction=0.02)
    vox, c3, num = oracle.voxelize(pts, geom, P, 40000, F=7, xcol=1)
    coords = np.concatenate([np.zeros((c3.shape[0], 1), np.int32), c3], axis=1)
    model_cfg = SimpleNamespace(USE_NORM=True, NUM_FILTERS=[64], **flags)
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints its feature list
        vfe = Radar7PillarVFE(model_cfg=model_cfg, num_point_features=7, voxel_size=list(cfg["voxel_size"]),
                              point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32))
    Cin = vfe.pfn_layers[0].linear.weight.shape[1]
    w = synthetic.make_pfn(Cin, 64, seed=len(name))
    sd = vfe.state_dict()
    sd["pfn_layers.0.linear.weight"] = torch.from_numpy(w.weight)
    sd["pfn_layers.0.norm.weight"] = torch.from_numpy(w.gamma)
    sd["pfn_layers.0.norm.bias"] = torch.from_numpy(w.beta)
    sd["pfn_layers.0.norm.running_mean"] = torch.from_numpy(w.running_mean)
    sd["pfn_layers.0.norm.running_var"] = torch.from_numpy(w.running_var)
    vfe.load_state_dict(sd)
    vfe.eval()
    bd = dict(voxels=torch.from_numpy(vox.copy()).float(), voxel_coords=torch.from_numpy(coords).float(),
              voxel_num_points=torch.from_numpy(num).float())
    with torch.no_grad():
        bd = vfe(bd)
    return dict(voxels=vox, voxel_coords=coords, voxel_num_points=num, weight=w.weight, gamma=w.gamma, beta=w.beta,
                running_mean=w.running_mean, running_var=w.running_var,
                ref_invstd=(1 / torch.sqrt(torch.from_numpy(w.running_var) + 1e-3)).numpy(),
                selected_indexes=vfe.selected_indexes.numpy().astype(np.int32),
                pillar_features=bd["pillar_features"].numpy(),
                voxels_after=bd["voxels"].numpy(),            # the reference zeroes z in place when USE_ELEVATION is False
                flags=np.asarray([f"{k}={int(v)}" for k, v in flags.items()]), P=np.asarray(P))


SPLIT_CASES = [
    # name, dataset, Fin, num_input (READER.NUM_INPUT_FEATURES, hgsfusion_vod.yaml:107 / hgsfusion_tj4d.yaml:103), virtual, encoding, frame order
    ("split_vod",          "vod",  17, 29, True,  "split",  "grouped"),
    ("split_tj4d",         "tj4d", 18, 31, True,  "split",  "grouped"),
    ("split_vod_unsorted", "vod",  17, 29, True,  "split",  "shuffled"),
    ("split_vod_mixed",    "vod",  17, 17, True,  "mixed",  "grouped"),
    ("split_vod_direct",   "vod",  17, 15, True,  "direct", "grouped"),
    ("split_plain7",       "vod",   7,  7, False, "split",  "grouped"),
]


def make_hybrid_points(dataset, Fin, B, n, seed, order):
    """Hybrid radar rows as vod_dataset.py:498-530 lays them out: xyz, features, two flag columns
    (1,1 raw radar point; 0,0 radar point inside a mask; 0,1 virtual point), frame index in column 0."""
    rng = np.random.default_rng(seed)
    cfg = synthetic.CONFIGS[dataset]
    lo, hi = np.asarray(cfg["pc_range"][:3]), np.asarray(cfg["pc_range"][3:])
    rows = []
    for b in range(B):
        nb = n + 37 * b
        p = rng.standard_normal((nb, 1 + Fin)).astype(np.float32)
        p[:, 0] = b
        p[:, 1:4] = (lo + (hi - lo) * rng.random((nb, 3)) * 1.04 - 0.02 * (hi - lo)).astype(np.float32)
        if Fin >= 5:
            kind = rng.integers(0, 3, nb)
            p[:, -2] = (kind == 0)
            p[:, -1] = (kind != 1)
        rows.append(p)
    pts = np.concatenate(rows)
    if order == "shuffled":
        pts = pts[rng.permutation(pts.shape[0])]
    return pts


def run_split(name, dataset, Fin, num_input, virtual, encoding, order):
    """Runs the reference's DynamicPillarFeatureNet.forward (dynamic_pillar_encoder.py:55-121) unmodified; the PillarMaxPooling
    it calls at :120 (needs spconv) is replaced by a recorder, so what is pinned is exactly the reader's input."""
    seen = {}

    class Recorder(torch.nn.Module):
        def __init__(self, **kw):
            super().__init__()

        def forward(self, xyz, xyz_batch_cnt, pt_feature):
            seen.update(xyz=xyz.numpy().copy(), cnt=xyz_batch_cnt.numpy().copy(), feat=pt_feature.numpy().copy())
            return None

    for modname in ("pcdet", "pcdet.ops", "pcdet.ops.pillar_ops"):
        sys.modules.setdefault(modname, types.ModuleType(modname))
    stub = types.ModuleType("pcdet.ops.pillar_ops.pillar_modules")
    stub.PillarMaxPooling = Recorder
    sys.modules["pcdet.ops.pillar_ops.pillar_modules"] = stub
    spec = importlib.util.spec_from_file_location(
        "ref_dynamic_pillar_encoder", f"{REF}/pcdet/models/backbones_3d/vfe/pillarnet_modules/dynamic_pillar_encoder.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = synthetic.CONFIGS[dataset]
    reader = mod.DynamicPillarFeatureNet(num_input_features=num_input, num_filters=[32], pillar_size=0.16,
                                         pc_range=list(cfg["pc_range"]), virtual=virtual, encoding_type=encoding, dataset=dataset)
    pts = make_hybrid_points(dataset, Fin, 3, 300, seed=len(name), order=order)
    t = torch.from_numpy(pts)
    # PillarNet.forward's split (vfe/pillarnet.py:51-58; the file itself needs spconv to import)
    batch_size = t[:, 0].max().item()
    frames = [t[t[:, 0] == i][:, 1:] for i in range(int(batch_size) + 1)]
    reader(dict(points=frames))
    return dict(points=pts, xyz=seen["xyz"], xyz_batch_cnt=seen["cnt"], pt_features=seen["feat"],
                pc_range=np.asarray(cfg["pc_range"], dtype=np.float64),
                meta=np.asarray([dataset, str(Fin), str(num_input), str(int(virtual)), encoding, order]))


TRAIN_CASES = [
    # name, config, n points, P, use_abs, with_dist
    ("train_vod_p10",  "vod",  600, 10, True,  False),
    ("train_vod_p32",  "vod",  500, 32, True,  False),
    ("train_tj4d_p5",  "tj4d", 600,  5, True,  False),
    ("train_vod_rel",  "vod",  400, 10, False, False),
    ("train_vod_dist", "vod",  400, 10, True,  True),
]


def run_train(PillarVFE, PointPillarScatter, name, config, n, P, use_abs, with_dist):
    """The reference's PillarVFE in TRAIN mode under torch autograd (CPU): forward on batch statistics, running-stat update,
    gradients of sum(pillar_features * R) and of a random cotangent on the canvas."""
    cfg = synthetic.CONFIGS[config]
    F = cfg["F"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    pts, offs = synthetic.make_batch(config, 2, n, "clustered", seed0=41, oob_fraction=0.02)
    vox, coords, num = [], [], []
    for b in range(2):
        v, c, k = oracle.voxelize(pts[offs[b]:offs[b + 1]], geom, P, 40000, F=F, xcol=1)
        vox.append(v); num.append(k)
        coords.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], axis=1))
    vox, coords, num = np.concatenate(vox), np.concatenate(coords), np.concatenate(num)
    Cin = (F if use_abs else F - 3) + 6 + (1 if with_dist else 0)
    w = synthetic.make_pfn(Cin, 64, seed=len(name))
    model_cfg = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=with_dist, USE_ABSLOTE_XYZ=use_abs, NUM_FILTERS=[64])
    vfe = PillarVFE(model_cfg=model_cfg, num_point_features=F, voxel_size=list(cfg["voxel_size"]),
                    point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32))
    sd = vfe.state_dict()
    sd["pfn_layers.0.linear.weight"] = torch.from_numpy(w.weight)
    sd["pfn_layers.0.norm.weight"] = torch.from_numpy(w.gamma)
    sd["pfn_layers.0.norm.bias"] = torch.from_numpy(w.beta)
    sd["pfn_layers.0.norm.running_mean"] = torch.from_numpy(w.running_mean)
    sd["pfn_layers.0.norm.running_var"] = torch.from_numpy(w.running_var)
    vfe.load_state_dict(sd)
    vfe.train()
    bd = dict(voxels=torch.from_numpy(vox).float(), voxel_coords=torch.from_numpy(coords).float(),
              voxel_num_points=torch.from_numpy(num).float())
    bd = vfe(bd)
    pf = bd["pillar_features"]
    g = grid = geom.grid
    sc = PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=np.array([g[0], g[1], g[2]]))
    bd = sc(bd)
    canvas = bd["spatial_features"]
    rng = np.random.default_rng(len(name))
    R = rng.standard_normal(pf.shape).astype(np.float32)
    # a sparse cotangent on the canvas: values at the occupied cells of every 3rd pillar, one channel band
    Rc = np.zeros(canvas.shape, dtype=np.float32)
    sel = np.arange(0, coords.shape[0], 3)
    Rc[coords[sel, 0], 8:24, coords[sel, 2], coords[sel, 3]] = rng.standard_normal((sel.size, 16)).astype(np.float32)
    loss = (pf * torch.from_numpy(R)).sum() + (canvas * torch.from_numpy(Rc)).sum()
    loss.backward()
    lin, bn = vfe.pfn_layers[0].linear, vfe.pfn_layers[0].norm
    return dict(voxels=vox, voxel_coords=coords, voxel_num_points=num, weight=w.weight, gamma=w.gamma, beta=w.beta,
                running_mean=w.running_mean, running_var=w.running_var,
                pillar_features=pf.detach().numpy(), grad_out=R, grad_canvas_idx=sel.astype(np.int32),
                grad_canvas_vals=Rc[coords[sel, 0], 8:24, coords[sel, 2], coords[sel, 3]],
                grad_weight=lin.weight.grad.numpy(), grad_gamma=bn.weight.grad.numpy(), grad_beta=bn.bias.grad.numpy(),
                running_mean_after=bn.running_mean.numpy().copy(), running_var_after=bn.running_var.numpy().copy(),
                meta=np.asarray([config, str(P), str(int(use_abs)), str(int(with_dist)), torch.__version__]))


def main():
    check = "--check" in sys.argv
    if "--only-split" not in sys.argv:
        PillarVFE, PointPillarScatter, _ = load_reference()
        for case in TRAIN_CASES:
            data = run_train(PillarVFE, PointPillarScatter, *case)
            path = os.path.join(HERE, f"{case[0]}.npz")
            np.savez_compressed(path, **data)
            print(f"{case[0]:20s} M={data['voxels'].shape[0]:6d} -> {os.path.basename(path)} ({os.path.getsize(path) / 1e3:.0f} kB)")
        if "--only-train" in sys.argv:
            return
    for case in SPLIT_CASES:
        data = run_split(*case)
        path = os.path.join(HERE, f"{case[0]}.npz")
        np.savez_compressed(path, **data)
        print(f"{case[0]:18s} L={data['xyz'].shape[0]:6d} Fout={data['pt_features'].shape[1]} -> {os.path.basename(path)} ({os.path.getsize(path) / 1e3:.0f} kB)")
    if "--only-split" in sys.argv:
        return
    PillarVFE, PointPillarScatter, Radar7PillarVFE = load_reference()
    for name, flags, P in RADAR7_CASES:
        data = run_radar7(Radar7PillarVFE, name, flags, P)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{name:18s} M={data['voxels'].shape[0]:6d} Cin={data['weight'].shape[1]} -> {os.path.basename(path)} ({os.path.getsize(path) / 1e3:.0f} kB)")
    torch.manual_seed(0)
    for case in CASES:
        data, canvas = run_case(PillarVFE, PointPillarScatter, *case)
        path = os.path.join(HERE, f"vfe_{case[0]}.npz")
        np.savez_compressed(path, **data)
        msg = f"{case[0]:14s} M={data['voxels'].shape[0]:6d} -> {os.path.basename(path)} ({os.path.getsize(path) / 1e3:.0f} kB)"
        if check:
            cfg = synthetic.CONFIGS[case[1]]
            geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
            pfn = oracle.PfnParams(data["weight"], data["gamma"], data["beta"], data["running_mean"], data["running_var"])
            got = oracle.pillar_vfe(data["voxels"], data["voxel_coords"], data["voxel_num_points"], geom, pfn,
                                    use_absolute_xyz=case[7], with_distance=case[8])
            ref = data["pillar_features"]
            ieee = (got.view(np.uint32) == ref.view(np.uint32)).mean()
            got = oracle.pillar_vfe(data["voxels"], data["voxel_coords"], data["voxel_num_points"], geom, pfn,
                                    use_absolute_xyz=case[7], with_distance=case[8],
                                    invstd_override=data["ref_invstd"])
            eq = (got.view(np.uint32) == ref.view(np.uint32)).mean()
            msg += f" | ieee-invstd bit-equal {100 * ieee:.3f}%"
            g = geom.grid
            cv = oracle.pointpillar_scatter(got, data["voxel_coords"], canvas.shape[0], 64, int(g[1]), int(g[0]))
            msg += f" | torch-invstd bit-equal {100 * eq:.3f}% max|d|={np.abs(got - ref).max():.3g} canvas_eq={np.array_equal(cv.view(np.uint32), canvas.view(np.uint32)) if eq == 1 else np.allclose(cv, canvas)}"
        print(msg)


if __name__ == "__main__":
    main()
