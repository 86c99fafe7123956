import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from types import SimpleNamespace
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
d = np.load(os.path.join(R, "tests/golden/vfe_tj4d_p32.npz"))
cfg = synthetic.CONFIGS["tj4d"]
dev = torch.device("cuda:0")
p = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, 8)
w = SimpleNamespace(weight=d["weight"], gamma=d["gamma"], beta=d["beta"], running_mean=d["running_mean"], running_var=d["running_var"], bias=None, eps=1e-3)
for fl in (True, False):
    co = torch.from_numpy(d["voxel_coords"]).to(dev); nu = torch.from_numpy(d["voxel_num_points"]).to(dev)
    if fl: co, nu = co.float(), nu.float()
    got = p.pillar_vfe(torch.from_numpy(d["voxels"]).to(dev), co, nu, device_pfn(w, dev)).cpu().numpy()
    ref = d["pillar_features"]
    bad = np.argwhere(np.abs(got - ref) > 1e-5 * np.abs(ref).max())
    print("float" if fl else "int", "bad elems", len(bad), "pillars", len(np.unique(bad[:, 0])), "channels", np.unique(bad[:, 1])[:10])
    if len(bad):
        i = bad[0]; print(i, got[i[0], i[1]], ref[i[0], i[1]], "num", d["voxel_num_points"][i[0]], "coords", d["voxel_coords"][i[0]])
        pil = np.unique(bad[:, 0]); print("nums of bad pillars", np.bincount(d["voxel_num_points"][pil])[:10], "coords y range", d["voxel_coords"][pil][:, 2].min(), d["voxel_coords"][pil][:, 2].max(), "x", d["voxel_coords"][pil][:, 3].min(), d["voxel_coords"][pil][:, 3].max())
