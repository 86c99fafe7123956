"""A few fused train steps for ncu launch lists: python scripts/prof_train.py [steps]"""
import os, sys
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hgsfusion_b200 import modules, synthetic
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = synthetic.CONFIGS["vod"]; dev = torch.device("cuda:0"); B = 16
pts, _ = synthetic.make_batch("vod", B, 30000, "clustered")
d = torch.from_numpy(pts).to(dev)
mc = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64], MAX_POINTS_PER_VOXEL=32,
                     MAX_NUMBER_OF_VOXELS={'train': 40000, 'test': 40000}, TRIM=False)
m = modules.FusedPillarVFE(model_cfg=mc, num_point_features=7, voxel_size=list(cfg["voxel_size"]),
                           point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(dev).train()
Rc = None
for _ in range(steps):
    m.zero_grad(set_to_none=True)
    out = m._forward_train_fused(d, B)
    if Rc is None:
        Rc = torch.randn_like(out['spatial_features'])
    torch.autograd.backward([out['spatial_features']], [Rc])
torch.cuda.synchronize()
print("ok")
