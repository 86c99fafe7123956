import os, sys, cProfile, pstats
from types import SimpleNamespace
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from hgsfusion_b200 import modules, synthetic
cfg = synthetic.CONFIGS["vod"]; dev = torch.device("cuda:0"); B = 16
pts, _ = synthetic.make_batch("vod", B, 30000, "clustered")
d = torch.from_numpy(pts).to(dev)
mc = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64], MAX_POINTS_PER_VOXEL=32,
                     MAX_NUMBER_OF_VOXELS={'train': 40000, 'test': 40000}, TRIM=False)
m = modules.FusedPillarVFE(model_cfg=mc, num_point_features=7, voxel_size=list(cfg["voxel_size"]),
                           point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(dev).train()
for _ in range(5):
    out = m._forward_train_fused(d, B); torch.cuda.synchronize()
def loop():
    for _ in range(50):
        out = m._forward_train_fused(d, B)
        torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); loop(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
