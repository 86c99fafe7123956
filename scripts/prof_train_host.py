"""Where the host time of the fused train forward goes: per-shape cost of torch.empty inside the call, and the rest."""
import os, sys, time, collections
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
real_empty = torch.empty
acc = collections.defaultdict(list)
def timed_empty(*a, **k):
    t0 = time.perf_counter(); r = real_empty(*a, **k); acc[tuple(r.shape)].append(time.perf_counter() - t0); return r
lib = m._path().lib
real_call = lib.hgsf_points_to_bev_train
def timed_call(*a):
    t0 = time.perf_counter(); r = real_call(*a); acc["hgsf_points_to_bev_train (3 launches)"].append(time.perf_counter() - t0); return r
for _ in range(5):
    out = m._forward_train_fused(d, B); torch.cuda.synchronize()
torch.empty = timed_empty
import hgsfusion_b200.ops as ops
ops.torch.empty = timed_empty
lib.hgsf_points_to_bev_train = timed_call
tot = []
for _ in range(50):
    t0 = time.perf_counter()
    out = m._forward_train_fused(d, B)
    tot.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
print("forward host total: median %.0f us" % (sorted(tot)[25] * 1e6))
for k, v in acc.items():
    print(k, "n=%d median %.1f us max %.1f us" % (len(v), sorted(v)[len(v) // 2] * 1e6, max(v) * 1e6))
