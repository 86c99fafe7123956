"""A few steps of the fused path at the bench shape, for ncu (launch list / --set full)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn

config = sys.argv[1] if len(sys.argv) > 1 else "vod"
mode = sys.argv[2] if len(sys.argv) > 2 else "clustered"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n = int(sys.argv[4]) if len(sys.argv) > 4 else 30000
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cfg = synthetic.CONFIGS[config]
F = cfg["F"]
pts, offs = synthetic.make_batch(config, B, n, mode)
dev = torch.device("cuda:0")
path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, F)
pf = device_pfn(synthetic.make_pfn(F + 6, 64), dev)
dpts = torch.from_numpy(pts).to(dev)
res = path.points_to_bev(dpts, B, pf)
for _ in range(steps):
    path.points_to_bev(dpts, B, pf, out=res)
torch.cuda.synchronize()
print("ok", int(res.num_pillars[0].item()))
