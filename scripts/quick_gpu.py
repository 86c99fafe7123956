import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R,'tests'))
import numpy as np, torch
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
cfg = synthetic.CONFIGS["vod"]
B, n = 16, 30000
pts, offs = synthetic.make_batch("vod", B, n, "clustered")
dev = torch.device("cuda:0")
w = synthetic.make_pfn(13, 64)
path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, 7)
dpts = torch.from_numpy(pts).to(dev)
pf = device_pfn(w, dev)
res = path.points_to_bev(dpts, B, pf)
torch.cuda.synchronize()
print("M", res.num_pillars.cpu().numpy())
for mode in ["clustered", "uniform"]:
    pts, offs = synthetic.make_batch("vod", B, n, mode)
    dpts = torch.from_numpy(pts).to(dev)
    res = path.points_to_bev(dpts, B, pf)
    for _ in range(5): path.points_to_bev(dpts, B, pf, out=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): path.points_to_bev(dpts, B, pf, out=res)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    M = int(res.num_pillars[0].item())
    alg = B * (4 * n * 7 + 4 * 64 * 320 * 320) + M * (4 * 64 + 16 + 4)
    print(mode, "ms/step", ms, "frames/s", B / ms * 1e3, "M", M, "GB/s", alg / ms / 1e6)
