"""ms/step of the fused path on one workload: python scripts/quick_cfg.py <config> <mode> <B> <n> [ring]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
# geometry experiments: the TJ4D range with 13 / 14 / 16 full tiles per row, feature counts swapped
synthetic.CONFIGS.update({
    "tj4d416": dict(pc_range=[0, -39.68, -4, 66.56, 39.68, 2], voxel_size=[0.16, 0.16, 6], F=8),
    "tj4d448": dict(pc_range=[0, -39.68, -4, 71.68, 39.68, 2], voxel_size=[0.16, 0.16, 6], F=8),
    "tj4d512": dict(pc_range=[0, -39.68, -4, 81.92, 39.68, 2], voxel_size=[0.16, 0.16, 6], F=8),
    "tj4d_f7": dict(pc_range=[0, -39.68, -4, 69.12, 39.68, 2], voxel_size=[0.16, 0.16, 6], F=7),
    "vod_f8": dict(pc_range=[0, -25.6, -3, 51.2, 25.6, 2], voxel_size=[0.16, 0.16, 5], F=8),
    "vod_y496": dict(pc_range=[0, -39.68, -3, 51.2, 39.68, 2], voxel_size=[0.16, 0.16, 5], F=7),
})
config, mode, B, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
ring = int(sys.argv[5]) if len(sys.argv) > 5 else 4
cfg = synthetic.CONFIGS[config]
F = cfg["F"]
dev = torch.device("cuda:0")
pf = device_pfn(synthetic.make_pfn(F + 6, 64), dev)
path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, F)
batches = [torch.from_numpy(synthetic.make_batch(config, B, n, mode, seed0=r * B)[0]).to(dev) for r in range(ring)]
res = path.points_to_bev(batches[0], B, pf)
for i in range(5): path.points_to_bev(batches[i % ring], B, pf, out=res)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100): path.points_to_bev(batches[i % ring], B, pf, out=res)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 100
print(f"{config} {mode} B={B} n={n}: ms/step {ms:.4f} frames/s {B / ms * 1e3:.0f}")
