"""Cycles per tile by class in k_emit (library built with -DHGSF_TILE_CLOCKS): python scripts/tile_clocks.py config mode B n"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
cfgname, mode, B, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
cfg = synthetic.CONFIGS[cfgname]; dev = torch.device("cuda:0")
path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, cfg["F"])
pf = device_pfn(synthetic.make_pfn(cfg["F"] + 6, 64), dev)
pts, _ = synthetic.make_batch(cfgname, B, n, mode)
d = torch.from_numpy(pts).to(dev)
res = path.points_to_bev(d, B, pf)
for _ in range(5): path.points_to_bev(d, B, pf, out=res)
torch.cuda.synchronize()
ws = path._ws
base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
off = base + 4096 if False else None
# workspace layout: ticket (512 B -> 512), state (256), then scan_desc [3*2048] u32
desc = base + 512 + 256
tail = ws[desc + 4 * (3 * 2048 - 128): desc + 4 * (3 * 2048)].view(torch.int64)
tail.zero_()
torch.cuda.synchronize()
path.points_to_bev(d, B, pf, out=res)
torch.cuda.synchronize()
v = tail.cpu().numpy().astype(np.float64)
names = ["empty", "1..10 points", "11.. points", "listed heavy", "skipped in window"]
warps, loop = v[21], v[20]
print(f"{cfgname} {mode} B={B} n={n}: warps {int(warps)}, mean loop cycles per warp {loop / warps:.0f}")
for c, nm in enumerate(names):
    cyc, cnt = v[c * 4: c * 4 + 3], v[c * 4 + 3]
    if cnt == 0: continue
    print(f"  {nm:18s} tiles {int(cnt):6d}  cycles/tile front {cyc[0] / cnt:7.0f} tile {cyc[1] / cnt:7.0f} hand-out {cyc[2] / cnt:7.0f}   share of warp time {100 * cyc.sum() / loop:5.1f} %")
