"""Reads k_front's per-phase %globaltimer stamps (library built with -DHGSF_PHASE_TIMES)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
cfg = synthetic.CONFIGS["vod"]; dev = torch.device("cuda:0"); B = 16
path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, 7)
pf = device_pfn(synthetic.make_pfn(13, 64), dev)
for mode in ("uniform", "clustered"):
    pts, _ = synthetic.make_batch("vod", B, 30000, mode)
    d = torch.from_numpy(pts).to(dev)
    res = path.points_to_bev(d, B, pf)
    for _ in range(5): path.points_to_bev(d, B, pf, out=res)
    torch.cuda.synchronize()
    ws = path._ws
    base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
    st = ws[base + 4 * 96: base + 4 * 96 + 8 * 6].view(torch.int64).cpu().numpy()      # k_front's stamps live at ticket[96..]
    t0, t1, t2, t3, t4, t5 = [int(v) for v in st]
    print(mode, f"ns: zero(skipped when clean) {t1-t0} count+sync {t2-t1} reduce+sync {t3-t2} apply+sync {t4-t3} fill {t5-t4} total {t5-t0}")
