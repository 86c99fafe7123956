"""Step and dominant-kernel time of the fused path on one workload (CUDA events, ring of distinct batches):
    python scripts/r2_step.py <config> <mode> <B> <n> [ring] [steps]
Prints ms/step, ms of the consumer kernel (k_emit), the rest (k_front + gap) and the roofline fractions (SURVEY 8d bytes)."""
import ctypes as C, json, os, statistics, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from hgsfusion_b200 import synthetic, _lib
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
config, mode, B, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
ring = int(sys.argv[5]) if len(sys.argv) > 5 else 10
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 100
P = int(os.environ.get("R2_P", "32")); MV = int(os.environ.get("R2_MV", "40000"))
cfg = synthetic.CONFIGS[config]; F = cfg["F"]
dev = torch.device("cuda:0")
lib = _lib.load()
pf = device_pfn(synthetic.make_pfn(F + 6, 64), dev)
path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, MV, F)
batches = [torch.from_numpy(synthetic.make_batch(config, B, n, mode, seed0=r * B)[0]).to(dev) for r in range(ring)]
res = path.points_to_bev(batches[0], B, pf)
for i in range(6): path.points_to_bev(batches[i % ring], B, pf, out=res)
torch.cuda.synchronize()
M = int(res.num_pillars[0].item())
lib.hgsf_emit_timing_begin(min(steps, 1024))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps): path.points_to_bev(batches[i % ring], B, pf, out=res)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
buf = (C.c_float * 1024)(); k = lib.hgsf_emit_timing_collect(buf, 1024); lib.hgsf_emit_timing_begin(0)
emit = statistics.fmean(buf[:k]) if k > 0 else float("nan")
alg = (4 * n * F + 4 * 64 * path.ny * path.nx + 84 * M / B) * B
peak = 6527.5
print(json.dumps(dict(workload=f"{config}_{mode}_b{B}_n{n}", ms_step=round(ms, 4), ms_emit=round(emit, 4), ms_rest=round(ms - emit, 4),
                      frames_per_s=round(B / ms * 1e3), pillars_per_frame=round(M / B), alg_mb=round(alg / 1e6, 1),
                      step_frac=round(alg / ms / 1e6 / peak, 3), emit_frac=round(alg / emit / 1e6 / peak, 3))))
