"""Train-step timing of FusedPillarVFE: the fused train path (hgsf_points_to_bev_train, 3 + 4 launches, no read-back) against the
contract-layout chain (pillarize -> trim -> batch statistics -> PFN -> scatter): python scripts/bench_train.py [config mode B n]"""
import json, os, sys
from types import SimpleNamespace
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from hgsfusion_b200 import modules, synthetic

cfgname, mode, B, n = (sys.argv[1:5] + ["vod", "clustered", "16", "30000"][len(sys.argv) - 1:])[:4]
B, n = int(B), int(n)
cfg = synthetic.CONFIGS[cfgname]
dev = torch.device("cuda:0")
pts, _ = synthetic.make_batch(cfgname, B, n, mode)
d = torch.from_numpy(pts).to(dev)
mc = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64], MAX_POINTS_PER_VOXEL=32,
                     MAX_NUMBER_OF_VOXELS={'train': 40000, 'test': 40000}, TRIM=False)
m = modules.FusedPillarVFE(model_cfg=mc, num_point_features=cfg["F"], voxel_size=list(cfg["voxel_size"]),
                           point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(dev).train()
Rc = None


def step(fused):
    global Rc
    m.zero_grad(set_to_none=True)
    out = m._forward_train_fused(d, B) if fused else m._forward_train(d, B)
    if Rc is None:
        Rc = torch.randn_like(out['spatial_features'])
    (out['spatial_features'] * Rc).sum().backward()


res = {}
for name, fused in (("fused_train", True), ("contract_chain", False)):
    for _ in range(5): step(fused)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 30
    e0.record()
    for _ in range(K): step(fused)
    e1.record(); torch.cuda.synchronize()
    res[name + "_ms_per_step"] = round(e0.elapsed_time(e1) / K, 4)
res["workload"] = f"{cfgname}_{mode}_b{B}_n{n}"
res["note"] = "forward + loss (canvas dot) + backward to linear.weight / norm.weight / norm.bias, torch autograd overhead included"
print(json.dumps(res))
