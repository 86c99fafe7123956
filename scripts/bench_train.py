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


def run(fused, what):
    """what: 'fwd' = the module's train-mode forward; 'bwd' = autograd.backward of the canvas with a ready cotangent (no loss
    kernels in the timed region); the forward of 'bwd' runs outside the events."""
    global Rc
    K = 30
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(-5, K):
        m.zero_grad(set_to_none=True)
        if what == 'fwd' and i >= 0: e[i][0].record()
        out = m._forward_train_fused(d, B) if fused else m._forward_train(d, B)
        if what == 'fwd' and i >= 0: e[i][1].record()
        if Rc is None:
            Rc = torch.randn_like(out['spatial_features'])
        if what == 'bwd' and i >= 0: e[i][0].record()
        torch.autograd.backward([out['spatial_features']], [Rc])
        if what == 'bwd' and i >= 0: e[i][1].record()
    torch.cuda.synchronize()
    return round(sum(a.elapsed_time(b) for a, b in e) / K, 4)


res = {"workload": f"{cfgname}_{mode}_b{B}_n{n}"}
for name, fused in (("fused_train", True), ("contract_chain", False)):
    res[name] = dict(forward_ms=run(fused, 'fwd'), backward_ms=run(fused, 'bwd'))
res["note"] = ("train-mode FusedPillarVFE, TRIM False; forward = points -> canvas on batch statistics (fused: 3 launches, no read-back; "
               "contract chain: pillarize -> trim (host sync) -> batch statistics -> PFN -> scatter); backward = canvas cotangent -> "
               "grads of linear.weight / norm.weight / norm.bias; torch autograd dispatch included, no loss kernels")
print(json.dumps(res))
