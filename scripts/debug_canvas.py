import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from oracle import oracle
from util import device_pfn, geom_for, oracle_pfn
dev = torch.device("cuda:0")
cfg = synthetic.CONFIGS["vod"]
for want_voxels in (False, True):
    B, n, P, mv = 3, 2000, 32, 40000
    pts, offs = synthetic.make_batch("vod", B, n, "clustered", seed0=4, oob_fraction=0.02)
    w = synthetic.make_pfn(13, 64, 2)
    ref = oracle.points_to_bev(pts, offs, geom_for("vod"), oracle_pfn(w), P, mv, F=7, xcol=1)
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, mv, 7)
    got = path.points_to_bev(torch.from_numpy(pts).to(dev), B, device_pfn(w, dev), want_voxels=want_voxels).trim()
    c = got["spatial_features"].cpu().numpy(); r = ref["spatial_features"]
    bad = np.argwhere(c.view(np.uint32) != r.view(np.uint32))
    print("want_voxels", want_voxels, "feats equal", np.array_equal(got["pillar_features"].cpu().numpy(), ref["pillar_features"]), "bad canvas elems", len(bad))
    if len(bad):
        cells = np.unique(bad[:, [0, 2, 3]], axis=0)
        print(" bad cells", len(cells), cells[:8].tolist(), "channels of first", np.unique(bad[(bad[:, [0, 2, 3]] == cells[0]).all(1)][:, 1])[:20])
        b0, y0, x0 = cells[0]
        print(" got", c[b0, :4, y0, x0], "ref", r[b0, :4, y0, x0], "num pillars in that tile row", (ref["voxel_coords"][:, 0] == b0).sum())
        occ = {(int(a), int(b_), int(c_)) for a, _, b_, c_ in ref["voxel_coords"]}
        print(" bad cell is a pillar cell:", [(tuple(int(v) for v in cc) in occ) for cc in cells[:8]])
