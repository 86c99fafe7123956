"""hgsf_assemble_hybrid_points (the step in front of the pillar path, on the device) against the reference fixtures and the
oracle: bit-exact rows, order and frame offsets."""
import glob
import os

import numpy as np
import pytest

from oracle import hybrid_oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "hybrid_*.npz")))


def _device_run(frames, Fr, use_virtual, calib_rows, pc_range, no_dup, dataset="vod"):
    import torch
    from hgsfusion_b200.hybrid_points import assemble_hybrid_points
    dev = torch.device("cuda:0")
    B = len(frames)
    cat = lambda k, w: torch.from_numpy(np.concatenate([f[k].reshape(-1, w) for f in frames], 0).astype(np.float32)).to(dev)
    offs = lambda k: np.concatenate([[0], np.cumsum([len(f[k]) for f in frames])]).astype(np.int32)
    kw = dict(batch_size=B, calib=None if calib_rows is None else np.stack(calib_rows), point_cloud_range=pc_range,
              no_dup=no_dup, dataset=dataset)
    if use_virtual:
        res = assemble_hybrid_points(cat(0, Fr), offs(0), cat(1, Fr + 8), offs(1), cat(2, Fr + 8), offs(2), **kw)
    else:
        res = assemble_hybrid_points(cat(0, Fr), offs(0), **kw)
    torch.cuda.synchronize()
    return res.trim().cpu().numpy(), res.frame_offsets.cpu().numpy()


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_device_assembly_matches_the_reference_fixtures(path):
    from hgsfusion_b200.hybrid_points import calib_row
    d = np.load(path)
    B, Fr = int(d["n_frames"]), int(d["Fr"])
    frames = [(d[f"real{b}"], d[f"gt{b}"], d[f"virt{b}"]) for b in range(B)]
    rows = [calib_row(d[f"V2C_{b}"], d[f"R0_{b}"], d[f"P2_{b}"], d["image_shape"]) for b in range(B)] if int(d["fov"]) else None
    got, offs = _device_run(frames, Fr, bool(int(d["use_virtual"])), rows, [float(v) for v in d["pc_range"]], bool(int(d["no_dup"])),
                            "tj4d" if Fr == 8 else "vod")
    ref = [d[f"points{b}"] for b in range(B)]
    assert list(offs) == list(np.concatenate([[0], np.cumsum([len(r) for r in ref])]))
    for b in range(B):
        mine = got[offs[b]:offs[b + 1]]
        assert (mine[:, 0] == b).all()
        assert np.array_equal(mine[:, 1:].view(np.uint32), ref[b].view(np.uint32)), (path, b)


def test_large_batch_empty_frames_and_pillar_path():
    """16 frames of ~30k candidates (some frames empty, some without masks), against the oracle; the result feeds
    hgsf_points_to_bev directly through frame_offsets (no host round trip)."""
    import torch
    from hgsfusion_b200.hybrid_points import calib_row
    import sys
    sys.path.insert(0, GOLDEN)
    from make_hybrid_golden import make_calib, make_frame      # input generators only (no reference access)
    rng = np.random.default_rng(7)
    pc_range = [0, -25.6, -3, 51.2, 25.6, 2]
    sizes = [(400, 300, 29000), (0, 0, 0), (350, 0, 500), (500, 200, 20000)] * 3 + [(0, 0, 0), (0, 5, 100), (300, 100, 0), (0, 0, 0)]
    frames, cal_rows, fovs = [], [], []
    for nr, ng, nv in sizes:
        frames.append(make_frame(rng, 7, nr, ng, nv, pc_range))
        c = make_calib(rng)
        cal_rows.append(calib_row(c["Tr_velo2cam"], c["R0"], c["P2"], (1216, 1936)))
        fovs.append((hybrid_oracle.lidar_to_rect_matrix(c["Tr_velo2cam"], c["R0"]), c["P2"], (1216, 1936)))
    ref, ref_offs = hybrid_oracle.assemble_batch([(f[0], f[1], f[2], fv) for f, fv in zip(frames, fovs)], no_dup=True,
                                                 pc_range=pc_range)
    got, offs = _device_run(frames, 7, True, cal_rows, pc_range, True)
    assert np.array_equal(offs, ref_offs)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # downstream: the assembled rows are what the shipped configuration's reader input prep takes (Path B, 17 -> 29 features)
    from hgsfusion_b200 import pillar_ops
    from oracle import pathb_oracle
    dev = torch.device("cuda:0")
    xyz, cnt, feat = pillar_ops.split_encode(torch.from_numpy(got).to(dev), pc_range, 29, virtual=True, encoding_type="split",
                                             dataset="vod", batch_size=len(sizes))
    torch.cuda.synchronize()
    oxyz, ocnt, ofeat = pathb_oracle.split_encode(ref, pc_range, "split", "vod", 29, True)
    assert np.array_equal(cnt.cpu().numpy()[:len(ocnt)], ocnt) and np.array_equal(np.diff(ref_offs)[:len(ocnt)], ocnt)
    assert np.array_equal(xyz.cpu().numpy(), oxyz) and np.array_equal(feat.cpu().numpy(), ofeat)


def test_sweep_only_and_argument_errors():
    import torch
    from hgsfusion_b200.hybrid_points import assemble_hybrid_points
    dev = torch.device("cuda:0")
    real = torch.randn(100, 7, device=dev)
    res = assemble_hybrid_points(real, [0, 60, 100], batch_size=2)
    torch.cuda.synchronize()
    out = res.trim().cpu().numpy()
    assert out.shape == (100, 8) and (out[:60, 0] == 0).all() and (out[60:, 0] == 1).all()
    assert np.array_equal(out[:, 1:], real.cpu().numpy())
    with pytest.raises(ValueError):
        assemble_hybrid_points(real.cpu(), [0, 100], batch_size=1)
    with pytest.raises(ValueError):       # the TJ4D broadcast error of the reference
        assemble_hybrid_points(torch.randn(10, 8, device=dev), [0, 10], torch.zeros(0, 16, device=dev), [0, 0],
                               torch.randn(5, 16, device=dev), [0, 5], batch_size=1, dataset="tj4d")
    # zero candidates: offsets are all zero, nothing launched
    res = assemble_hybrid_points(torch.zeros(0, 7, device=dev), [0, 0, 0], batch_size=2)
    torch.cuda.synchronize()
    assert res.frame_offsets.cpu().tolist() == [0, 0, 0]
