"""GPU parity of the Path B (PillarNet reader) ops: against the numpy restatement, and -- when
oracle/_ref/libref_pillar_ops.so is present -- against the REFERENCE's own CUDA kernels compiled unmodified from
pcdet/ops/pillar_ops/src (oracle/build_ref_pillar_ops.sh) and run on the same GPU."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from hgsfusion_b200 import pillar_ops as po
from oracle import pathb_oracle as pb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_pillar_ops.so")


def make_points(B, n, seed, H=320, W=320, s=0.16, spread=1.15):
    rng = np.random.default_rng(seed)
    cnt = rng.integers(n // 2, n, size=B).astype(np.int32)
    N = int(cnt.sum())
    xyz = np.stack([rng.uniform(-0.1 * W * s, spread * W * s, N), rng.uniform(-0.1 * H * s, spread * H * s, N),
                    rng.uniform(-3, 2, N)], axis=1).astype(np.float32)
    # clusters, so that pillars hold several points
    k = N // 3
    xyz[:k, :2] = (rng.uniform(0, W * s, size=(1, 2)) + rng.normal(0, 0.5, size=(k, 2))).astype(np.float32)
    rng.shuffle(xyz)
    return xyz, cnt


@pytest.mark.parametrize("B,n,H,W,s", [(2, 3000, 320, 320, 0.16), (3, 5000, 496, 432, 0.16), (1, 1, 8, 8, 1.0), (4, 200, 16, 40, 0.5)])
def test_indices_match_numpy_oracle(cuda, B, n, H, W, s):
    xyz, cnt = make_points(B, n, seed=B * 7 + n, H=H, W=W, s=s)
    ref = pb.gen_indice_pairs_flat(xyz, cnt, s, H, W)
    got = po.gen_indice_pairs_flat(torch.from_numpy(xyz).to(cuda), torch.from_numpy(cnt).to(cuda), s, (H, W))
    assert got["counts"].tolist() == [ref["M"], ref["L"]]
    for k in ("pillars", "pillar_bev_indices", "indice_pairs", "point_set_indices", "pillar_set_indices"):
        assert np.array_equal(got[k].cpu().numpy(), ref[k]), k


def test_surplus_points_and_empty(cuda):
    xyz = np.array([[0.1, 0.1, 0]] * 5, dtype=np.float32)
    got = po.gen_indice_pairs_flat(torch.from_numpy(xyz).to(cuda), torch.tensor([2, 1], dtype=torch.int32, device=cuda), 1.0, (2, 2))
    assert got["pillars"].tolist() == [[0, 0, 0], [1, 0, 0]] and got["indice_pairs"].view(-1).tolist() == [0, 0, 1, 1, 1]
    got = po.gen_indice_pairs_flat(torch.zeros((0, 3), device=cuda), torch.tensor([0], dtype=torch.int32, device=cuda), 1.0, (4, 4))
    assert got["counts"].tolist() == [0, 0] and (got["pillar_bev_indices"] == -1).all()


def test_gather_scatter_forward_backward(cuda):
    rng = np.random.default_rng(3)
    N, Cf, L, M, Cc = 4000, 29, 3500, 900, 32
    feats = torch.from_numpy(rng.normal(size=(N, Cf)).astype(np.float32)).to(cuda).requires_grad_(True)
    idx_np = rng.integers(0, N, size=L).astype(np.int32)
    idx = torch.from_numpy(idx_np).to(cuda)
    g = po.gather_feature(feats, idx)
    assert np.array_equal(g.detach().cpu().numpy(), pb.gather_feature(feats.detach().cpu().numpy(), idx_np))
    go = torch.from_numpy(rng.normal(size=(L, Cf)).astype(np.float32)).to(cuda)
    g.backward(go)
    assert np.allclose(feats.grad.cpu().numpy(), pb.gather_feature_grad(idx_np, go.cpu().numpy(), N), atol=1e-5)

    src_np = rng.normal(size=(Cc, L)).astype(np.float32)
    pidx_np = rng.integers(0, M - 5, size=L).astype(np.int32)
    src = torch.from_numpy(src_np).to(cuda).requires_grad_(True)
    pidx = torch.from_numpy(pidx_np).to(cuda)
    out = po.scatter_max(src, pidx, M)
    ref = pb.scatter_max(src_np, pidx_np, M)
    assert np.array_equal(out.detach().cpu().numpy(), ref)          # max is order independent: exact
    gout = torch.from_numpy(rng.normal(size=(Cc, M)).astype(np.float32)).to(cuda)
    out.backward(gout)
    gs = src.grad.cpu().numpy()
    # every positive maximum routes its gradient to exactly one arg-max element; nothing else receives any
    nz = np.argwhere(gs != 0)
    assert len(nz) == int((ref > 0).sum())
    for c, p in nz[:: max(1, len(nz) // 300)]:
        assert abs(src_np[c, p] - ref[c, pidx_np[p]]) < 1e-5 and gs[c, p] == gout[c, pidx_np[p]].item()


# Path B geometries: VoD (square 320 x 320) and TJ4D (hgsfusion_tj4d.yaml: x 69.12 m -> Nx = 432, y 79.36 m -> Ny = 496; the
# 'split' encoding gives 29 / 31 point features).  bev_spatial_shape returns (W, H) = (Ny, Nx) and the reference hands exactly
# that to SparseConvTensor (pillar_modules.py:82), matching the (b, y, x) indices.
PATHB_GEOM = {
    "vod": dict(pc_range=[0, -25.6, -3, 51.2, 25.6, 2], Ny=320, Nx=320, Cf=29),
    "tj4d": dict(pc_range=[0, -39.68, -4, 69.12, 39.68, 2], Ny=496, Nx=432, Cf=31),
}


@pytest.mark.parametrize("geom", ["vod", "tj4d"])
def test_pillar_max_pooling_module(cuda, geom):
    """PillarMaxPooling (Path B reader, 'split' width 29+6 / 31+6 -> 32) against a plain torch composition."""
    gm = PATHB_GEOM[geom]
    Ny, Nx, Cf = gm["Ny"], gm["Nx"], gm["Cf"]
    rng = np.random.default_rng(5)
    rng_pc = gm["pc_range"]
    xyz_np, cnt_np = make_points(2, 4000, 11, H=Ny, W=Nx, spread=1.0)
    N = xyz_np.shape[0]
    pf_np = rng.normal(size=(N, Cf)).astype(np.float32)
    m = po.PillarMaxPooling([Cf + 6, 32], 0.16, rng_pc).to(cuda).eval()
    assert (m.bev_width, m.bev_height) == (Ny, Nx)
    with torch.no_grad():
        m.shared_mlps[1].running_mean.normal_(); m.shared_mlps[1].running_var.uniform_(0.5, 2.0)
        xyz, cnt, pf = torch.from_numpy(xyz_np).to(cuda), torch.from_numpy(cnt_np).to(cuda), torch.from_numpy(pf_np).to(cuda)
        res = m(xyz, cnt, pf)
        feats, pillars, shape, B = res if isinstance(res, tuple) else (res.features, res.indices, res.spatial_shape, res.batch_size)
        ref = pb.gen_indice_pairs_flat(xyz_np, cnt_np, 0.16, Ny, Nx)
        assert np.array_equal(pillars.cpu().numpy(), ref["pillars"]) and tuple(shape) == (Ny, Nx) and B == 2
        # the indices fit the shape the tensor is wrapped with: y < shape[0], x < shape[1] (and both extremes are reached)
        assert int(pillars[:, 1].max()) < shape[0] and int(pillars[:, 2].max()) < shape[1]
        assert Ny == Nx or int(pillars[:, 1].max()) >= Nx          # a transposed (Nx, Ny) shape could not hold these rows
        assert tuple(m.pillar_bev_indices.shape) == (2, Ny, Nx)
        pi = torch.from_numpy(ref["point_set_indices"]).long().to(cuda)
        qi = torch.from_numpy(ref["pillar_set_indices"]).long().to(cuda)
        centers = torch.zeros((ref["M"], 3), device=cuda)
        P = torch.from_numpy(ref["pillars"]).to(cuda)
        centers[:, 0] = (P[:, 2] + 0.5) * 0.16; centers[:, 1] = (P[:, 1] + 0.5) * 0.16; centers[:, 2] = (rng_pc[5] + rng_pc[2]) / 2
        gfeat = torch.cat([pf[pi], xyz[pi], xyz[pi] - centers[qi]], dim=1)
        h = m.shared_mlps(gfeat)
        exp = torch.zeros((ref["M"], 32), device=cuda)
        exp.index_reduce_(0, qi, h, "amax", include_self=True)
        # eval + no_grad is the fused native reader (hgsf_pillarnet_reader): its Linear is a sequential fp32 FMA chain, torch's is
        # cuBLAS -- same inputs, different summation order: 1e-5 relative (SURVEY.md 8c parity rule for Path B features)
        assert m._fused_ok(pf)
        assert torch.allclose(feats, exp, rtol=1e-5, atol=1e-6 * float(exp.abs().max()))
        assert torch.equal(feats == 0, exp == 0)                 # the ReLU / empty pattern is identical
    # with gradients enabled the module runs the reference's own composition (native gather / scatter_max around torch's MLP)
    assert not m._fused_ok(pf)
    res2 = m(xyz, cnt, pf)
    feats2 = res2[0] if isinstance(res2, tuple) else res2.features
    assert feats2.requires_grad and torch.equal(feats2.detach(), exp)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_pillar_ops.so not built (needs /root/reference)")
def test_against_the_reference_kernels(cuda):
    """Bit-for-bit against the reference's own kernels (compiled from its unmodified sources) on this GPU."""
    ref = C.CDLL(REF_SO)
    B, H, W, s = 3, 320, 320, 0.16
    xyz_np, cnt_np = make_points(B, 6000, 21)
    N = xyz_np.shape[0]
    xyz, cnt = torch.from_numpy(xyz_np).to(cuda), torch.from_numpy(cnt_np).to(cuda)
    p = lambda t: C.c_void_p(t.data_ptr())
    # the reference's Python glue (pillar_utils.py:99-124, group_utils.py:20-29) around its kernels
    mask = torch.zeros((B, H, W), dtype=torch.bool, device=cuda)
    torch.cuda.synchronize()
    ref.ref_create_pillar_indices_stack(N, B, H, W, C.c_float(s), p(xyz), p(cnt), p(mask))
    location = torch.cumsum(mask.view(-1), 0).int()
    M = location[-1].item()
    bev = (location.view(B, H, W) * mask - 1).int().contiguous()
    pillars = torch.zeros((M, 3), dtype=torch.int32, device=cuda)
    ref.ref_create_pillar_indices(B, H, W, p(bev), p(pillars))
    pairs = torch.full((N, 1), -1, dtype=torch.int32, device=cuda)
    ref.ref_create_pillar_indice_pairs_stack(N, B, H, W, C.c_float(s), p(xyz), p(cnt), p(bev), p(pairs))
    valid = pairs.view(-1) > -1
    position = torch.cumsum(valid, 0).int()
    L = position[-1].item()
    position = (position * valid - 1).int().contiguous()
    first = torch.zeros(L, dtype=torch.int32, device=cuda)
    second = torch.zeros(L, dtype=torch.int32, device=cuda)
    ref.ref_flatten_indice_pairs(N, 1, p(pairs), p(position), p(first), p(second))
    torch.cuda.synchronize()
    got = po.gen_indice_pairs_flat(xyz, cnt, s, (H, W))
    assert got["counts"].tolist() == [M, L]
    assert torch.equal(got["pillars"], pillars) and torch.equal(got["pillar_bev_indices"], bev)
    assert torch.equal(got["indice_pairs"], pairs)
    assert torch.equal(got["point_set_indices"], first) and torch.equal(got["pillar_set_indices"], second)
    # gather / scatter_max
    feats = torch.randn((N, 31), device=cuda)
    out_ref = torch.zeros((L, 31), device=cuda)
    ref.ref_gather_feature(L, 31, p(first), p(feats), p(out_ref))
    torch.cuda.synchronize()
    assert torch.equal(po.gather_feature(feats, first), out_ref)
    src = torch.randn((32, L), device=cuda)
    arg_ref = torch.full((32, M), -1, dtype=torch.int32, device=cuda)
    o_ref = torch.zeros((32, M), device=cuda)
    ref.ref_scatter_max(32, L, M, p(second), p(src), p(arg_ref), p(o_ref))
    torch.cuda.synchronize()
    o = po.scatter_max(src, second, M)
    assert torch.equal(o, o_ref)
    assert pb.check_arg(arg_ref.cpu().numpy(), src.cpu().numpy(), second.cpu().numpy(), o_ref.cpu().numpy())
    # our arg through the backward pass: same rule
    gout = torch.randn((32, M), device=cuda)
    gs_ref = torch.zeros((32, L), device=cuda)
    ref.ref_scatter_max_grad(32, M, p(arg_ref), p(gout), p(gs_ref))
    torch.cuda.synchronize()
    src2 = src.clone().requires_grad_(True)
    po.scatter_max(src2, second, M).backward(gout)
    # identical wherever the maximum is attained by a single element (ties may route differently, as in the reference)
    same = (src2.grad == gs_ref).float().mean().item()
    assert same > 0.999
    assert torch.equal((src2.grad != 0).sum(), (gs_ref != 0).sum())


# ---- split_encode (a13): PillarNet.forward's frame split + DynamicPillarFeatureNet.forward's encoding -----------------
import glob

SPLIT_FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "split_*.npz")))


@pytest.mark.parametrize("path", SPLIT_FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("give_batch_size", [False, True])
def test_split_encode_matches_reference_fixture(cuda, path, give_batch_size):
    d = np.load(path)
    dataset, Fin, num_input, virtual, encoding, order = d["meta"]
    B = len(d["xyz_batch_cnt"]) if give_batch_size else None
    xyz, cnt, feat = po.split_encode(torch.from_numpy(d["points"]).to(cuda), d["pc_range"], int(num_input), bool(int(virtual)),
                                     str(encoding), str(dataset), B)
    assert cnt.dtype == torch.int32 and np.array_equal(cnt.cpu().numpy(), d["xyz_batch_cnt"])
    assert np.array_equal(xyz.cpu().numpy().view(np.uint32), d["xyz"].view(np.uint32))
    assert np.array_equal(feat.cpu().numpy().view(np.uint32), d["pt_features"].view(np.uint32))


@pytest.mark.parametrize("dataset,Fin,Fout", [("vod", 17, 29), ("tj4d", 18, 31)])
def test_split_encode_large_and_dropped_rows(cuda, dataset, Fin, Fout):
    rng = np.random.default_rng(Fin)
    B, L = 16, 480_000 + 13
    pts = rng.standard_normal((L, 1 + Fin)).astype(np.float32)
    pts[:, 0] = np.sort(rng.integers(0, B, L))
    pts[:, -2] = rng.integers(0, 2, L)
    pts[:, -1] = rng.integers(0, 2, L)
    pts[-5:, 0] = [B, -1, 0.5, np.nan, 99]                        # match no `points[:,0] == i`: dropped (tail -> no reorder needed)
    rng_range = [0, -25.6, -3, 51.2, 25.6, 2]
    ref = pb.split_encode(pts[:-5], rng_range, "split", dataset, Fout)
    xyz, cnt, feat = po.split_encode(torch.from_numpy(pts).to(cuda), rng_range, Fout, True, "split", dataset, B)
    assert np.array_equal(cnt.cpu().numpy(), ref[1])
    assert np.array_equal(xyz.cpu().numpy().view(np.uint32), ref[0].view(np.uint32))
    assert np.array_equal(feat.cpu().numpy().view(np.uint32), ref[2].view(np.uint32))
    # a dropped row in the middle and frames out of order take the ordered second pass
    pts2 = pts[rng.permutation(L)]
    ref2 = pb.split_encode(pts2[np.isfinite(pts2[:, 0])], rng_range, "split", dataset, Fout)   # (the oracle's max() must not see the NaN)
    xyz, cnt, feat = po.split_encode(torch.from_numpy(pts2).to(cuda), rng_range, Fout, True, "split", dataset, B)
    keep = ref2[1][:B]
    assert np.array_equal(cnt.cpu().numpy(), keep)
    n = int(keep.sum())
    assert xyz.shape[0] == n and np.array_equal(feat.cpu().numpy().view(np.uint32), ref2[2][:n].view(np.uint32))


@pytest.mark.parametrize("geom", ["vod", "tj4d"])
def test_dynamic_pillar_feature_net_reader(cuda, geom):
    """The module mirror end to end: collated points in, the reader's (features, pillars, shape, B) out; both input forms agree.
    TJ4D: 18 -> 31 'split' columns, 31 + 6 -> 32 reader, non-square (496, 432) shape."""
    gm = PATHB_GEOM[geom]
    d = np.load(os.path.join(ROOT, "tests", "golden", f"split_{geom}.npz"))
    nb = len(d["xyz_batch_cnt"])
    torch.manual_seed(0)
    net = po.DynamicPillarFeatureNet(num_input_features=gm["Cf"], num_filters=[32], pillar_size=0.16, virtual=True,
                                     pc_range=gm["pc_range"], encoding_type="split", dataset=geom).to(cuda).eval()
    assert "pfn_layers.shared_mlps.0.weight" in net.state_dict()
    assert net.state_dict()["pfn_layers.shared_mlps.0.weight"].shape == (32, gm["Cf"] + 6)
    pts = torch.from_numpy(d["points"]).to(cuda)
    with torch.no_grad():
        a = net(dict(points=pts, batch_size=nb))
        frames = [pts[pts[:, 0] == i][:, 1:] for i in range(nb)]
        b = net(dict(points=frames))
    fa, pa, sa = (a.features, a.indices, a.spatial_shape) if hasattr(a, "features") else (a[0], a[1], a[2])
    fb, pb_ = (b.features, b.indices) if hasattr(b, "features") else (b[0], b[1])
    assert torch.equal(fa, fb) and torch.equal(pa, pb_) and fa.shape[1] == 32 and pa.shape[1] == 3
    assert tuple(sa) == (gm["Ny"], gm["Nx"])
    assert int(pa[:, 1].max()) < sa[0] and int(pa[:, 2].max()) < sa[1] and int(pa[:, 0].max()) < nb
    # the reader's pillar list against the numpy restatement of the reference's index generation on the fixture's own xyz
    ref = pb.gen_indice_pairs_flat(d["xyz"], d["xyz_batch_cnt"], 0.16, gm["Ny"], gm["Nx"])
    assert np.array_equal(pa.cpu().numpy(), ref["pillars"])


@pytest.mark.parametrize("Cf,dataset", [(29, "vod"), (31, "tj4d"), (7, "vod")])
def test_fused_reader_entry_point(cuda, Cf, dataset):
    """hgsf_pillarnet_reader against a float64 restatement of gather + centre offsets + Linear + BN + ReLU + scatter_max."""
    rng = np.random.default_rng(Cf)
    xyz_np, cnt_np = make_points(3, 5000, Cf, spread=1.0)
    N = xyz_np.shape[0]
    pf_np = rng.normal(size=(N, Cf)).astype(np.float32)
    m = po.PillarMaxPooling([Cf + 6, 32], 0.16, [0, -25.6, -3, 51.2, 25.6, 2]).to(cuda).eval()
    with torch.no_grad():
        bn = m.shared_mlps[1]
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.5)
        res = m(torch.from_numpy(xyz_np).to(cuda), torch.from_numpy(cnt_np).to(cuda), torch.from_numpy(pf_np).to(cuda))
    feats = (res[0] if isinstance(res, tuple) else res.features).cpu().numpy()
    ref = pb.gen_indice_pairs_flat(xyz_np, cnt_np, 0.16, 320, 320)
    pi, qi, P = ref["point_set_indices"], ref["pillar_set_indices"], ref["pillars"]
    centers = np.stack([(P[:, 2] + 0.5) * np.float32(0.16), (P[:, 1] + 0.5) * np.float32(0.16), np.full(P.shape[0], -0.5)], axis=1)
    g = np.concatenate([pf_np[pi], xyz_np[pi], xyz_np[pi] - centers[qi].astype(np.float32)], axis=1).astype(np.float64)
    W = m.shared_mlps[0].weight.detach().cpu().numpy().astype(np.float64)
    x = g @ W.T
    y = (x - bn.running_mean.cpu().numpy()) / np.sqrt(bn.running_var.cpu().numpy().astype(np.float64) + bn.eps) * bn.weight.detach().cpu().numpy() + bn.bias.detach().cpu().numpy()
    exp = np.zeros((ref["M"], 32))
    np.maximum.at(exp, qi, np.maximum(y, 0))
    assert feats.shape == exp.shape
    assert np.allclose(feats, exp, rtol=1e-5, atol=2e-6 * np.abs(exp).max())


@pytest.mark.parametrize("C,ny,nx,B", [(32, 320, 320, 2), (64, 160, 160, 3), (128, 80, 80, 2), (256, 40, 40, 1), (64, 62, 54, 2)])
def test_sparse_to_dense(C, ny, nx, B):
    """the .dense() of the PillarNet branch (lss_fpn.py:111-113) at strides 1/2/4/8."""
    import torch
    from hgsfusion_b200 import pillar_ops
    from oracle import pathb_oracle
    rng = np.random.default_rng(C + nx)
    cells = rng.choice(B * ny * nx, size=min(B * ny * nx // 5, 20000), replace=False)
    idx = np.stack([cells // (ny * nx), (cells % (ny * nx)) // nx, cells % nx], 1).astype(np.int32)
    feats = rng.normal(0, 1, (len(idx), C)).astype(np.float32)
    dev = torch.device("cuda:0")
    got = pillar_ops.sparse_to_dense(torch.from_numpy(feats).to(dev), torch.from_numpy(idx).to(dev), (ny, nx), B)
    torch.cuda.synchronize()
    ref = pathb_oracle.sparse_to_dense(feats, idx, (ny, nx), B)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    # empty tensor and out-of-range rows
    got = pillar_ops.sparse_to_dense(torch.zeros((0, C), device=dev), torch.zeros((0, 3), dtype=torch.int32, device=dev), (ny, nx), B)
    assert float(got.abs().max()) == 0.0
    with pytest.raises(Exception):
        pillar_ops.sparse_to_dense(torch.zeros((4, 48), device=dev), torch.zeros((4, 3), dtype=torch.int32, device=dev), (ny, nx), B)
