"""GPU parity of the pillar-list consumer: hgsf_subm_neighbors / hgsf_subm_conv3x3 and the
SpMiddlePillarEncoder18.conv1 mirror against the numpy oracle (itself pinned against torch's dense conv2d on the active set,
tests/test_pathb_oracle.py), and at full size against cuDNN's dense fp32 convolution on the same GPU.

Tolerance: 1e-5 of the largest output (fp32; the summation order of a 288-term dot product differs between
implementations)."""
import ctypes as C

import numpy as np
import pytest
import torch

from hgsfusion_b200 import _lib, pillar_ops as po
from oracle import pathb_oracle as pb
from test_pathb_oracle import _random_pillars

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _close(got, ref):
    return np.abs(got - ref).max() <= TOL * max(1.0, np.abs(ref).max())


def _bn(rng, Cc, dev):
    bn = torch.nn.BatchNorm1d(Cc, eps=1e-3, momentum=0.01).to(dev).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, Cc).astype(np.float32)))
        bn.bias.copy_(torch.from_numpy(rng.normal(0, 0.5, Cc).astype(np.float32)))
        bn.running_mean.copy_(torch.from_numpy(rng.normal(0, 1, Cc).astype(np.float32)))
        bn.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 2, Cc).astype(np.float32)))
    return bn


def _bn_np(bn):
    return tuple(t.detach().cpu().numpy() for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))


@pytest.mark.parametrize("B,H,W,M_per", [(2, 40, 40, 500), (3, 17, 33, 561), (1, 1, 1, 1), (2, 64, 64, 70)])
def test_neighbors_bit_exact(cuda, B, H, W, M_per):
    rng = np.random.default_rng(B * 100 + H)
    pillars, bev = _random_pillars(rng, B, H, W, M_per)
    got = po.subm_neighbors(torch.from_numpy(bev).to(cuda), torch.from_numpy(pillars).to(cuda))
    assert np.array_equal(got.cpu().numpy(), pb.subm_neighbors(bev, pillars))


@pytest.mark.parametrize("Cin,Cout", [(32, 32), (64, 64), (32, 64), (64, 128), (128, 128), (128, 256), (256, 256)])
@pytest.mark.parametrize("layout", ["KRSC", "RSCK"])
@pytest.mark.parametrize("opts", [dict(), dict(bias=True), dict(bias=True, bn=True, relu=True), dict(bias=True, bn=True, res=True, relu=True)])
def test_conv_matches_oracle(cuda, Cin, Cout, layout, opts):
    rng = np.random.default_rng(Cin * 7 + Cout + len(opts))
    pillars, bev = _random_pillars(rng, 2, 37, 29, 333)       # M = 666: ragged last tile
    M = pillars.shape[0]
    feats = rng.normal(size=(M, Cin)).astype(np.float32)
    w = (rng.normal(size=(Cout, 3, 3, Cin)) * 0.1).astype(np.float32)
    bias = rng.normal(size=Cout).astype(np.float32) if opts.get("bias") else None
    bn = _bn(rng, Cout, cuda) if opts.get("bn") else None
    res = rng.normal(size=(M, Cout)).astype(np.float32) if opts.get("res") else None
    nbr = pb.subm_neighbors(bev, pillars)
    ref = pb.subm_conv3x3(feats, nbr, w, bias=bias, bn=_bn_np(bn) if bn else None, residual=res, relu=bool(opts.get("relu")))
    wt = torch.from_numpy(w if layout == "KRSC" else np.ascontiguousarray(w.transpose(1, 2, 3, 0))).to(cuda)
    got = po.subm_conv3x3(torch.from_numpy(feats).to(cuda), torch.from_numpy(nbr).to(cuda), wt,
                          bias=None if bias is None else torch.from_numpy(bias).to(cuda), bn=bn,
                          residual=None if res is None else torch.from_numpy(res).to(cuda), relu=bool(opts.get("relu")),
                          weight_layout=layout)
    assert _close(got.cpu().numpy(), ref)


def test_isolated_pillars_and_empty(cuda):
    """A pillar without neighbours sees only its centre tap; M = 0 launches nothing."""
    rng = np.random.default_rng(5)
    pillars = np.array([[0, 0, 0], [0, 5, 5], [1, 9, 9]], dtype=np.int32)
    bev = np.full((2, 10, 10), -1, dtype=np.int32)
    bev[pillars[:, 0], pillars[:, 1], pillars[:, 2]] = np.arange(3)
    feats = rng.normal(size=(3, 32)).astype(np.float32)
    w = rng.normal(size=(32, 3, 3, 32)).astype(np.float32)
    nbr = po.subm_neighbors(torch.from_numpy(bev).to(cuda), torch.from_numpy(pillars).to(cuda))
    got = po.subm_conv3x3(torch.from_numpy(feats).to(cuda), nbr, torch.from_numpy(w).to(cuda)).cpu().numpy()
    assert _close(got, feats @ w[:, 1, 1, :].T)
    e = po.subm_conv3x3(torch.zeros((0, 32), device=cuda), torch.zeros((0, 9), dtype=torch.int32, device=cuda), torch.from_numpy(w).to(cuda))
    assert e.shape == (0, 32) and _lib.load().hgsf_last_launch_count() == 0


def test_device_row_count_leaves_the_tail_untouched(cuda):
    rng = np.random.default_rng(6)
    pillars, bev = _random_pillars(rng, 2, 20, 20, 100)
    M = pillars.shape[0]
    cap = M + 77
    pad = np.zeros((cap, 3), dtype=np.int32); pad[:M] = pillars; pad[M:] = -7          # garbage rows past the count
    feats = rng.normal(size=(cap, 32)).astype(np.float32)
    w = (rng.normal(size=(32, 3, 3, 32)) * 0.1).astype(np.float32)
    m_dev = torch.tensor([M], dtype=torch.int32, device=cuda)
    nbr = torch.full((cap, 9), -5, dtype=torch.int32, device=cuda)
    lib = _lib.load()
    bev_t, pad_t = torch.from_numpy(bev).to(cuda), torch.from_numpy(pad).to(cuda)
    st = lib.hgsf_subm_neighbors(C.c_void_p(bev_t.data_ptr()), C.c_void_p(pad_t.data_ptr()), cap, C.c_void_p(m_dev.data_ptr()), 2, 20, 20, C.c_void_p(nbr.data_ptr()), po._s())
    assert st == 0
    assert np.array_equal(nbr[:M].cpu().numpy(), pb.subm_neighbors(bev, pillars)) and (nbr[M:] == -5).all()
    out = torch.full((cap, 32), 123.0, device=cuda)
    po.subm_conv3x3(torch.from_numpy(feats).to(cuda), nbr, torch.from_numpy(w).to(cuda), num_rows_dev=m_dev, out=out)
    assert _close(out[:M].cpu().numpy(), pb.subm_conv3x3(feats[:M], pb.subm_neighbors(bev, pillars), w)) and (out[M:] == 123.0).all()


def _conv1_oracle(enc, feats, nbr):
    """Sparse2DBasicBlockV.forward + Sparse2DBasicBlock.forward (pcnres18.py:139-151,176-187) composed from the oracle's convolution."""
    def cb(seq, x, res):
        return pb.subm_conv3x3(x, nbr, seq[0].weight.detach().cpu().numpy(), bias=seq[0].bias.detach().cpu().numpy(), bn=_bn_np(seq[1]),
                               residual=res, relu=True)
    b0, b1 = getattr(enc, "0"), getattr(enc, "1")
    identity = cb(b0.conv0, feats, None)
    x = cb(b0.conv2, cb(b0.conv1, identity, None), identity)
    return cb(b1.conv2, cb(b1.conv1, x, None), x)


def test_encoder_conv1_on_the_reader_output(cuda):
    """points -> hgsf_pillarnet_indices -> (random reader features) -> conv1 stage, against the oracle's composition."""
    from test_gpu_pillarnet import make_points
    rng = np.random.default_rng(11)
    xyz, cnt = make_points(2, 3000, seed=4)
    r = po.gen_indice_pairs_flat(torch.from_numpy(xyz).to(cuda), torch.from_numpy(cnt).to(cuda), 0.16, (320, 320))
    pillars, bev = r["pillars"], r["pillar_bev_indices"]
    M = int(pillars.shape[0])
    feats = np.abs(rng.normal(size=(M, 32))).astype(np.float32)
    torch.manual_seed(0)
    enc = po.PillarEncoderConv1(32).to(cuda).eval()
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            src = _bn(rng, 32, cuda)
            m.load_state_dict(src.state_dict())
    keys = set(enc.state_dict().keys())
    assert {"0.conv0.0.weight", "0.conv0.0.bias", "0.conv0.1.running_var", "0.conv2.1.weight", "1.conv1.0.weight", "1.conv2.1.bias"} <= keys
    got = enc(torch.from_numpy(feats).to(cuda), pillars, bev).cpu().numpy()
    ref = _conv1_oracle(enc, feats, pb.subm_neighbors(bev.cpu().numpy(), pillars.cpu().numpy()))
    assert got.shape == (M, 32) and _close(got, ref)


def test_full_size_against_cudnn_dense_fp32(cuda):
    """Config-2 size (16 frames, ~11k pillars each): the kernel against cuDNN's dense fp32 convolution (TF32 off) of the
    densified input, sampled at the active cells, on the same GPU."""
    import torch.nn.functional as F
    rng = np.random.default_rng(21)
    B, H, W = 16, 320, 320
    pillars, bev = _random_pillars(rng, B, H, W, 11400)
    pillars_t, bev_t = torch.from_numpy(pillars).to(cuda), torch.from_numpy(bev).to(cuda)
    M = pillars.shape[0]
    feats = torch.from_numpy(rng.normal(size=(M, 32)).astype(np.float32)).to(cuda)
    w = torch.from_numpy((rng.normal(size=(32, 3, 3, 32)) * 0.1).astype(np.float32)).to(cuda)
    bias = torch.from_numpy(rng.normal(size=32).astype(np.float32)).to(cuda)
    got = po.subm_conv3x3(feats, po.subm_neighbors(bev_t, pillars_t), w, bias=bias)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        dense = po.sparse_to_dense(feats, pillars_t, (H, W), B)
        ref = F.conv2d(dense, w.permute(0, 3, 1, 2).contiguous(), bias, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    idx = pillars_t.long()
    ref = ref[idx[:, 0], :, idx[:, 1], idx[:, 2]]
    assert (got - ref).abs().max().item() <= TOL * max(1.0, ref.abs().max().item())


def test_invalid_arguments(cuda):
    lib = _lib.load()
    cv = _lib.SubmConv()
    w = torch.zeros((48, 3, 3, 48), device=cuda)
    f = torch.zeros((4, 48), device=cuda)
    n = torch.zeros((4, 9), dtype=torch.int32, device=cuda)
    o = torch.zeros((4, 48), device=cuda)
    cv.weight, cv.in_channels, cv.out_channels = w.data_ptr(), 48, 48
    args = lambda feat, out: (C.byref(cv), C.c_void_p(feat.data_ptr()), C.c_void_p(n.data_ptr()), 4, None, None, C.c_void_p(out.data_ptr()), po._s())
    assert lib.hgsf_subm_conv3x3(*args(f, o)) == _lib.ERR_UNSUPPORTED          # 48 channels: outside the compiled set
    cv.in_channels = cv.out_channels = 32
    assert lib.hgsf_subm_conv3x3(*args(f, f)) == _lib.ERR_INVALID_ARG          # in-place
    cv.bn_weight = w.data_ptr()
    assert lib.hgsf_subm_conv3x3(*args(f, o)) == _lib.ERR_INVALID_ARG          # BatchNorm pointers: all or none
    with pytest.raises(ValueError):
        po.subm_conv3x3(f.cpu(), n, w)                                         # no CPU path


@pytest.mark.parametrize("B,H,W,M_per", [(2, 40, 40, 300), (3, 17, 33, 200), (1, 1, 1, 1), (2, 2, 1, 2), (4, 64, 64, 4096), (2, 31, 8, 5)])
def test_stride2_indices_bit_exact(cuda, B, H, W, M_per):
    rng = np.random.default_rng(B * 1000 + H * 7 + W)
    pillars, bev = _random_pillars(rng, B, H, W, M_per)
    r = po.sparse_conv_s2_indices(torch.from_numpy(bev).to(cuda), torch.from_numpy(pillars).to(cuda))
    out_pillars, out_bev, nbr = pb.sparse_conv_s2_indices(bev, pillars)
    assert int(r["counts"][0].item()) == out_pillars.shape[0]
    assert np.array_equal(r["pillars"].cpu().numpy(), out_pillars)
    assert np.array_equal(r["pillar_bev_indices"].cpu().numpy(), out_bev)
    assert np.array_equal(r["neighbors"].cpu().numpy(), nbr)


def test_stride2_indices_with_a_device_row_count(cuda):
    rng = np.random.default_rng(77)
    pillars, bev = _random_pillars(rng, 2, 20, 20, 90)
    M = pillars.shape[0]
    pad = np.full((M + 50, 3), -3, dtype=np.int32); pad[:M] = pillars
    m_dev = torch.tensor([M], dtype=torch.int32, device=cuda)
    r = po.sparse_conv_s2_indices(torch.from_numpy(bev).to(cuda), torch.from_numpy(pad).to(cuda), num_rows_dev=m_dev)
    out_pillars, out_bev, nbr = pb.sparse_conv_s2_indices(bev, pillars)
    assert np.array_equal(r["pillars"].cpu().numpy(), out_pillars) and np.array_equal(r["neighbors"].cpu().numpy(), nbr)


def test_encoder_conv2_against_oracle_and_cudnn(cuda):
    """conv2 stage (stride-2 SparseConv2d 32 -> 64 + BN + ReLU + two residual blocks at 64 channels) on the reader's pillar list."""
    import torch.nn.functional as F
    from test_gpu_pillarnet import make_points
    rng = np.random.default_rng(13)
    xyz, cnt = make_points(2, 3000, seed=9)
    g = po.gen_indice_pairs_flat(torch.from_numpy(xyz).to(cuda), torch.from_numpy(cnt).to(cuda), 0.16, (320, 320))
    pillars, bev = g["pillars"], g["pillar_bev_indices"]
    M = int(pillars.shape[0])
    feats = np.abs(rng.normal(size=(M, 32))).astype(np.float32)
    torch.manual_seed(1)
    enc = po.PillarEncoderConv2(32, 64).to(cuda).eval()
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.load_state_dict(_bn(rng, 64, cuda).state_dict())
    assert {"0.weight", "1.running_mean", "3.conv1.0.weight", "3.conv1.0.bias", "4.conv2.1.weight"} <= set(enc.state_dict().keys())
    assert "0.bias" not in enc.state_dict()
    got, out_pillars, out_bev = enc(torch.from_numpy(feats).to(cuda), pillars, bev)
    # oracle composition
    op, ob, nbr2 = pb.sparse_conv_s2_indices(bev.cpu().numpy(), pillars.cpu().numpy())
    assert np.array_equal(out_pillars.cpu().numpy(), op) and np.array_equal(out_bev.cpu().numpy(), ob)
    conv, bn = getattr(enc, "0"), getattr(enc, "1")
    x = pb.subm_conv3x3(feats, nbr2, conv.weight.detach().cpu().numpy(), bn=_bn_np(bn), relu=True)
    nbr = pb.subm_neighbors(ob, op)
    def cb(seq, t, res):
        return pb.subm_conv3x3(t, nbr, seq[0].weight.detach().cpu().numpy(), bias=seq[0].bias.detach().cpu().numpy(), bn=_bn_np(seq[1]),
                               residual=res, relu=True)
    for name in ("3", "4"):
        blk = getattr(enc, name)
        x = cb(blk.conv2, cb(blk.conv1, x, None), x)
    assert got.shape == x.shape and _close(got.cpu().numpy(), x)
    # the stride-2 layer alone against cuDNN's dense fp32 convolution at the active output cells
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        dense = po.sparse_to_dense(torch.from_numpy(feats).to(cuda), pillars, (320, 320), 2)
        ref = F.conv2d(dense, conv.weight.detach().permute(0, 3, 1, 2).contiguous(), None, stride=2, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    idx = out_pillars.long()
    ref = ref[idx[:, 0], :, idx[:, 1], idx[:, 2]]
    r = po.sparse_conv_s2_indices(bev, pillars)
    one = po.subm_conv3x3(torch.from_numpy(feats).to(cuda), r["neighbors"], conv.weight.detach())
    assert (one - ref).abs().max().item() <= TOL * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("geom", ["vod", "tj4d"])
def test_reader_to_conv1_conv2_to_dense_end_to_end(cuda, geom):
    """Path B behind the reader: PillarMaxPooling (native) -> conv1 -> conv2 on the pillar list -> dense [B,64,H/2,W/2], against the
    dense torch composition of the same modules (cuDNN fp32) fed with the reader's densified output.  TJ4D = the non-square
    432 x 496 grid of hgsfusion_tj4d.yaml: spatial_shape must come out as (Ny, Nx) = (496, 432) (pillar_modules.py:82)."""
    import torch.nn.functional as F
    from test_gpu_pillarnet import PATHB_GEOM, make_points
    gm = PATHB_GEOM[geom]
    rng = np.random.default_rng(31)
    xyz, cnt = make_points(2, 4000, seed=3, H=gm["Ny"], W=gm["Nx"])
    N = xyz.shape[0]
    feat = rng.normal(size=(N, gm["Cf"])).astype(np.float32)
    torch.manual_seed(2)
    reader = po.PillarMaxPooling([gm["Cf"] + 6, 32], 0.16, gm["pc_range"]).to(cuda).eval()
    c1, c2 = po.PillarEncoderConv1(32).to(cuda).eval(), po.PillarEncoderConv2(32, 64).to(cuda).eval()
    for mod in (reader, c1, c2):
        for m in mod.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.load_state_dict(_bn(rng, m.num_features, cuda).state_dict())
    with torch.no_grad():
        out = reader(torch.from_numpy(xyz).to(cuda), torch.from_numpy(cnt).to(cuda), torch.from_numpy(feat).to(cuda))
        pf, pillars, (H, W), B = out if isinstance(out, tuple) else (out.features, out.indices, out.spatial_shape, out.batch_size)
        bev = reader.pillar_bev_indices
        assert (H, W) == (gm["Ny"], gm["Nx"])
        assert bev is not None and tuple(bev.shape) == (B, H, W)
        x1 = c1(pf, pillars, bev)
        x2, p2, bev2 = c2(x1, pillars, bev)
        got = po.sparse_to_dense(x2, p2, (bev2.shape[1], bev2.shape[2]), B)
        assert tuple(got.shape) == (B, 64, (H + 1) // 2, (W + 1) // 2)
        # dense composition
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            idx = pillars.long()
            mask = torch.zeros((B, 1, H, W), device=cuda); mask[idx[:, 0], 0, idx[:, 1], idx[:, 2]] = 1
            mask2 = (F.max_pool2d(mask, 3, 2, 1) > 0).float()
            def cb(seq, x, res, msk, stride=1):
                conv, bn = seq
                y = F.conv2d(x, conv.weight.permute(0, 3, 1, 2).contiguous(), conv.bias, stride=stride, padding=1)
                y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
                if res is not None: y = y + res
                return torch.relu(y) * msk
            x = po.sparse_to_dense(pf, pillars, (H, W), B)
            b0, b1 = getattr(c1, "0"), getattr(c1, "1")
            i = cb(b0.conv0, x, None, mask); x = cb(b0.conv2, cb(b0.conv1, i, None, mask), i, mask)
            x = cb(b1.conv2, cb(b1.conv1, x, None, mask), x, mask)
            x = cb((getattr(c2, "0"), getattr(c2, "1")), x, None, mask2, stride=2)
            for name in ("3", "4"):
                blk = getattr(c2, name)
                x = cb(blk.conv2, cb(blk.conv1, x, None, mask2), x, mask2)
        finally:
            torch.backends.cudnn.allow_tf32 = old
    assert got.shape == x.shape
    assert (got - x).abs().max().item() <= 5e-5 * max(1.0, x.abs().max().item())      # eleven fp32 layers deep


def test_whole_sparse_encoder_conv1_to_conv4_against_dense_cudnn(cuda):
    """SpMiddlePillarEncoder18 (pcnres18.py:200-285) on pillar lists, all four stages (32 -> 32, 64, 128, 256 channels at strides
    1 / 2 / 4 / 8), against the dense torch composition of the same modules (cuDNN fp32, masked to the active sets) on the same
    GPU; the module tree carries the reference's parameter names (conv1.0.conv0.0.weight ... conv4.4.conv2.1.running_var)."""
    import torch.nn.functional as F
    rng = np.random.default_rng(77)
    B, H, W = 2, 96, 80
    pillars_np, bev_np = _random_pillars(rng, B, H, W, 900)
    pillars, bev = torch.from_numpy(pillars_np).to(cuda), torch.from_numpy(bev_np).to(cuda)
    M = pillars.shape[0]
    feats = torch.from_numpy(np.abs(rng.normal(size=(M, 32))).astype(np.float32)).to(cuda)
    torch.manual_seed(5)
    enc = po.SpMiddlePillarEncoder18(32, out_indices=(0, 1, 2, 3)).to(cuda).eval()
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.load_state_dict(_bn(rng, m.num_features, cuda).state_dict())
    keys = set(enc.state_dict().keys())
    assert {"conv1.0.conv0.0.weight", "conv2.0.weight", "conv3.0.weight", "conv3.3.conv1.0.bias", "conv4.4.conv2.1.running_var"} <= keys
    assert enc.state_dict()["conv4.0.weight"].shape == (256, 3, 3, 128) and "conv3.0.bias" not in keys
    with torch.no_grad():
        outs = enc(feats, pillars, bev)
        assert [o[0].shape[1] for o in outs] == [32, 64, 128, 256]
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            idx = pillars.long()
            mask = torch.zeros((B, 1, H, W), device=cuda); mask[idx[:, 0], 0, idx[:, 1], idx[:, 2]] = 1

            def cb(seq, x, res, msk, stride=1):
                conv, bn = seq
                y = F.conv2d(x, conv.weight.permute(0, 3, 1, 2).contiguous(), conv.bias, stride=stride, padding=1)
                y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
                if res is not None: y = y + res
                return torch.relu(y) * msk
            x = po.sparse_to_dense(feats, pillars, (H, W), B)
            b0, b1 = getattr(enc.conv1, "0"), getattr(enc.conv1, "1")
            i = cb(b0.conv0, x, None, mask); x = cb(b0.conv2, cb(b0.conv1, i, None, mask), i, mask)
            x = cb(b1.conv2, cb(b1.conv1, x, None, mask), x, mask)
            dense = [x]
            for stage in (enc.conv2, enc.conv3, enc.conv4):
                mask = (F.max_pool2d(mask, 3, 2, 1) > 0).float()
                x = cb((getattr(stage, "0"), getattr(stage, "1")), x, None, mask, stride=2)
                for name in ("3", "4"):
                    blk = getattr(stage, name)
                    x = cb(blk.conv2, cb(blk.conv1, x, None, mask), x, mask)
                dense.append(x)
        finally:
            torch.backends.cudnn.allow_tf32 = old
        for (f, pl, bv), ref in zip(outs, dense):
            got = po.sparse_to_dense(f, pl, (bv.shape[1], bv.shape[2]), B)
            assert got.shape == ref.shape
            assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())      # up to 21 fp32 layers deep, K up to 2304
            assert int((mask_of(got) != mask_of(ref)).sum()) == 0


def mask_of(t):
    return (t != 0).any(dim=1)
