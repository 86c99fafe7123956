"""Host-side logic: geometry arithmetic, module construction / state_dict names, sharding, synthetic inputs.
Everything here runs without a GPU; nothing computes on the oracle's behalf of the product."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from hgsfusion_b200 import modules, sharding, synthetic
from hgsfusion_b200.geometry import grid_size, make_geometry
from oracle import oracle


def cfg_ns(**kw):
    base = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64])
    base.update(kw)
    return SimpleNamespace(**base)


@pytest.mark.parametrize("name", list(synthetic.CONFIGS))
def test_geometry_matches_oracle_and_reference_expressions(name):
    c = synthetic.CONFIGS[name]
    rng = np.array(c["pc_range"], dtype=np.float32)
    og = oracle.Geometry(c["pc_range"], c["voxel_size"])
    assert np.array_equal(grid_size(rng, c["voxel_size"]), og.grid)
    g = make_geometry(rng, c["voxel_size"])
    assert list(g.grid) == list(og.grid)
    # pillar_vfe.py:79-81 evaluated with an np.float32 range, as OpenPCDet passes it
    assert tuple(g.centre_off) == og.centre_offsets()
    assert list(g.voxel_size) == [float(np.float32(v)) for v in c["voxel_size"]]


def test_centre_offset_follows_the_callers_range_dtype():
    vs = [0.16, 0.16, 5]
    g32 = make_geometry(np.array([0, -25.6, -3, 51.2, 25.6, 2], dtype=np.float32), vs)
    g64 = make_geometry([0, -25.6, -3, 51.2, 25.6, 2], vs)
    assert g64.centre_off[1] == float(np.float32(0.16 / 2 + -25.6))
    assert g32.centre_off[1] == float(np.float32(0.16 / 2 + np.float32(-25.6)))


def test_pillar_vfe_module_has_the_reference_parameter_names():
    m = modules.PillarVFE(model_cfg=cfg_ns(), num_point_features=7, voxel_size=[0.16, 0.16, 5],
                          point_cloud_range=np.array([0, -25.6, -3, 51.2, 25.6, 2], dtype=np.float32))
    keys = set(m.state_dict().keys())
    # pillar_vfe.py:22-23,74
    assert keys == {"pfn_layers.0.linear.weight", "pfn_layers.0.norm.weight", "pfn_layers.0.norm.bias",
                    "pfn_layers.0.norm.running_mean", "pfn_layers.0.norm.running_var",
                    "pfn_layers.0.norm.num_batches_tracked"}
    assert tuple(m.state_dict()["pfn_layers.0.linear.weight"].shape) == (64, 13)
    assert m.get_output_feature_dim() == 64
    m2 = modules.PillarVFE(model_cfg=cfg_ns(USE_NORM=False, USE_ABSLOTE_XYZ=False, WITH_DISTANCE=True), num_point_features=8,
                           voxel_size=[0.16, 0.16, 6], point_cloud_range=np.array([0, -39.68, -4, 69.12, 39.68, 2], dtype=np.float32))
    assert set(m2.state_dict().keys()) == {"pfn_layers.0.linear.weight", "pfn_layers.0.linear.bias"}
    assert tuple(m2.state_dict()["pfn_layers.0.linear.weight"].shape) == (64, 8 + 3 + 1)


def test_module_guards():
    # a stacked PFN builds the reference's parameter tree (pillar_vfe.py:18-19,63-74): the non-last layer halves its width
    st = modules.PillarVFE(model_cfg=cfg_ns(NUM_FILTERS=[64, 128]), num_point_features=7, voxel_size=[0.16, 0.16, 5],
                           point_cloud_range=np.array([0, -25.6, -3, 51.2, 25.6, 2], dtype=np.float32))
    sd = st.state_dict()
    assert sd["pfn_layers.0.linear.weight"].shape == (32, 13) and sd["pfn_layers.1.linear.weight"].shape == (128, 64)
    assert st.stacked and not st.fused_ok and st.get_output_feature_dim() == 128
    with pytest.raises(NotImplementedError):                      # three layers: not built
        modules.PillarVFE(model_cfg=cfg_ns(NUM_FILTERS=[64, 64, 64]), num_point_features=7, voxel_size=[0.16, 0.16, 5],
                          point_cloud_range=np.array([0, -25.6, -3, 51.2, 25.6, 2], dtype=np.float32))
    m = modules.FusedPillarVFE(model_cfg=cfg_ns(MAX_POINTS_PER_VOXEL=32, MAX_NUMBER_OF_VOXELS={'train': 16000, 'test': 40000}),
                               num_point_features=7, voxel_size=[0.16, 0.16, 5],
                               point_cloud_range=np.array([0, -25.6, -3, 51.2, 25.6, 2], dtype=np.float32))
    m.train()
    assert m._path().max_voxels == 16000                         # MAX_NUMBER_OF_VOXELS['train'] (data_processor.py:146)
    with pytest.raises(ValueError, match="CUDA tensor"):         # train mode goes native too: CPU tensors are refused
        m({'points': torch.zeros(1, 8), 'batch_size': 1})
    m.eval()
    assert m._path().max_voxels == 40000
    with pytest.raises(ValueError, match="CUDA tensor"):
        with torch.no_grad():
            m({'points': torch.zeros(4, 8), 'batch_size': 1})     # CPU tensors are refused: no CPU path
    sc = modules.PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=[320, 320, 1])
    assert sc.num_bev_features == 64
    with pytest.raises(AssertionError):
        modules.PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=[320, 320, 2])
    with pytest.raises(KeyError):
        modules.PillarScatterPassthrough(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=[320, 320, 1])({})


def test_registry_hook():
    vfe, m2b = {}, {}
    modules.register(vfe, m2b, override=True)
    assert vfe["PillarVFE"] is modules.PillarVFE and vfe["FusedPillarVFE"] is modules.FusedPillarVFE
    assert m2b["PointPillarScatter"] is modules.PointPillarScatter


def test_frame_range_partitions_exactly():
    for B in (1, 7, 16, 64):
        for W in (1, 2, 3, 4, 8):
            spans = [sharding.frame_range(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_range(4, 2, 2)


def test_shard_points_rebases_batch_column():
    pts, offs = synthetic.make_batch("vod", 5, 50, "uniform")
    sp, so = sharding.shard_points(pts, offs, 2, 4)
    assert sp.shape[0] == 100 and list(so) == [0, 50, 100]
    assert set(np.unique(sp[:, 0])) == {0.0, 1.0}
    assert np.array_equal(sp[:, 1:], pts[100:200, 1:])


def test_synthetic_is_deterministic_and_shaped():
    a, oa = synthetic.make_batch("tj4d", 2, 300, "clustered", seed0=3)
    b, ob = synthetic.make_batch("tj4d", 2, 300, "clustered", seed0=3)
    assert np.array_equal(a, b) and np.array_equal(oa, ob)
    assert a.shape == (600, 9) and a.dtype == np.float32 and list(oa) == [0, 300, 600]
    w = synthetic.make_pfn(14, 64)
    assert w.weight.shape == (64, 14) and (w.running_var > 0).all()
