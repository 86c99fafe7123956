"""The hybrid point assembly oracle (oracle/hybrid_oracle.py) against the fixtures produced by the reference's own statements
(tests/golden/make_hybrid_golden.py: vod_dataset.py:498-529, tj4d_dataset.py:588-618, get_fov_flag, mask_points_by_range,
Calibration), bit for bit."""
import glob
import os

import numpy as np
import pytest

from oracle import hybrid_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "hybrid_*.npz")))


def test_fixtures_present():
    assert len(FIXTURES) >= 5


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_the_reference(path):
    d = np.load(path)
    for b in range(int(d["n_frames"])):
        got = hybrid_oracle.assemble_frame_from_fixture(d, b)
        ref = d[f"points{b}"]
        assert got.dtype == np.float32 and got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_reference_quirks_are_kept():
    rng = np.random.default_rng(0)
    real = rng.normal(0, 1, (5, 7)).astype(np.float32)
    gt = rng.normal(0, 1, (3, 15)).astype(np.float32)
    virt = rng.normal(0, 1, (4, 15)).astype(np.float32)
    # no mask points: the virtual points are dropped, every extra column is 1 (vod_dataset.py:507-509)
    p = hybrid_oracle.assemble_frame(real, gt[:0], virt)
    assert p.shape == (5, 17) and (p[:, 7:] == 1).all() and np.array_equal(p[:, :7], real)
    # mask points but no virtual points: points[-0:, -1] = 1 hits every row (vod_dataset.py:521)
    p = hybrid_oracle.assemble_frame(real, gt, virt[:0])
    assert p.shape == (8, 17) and (p[:, 16] == 1).all() and (p[5:, 15] == 0).all() and (p[:5, 15] == 1).all()
    # the normal case: flags (1,1) / (0,0) / (0,1)
    p = hybrid_oracle.assemble_frame(real, gt, virt)
    assert [tuple(r) for r in p[[0, 5, 8]][:, 15:]] == [(1, 1), (0, 0), (0, 1)]
    # TJ4D with virtual points but no mask points raises in the reference (tj4d_dataset.py:603-605)
    with pytest.raises(ValueError):
        hybrid_oracle.assemble_frame(rng.normal(0, 1, (5, 8)).astype(np.float32), np.zeros((0, 16), np.float32),
                                     rng.normal(0, 1, (4, 16)).astype(np.float32), dataset="tj4d")


def test_range_mask_limits_are_float32():
    # The reference masks the float64 point array against DatasetTemplate.point_cloud_range = np.array(..., dtype=np.float32)
    # (dataset.py:26, data_processor.py:83-85, common_utils.py:78-81): the limits are float32(25.6), float32(-25.6),
    # float32(51.2), widened exactly, and the compare is inclusive -- points exactly on them are kept, one ulp beyond is dropped.
    real = np.zeros((6, 7), np.float32)
    real[0, 1] = np.float32(25.6); real[1, 1] = np.float32(-25.6); real[2, 0] = np.float32(51.2)
    real[3, 0] = np.nextafter(np.float32(51.2), np.float32(0))
    real[4, 0] = np.nextafter(np.float32(51.2), np.float32(100))          # outside
    real[5, 1] = np.nextafter(np.float32(-25.6), np.float32(-100))        # outside
    p = hybrid_oracle.assemble_frame(real, None, None, use_virtual=False, pc_range=[0, -25.6, -3, 51.2, 25.6, 2])
    assert len(p) == 4 and np.array_equal(p, real[:4])
    # the same through the reference's own statement
    import ast
    src = open("/root/reference/pcdet/utils/common_utils.py").read() if __import__("os").path.exists("/root/reference") else None
    if src is not None:
        tree = ast.parse(src)
        fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "mask_points_by_range")
        ns = {"np": np}
        exec(compile(ast.Module(body=[fn], type_ignores=[]), "common_utils.py", "exec"), ns)
        m = ns["mask_points_by_range"](real.astype(np.float64), np.array([0, -25.6, -3, 51.2, 25.6, 2], dtype=np.float32))
        assert m.tolist() == [True, True, True, True, False, False]
