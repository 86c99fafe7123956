"""PINS THE VOXELIZER.  Run this on any machine where `spconv` imports (it is not installed in the build image, and the
reference pins no version: setup.py:48, docs/INSTALL.md:9,30-33):

    python tests/golden/make_spconv_golden.py          # writes tests/golden/spconv_v<major>_*.npz

It calls spconv exactly the way the reference's VoxelGeneratorWrapper does (pcdet/datasets/processor/data_processor.py:16-61:
spconv 1.x `VoxelGeneratorV2` / `VoxelGenerator(voxel_size, point_cloud_range, max_num_points, max_voxels).generate(points)`,
spconv 2.x `Point2VoxelCPU3d(vsize_xyz, coors_range_xyz, num_point_features, max_num_points_per_voxel,
max_num_voxels).point_to_voxel(tv.from_numpy(points))`) on the hand-computed cases of voxelize_cases.json and on seeded
frames of the four BASELINE configs (including frames that overflow max_voxels and pillars beyond max_points), and stores
inputs + spconv's outputs.  tests/test_oracle_spconv_golden.py then checks the oracle -- and, on a GPU, the kernels --
against those files bit for bit; commit the .npz files and "parity unpinned" becomes "pinned against spconv <version>".
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from hgsfusion_b200 import synthetic  # noqa: E402


def make_generator(vsize, pc_range, F, P, max_voxels):
    """-> (major version, generate(points) -> voxels, coords (z,y,x), num_points); import order as data_processor.py:18-26."""
    try:
        from spconv.utils import VoxelGeneratorV2 as VoxelGenerator
        ver = 1
    except Exception:
        try:
            from spconv.utils import VoxelGenerator
            ver = 1
        except Exception:
            from spconv.utils import Point2VoxelCPU3d as VoxelGenerator
            ver = 2
    if ver == 1:
        gen = VoxelGenerator(voxel_size=vsize, point_cloud_range=pc_range, max_num_points=P, max_voxels=max_voxels)

        def generate(points):
            out = gen.generate(points)
            if isinstance(out, dict):
                return out["voxels"], out["coordinates"], out["num_points_per_voxel"]
            return out
    else:
        from cumm import tensorview as tv
        gen = VoxelGenerator(vsize_xyz=vsize, coors_range_xyz=pc_range, num_point_features=F, max_num_points_per_voxel=P,
                             max_num_voxels=max_voxels)

        def generate(points):
            v, c, n = gen.point_to_voxel(tv.from_numpy(points))
            return v.numpy(), c.numpy(), n.numpy()
    return ver, generate


def cases():
    d = json.load(open(os.path.join(HERE, "voxelize_cases.json")))
    for c in d["cases"]:
        pts = np.asarray(c["points"], dtype=np.float32).reshape(-1, 4)
        yield "hand_" + c["name"], d["pc_range"], d["voxel_size"], 4, c["P"], c["max_voxels"], pts
    for name, config, n, P, mv, mode in [("cfg1_vod_2000", "vod", 2000, 32, 40000, "clustered"),
                                         ("cfg2_vod_30000", "vod", 30000, 32, 40000, "clustered"),
                                         ("cfg2_vod_uniform", "vod", 30000, 32, 40000, "uniform"),
                                         ("cfg2_vod_overflow", "vod", 30000, 5, 3000, "clustered"),
                                         ("cfg3_tj4d", "tj4d", 30000, 32, 40000, "clustered"),
                                         ("cfg3_tj4d_overflow", "tj4d", 30000, 10, 1000, "uniform"),
                                         ("cfg4_stress", "stress", 200000, 32, 40000, "clustered"),
                                         ("cfg4_stress_p100", "stress", 200000, 100, 40000, "clustered")]:
        cfg = synthetic.CONFIGS[config]
        yield name, cfg["pc_range"], cfg["voxel_size"], cfg["F"], P, mv, synthetic.make_frame(n, cfg["pc_range"], cfg["F"], 7, mode, 0.02)


def main():
    import spconv
    version = getattr(spconv, "__version__", "unknown")
    for name, pc_range, vsize, F, P, mv, pts in cases():
        ver, generate = make_generator(list(vsize), np.asarray(pc_range, dtype=np.float32), F, P, mv)
        voxels, coords, num = generate(pts)
        path = os.path.join(HERE, f"spconv_v{ver}_{name}.npz")
        np.savez_compressed(path, points=pts, pc_range=np.asarray(pc_range, dtype=np.float32), voxel_size=np.asarray(vsize, dtype=np.float64),
                            P=P, max_voxels=mv, voxels=np.asarray(voxels), coords=np.asarray(coords), num_points=np.asarray(num),
                            spconv_version=str(version), spconv_major=ver)
        print(f"{name:24s} spconv {version}: {len(num)} pillars -> {os.path.basename(path)}")


if __name__ == "__main__":
    main()
