"""Generates tests/golden/hybrid_*.npz by EXECUTING the reference's own statements for the hybrid point assembly
 on seeded synthetic inputs.  Runs only in the build container.

`import pcdet` fails here, and the assembly is inline code of `__getitem__`, not a function.  So the reference
sources are parsed with `ast` where they lie under /root/reference and the relevant nodes are compiled and run unchanged:

  * `calc_dist`                                         pcdet/datasets/kitti/vod_dataset.py:13-19
  * the `if "points" in get_item_list:` statement       vod_dataset.py:498-529 / tj4d_dataset.py:588-618
  * `get_fov_flag`                                      vod_dataset.py:181-197
  * `mask_points_by_range`                              pcdet/utils/common_utils.py:78-81
  * class `Calibration` (imported by path; numpy only)  pcdet/utils/calibration_kitti.py:23-88

with a stand-in `self` that serves the synthetic arrays where the dataset would read files.  Nothing of the reference is
copied into this repository; the committed fixtures hold inputs and the reference's outputs.

    python tests/golden/make_hybrid_golden.py [--check]
"""
import ast
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _load_calibration():
    spec = importlib.util.spec_from_file_location("ref_calibration_kitti", f"{REF}/pcdet/utils/calibration_kitti.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Calibration


def _find(tree, kind, name):
    for node in ast.walk(tree):
        if isinstance(node, kind) and getattr(node, "name", None) == name:
            return node
    raise KeyError(name)


def _compile_function(node, filename, ns):
    node.decorator_list = []
    mod = ast.Module(body=[node], type_ignores=[])
    exec(compile(mod, filename, "exec"), ns)
    return ns[node.name]


def load_reference_pieces(dataset: str):
    """-> (run_points_block(self, sample_idx, calib, img_shape) -> points, mask_points_by_range)"""
    path = f"{REF}/pcdet/datasets/kitti/{dataset}_dataset.py"
    tree = ast.parse(open(path).read(), filename=path)
    ns = {"np": np}
    if dataset == "vod":
        _compile_function(_find(tree, ast.FunctionDef, "calc_dist"), path, ns)
    cls = _find(tree, ast.ClassDef, "VODDataset" if dataset == "vod" else "TJ4DDataset")
    fov = _compile_function(_find(cls, ast.FunctionDef, "get_fov_flag"), path, ns)
    getitem = _find(cls, ast.FunctionDef, "__getitem__")
    block = None
    for node in ast.walk(getitem):
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Constant) \
                and node.test.left.value == "points":
            block = node
    assert block is not None
    code = compile(ast.Module(body=[block], type_ignores=[]), path, "exec")

    def run(self, sample_idx, calib, img_shape):
        self.get_fov_flag = fov
        loc = dict(self=self, sample_idx=sample_idx, calib=calib, img_shape=img_shape, input_dict={},
                   get_item_list=["points"])
        exec(code, ns, loc)
        return loc["input_dict"]["points"]

    cpath = f"{REF}/pcdet/utils/common_utils.py"
    ctree = ast.parse(open(cpath).read(), filename=cpath)
    mask = _compile_function(_find(ctree, ast.FunctionDef, "mask_points_by_range"), cpath, {"np": np})
    return run, mask


# ---- synthetic frames -----------------------------------------------------------------------------------------------------
def make_calib(rng):
    """A KITTI / VoD style calibration: radar x forward, y left, z up -> camera z forward; small random perturbation."""
    R = np.array([[0, -1, 0], [0, 0, -1], [1, 0, 0]], dtype=np.float64)
    a = rng.normal(0, 0.02, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    V2C = np.concatenate([Rx @ Ry @ R, rng.normal(0, 0.3, (3, 1))], 1).astype(np.float32)
    R0 = (np.eye(3) + rng.normal(0, 1e-3, (3, 3))).astype(np.float32)
    P2 = np.array([[1495.47, 0, 961.27, rng.normal(0, 40)], [0, 1495.47, 624.9, rng.normal(0, 2)], [0, 0, 1, rng.normal(0, 0.01)]],
                  dtype=np.float32)
    return dict(P2=P2, R0=R0, Tr_velo2cam=V2C)


def make_frame(rng, Fr, n_real, n_gt, n_virt, pc_range, dup_fraction=0.3):
    W = Fr + 8
    lo = np.array([pc_range[0] - 4, pc_range[1] - 4, pc_range[2]]), np.array([pc_range[3] + 4, pc_range[4] + 4, pc_range[5]])
    real = np.concatenate([rng.uniform(lo[0], lo[1], (n_real, 3)), rng.normal(0, 1, (n_real, Fr - 3))], 1).astype(np.float32)
    if n_real >= 4:
        # points exactly on the range boundary: the reference's limits are float32 too (dataset.py:26), so these are kept
        real[0, 0] = np.float32(pc_range[3]); real[1, 1] = np.float32(pc_range[1]); real[2, 1] = np.float32(pc_range[4])
        real[3, 0] = np.float32(pc_range[0])
    gt = np.concatenate([rng.uniform(lo[0], lo[1], (n_gt, 3)), rng.normal(0, 1, (n_gt, Fr - 3)),
                         np.eye(8)[rng.integers(0, 8, n_gt)]], 1).astype(np.float32).reshape(n_gt, W)
    # real points inside ground-truth masks also sit in the sweep (NO_DUP removes them): exact and near duplicates
    k = min(int(dup_fraction * n_gt), max(0, n_real - 6))
    if k > 0:
        real[4:4 + k, :Fr] = gt[:k, :Fr]
        if n_gt >= k + 2:
            real[4 + k, :3] = gt[k, :3] + np.float32(0.01)      # squared distance 3e-4 <= 0.001: still a duplicate
            real[5 + k, :3] = gt[k + 1, :3] + np.float32(0.03)  # 2.7e-3 > 0.001: kept
    virt = np.concatenate([rng.uniform(lo[0], lo[1], (n_virt, 3)), rng.normal(0, 1, (n_virt, Fr - 3)),
                           np.eye(8)[rng.integers(0, 8, n_virt)]], 1).astype(np.float32).reshape(n_virt, W)
    return real, gt, virt


CASES = [
    # name, dataset, Fr, pc_range, frames [(n_real, n_gt, n_virt)], fov, no_dup, use_virtual
    ("vod_fov_nodup", "vod", 7, [0, -25.6, -3, 51.2, 25.6, 2], [(400, 60, 1500), (350, 0, 800), (5, 3, 9), (300, 40, 0)], True, True, True),
    ("vod_fov", "vod", 7, [0, -25.6, -3, 51.2, 25.6, 2], [(500, 80, 2500), (450, 30, 1000)], True, False, True),
    ("vod_nofov", "vod", 7, [0, -25.6, -3, 51.2, 25.6, 2], [(300, 20, 700), (0, 4, 30), (200, 10, 600)], False, True, True),
    ("vod_real_only", "vod", 7, [0, -25.6, -3, 51.2, 25.6, 2], [(600, 0, 0), (550, 0, 0)], True, False, False),
    ("tj4d_fov", "tj4d", 8, [0, -39.68, -4, 69.12, 39.68, 2], [(700, 90, 2200), (650, 50, 1800), (600, 0, 0)], True, False, True),
]


def run_case(name, dataset, Fr, pc_range, frames, fov, no_dup, use_virtual, seed):
    Calibration = _load_calibration()
    run_block, mask_by_range = load_reference_pieces(dataset)
    rng = np.random.default_rng(seed)
    img_shape = np.array([1216, 1936], dtype=np.int32)
    out = dict(Fr=Fr, pc_range=np.asarray(pc_range, dtype=np.float64), fov=int(fov), no_dup=int(no_dup),
               use_virtual=int(use_virtual), image_shape=img_shape, n_frames=len(frames))
    for b, (nr, ng, nv) in enumerate(frames):
        real, gt, virt = make_frame(rng, Fr, nr, ng, nv, pc_range)
        cal = make_calib(rng)
        calib = Calibration(cal)
        self = SimpleNamespace(use_virtual_points=use_virtual, no_dup=no_dup, only_virtual=False,
                               dataset_cfg=SimpleNamespace(FOV_POINTS_ONLY=fov),
                               get_virtual_point=lambda idx, v=virt, g=gt: (v, g),
                               get_lidar=lambda idx, r=real: (r - np.zeros(Fr)) / np.ones(Fr))      # float64, as get_lidar's (points - means) / stds
        pts = run_block(self, "00000", calib, img_shape)                 # vod_dataset.py:498-529 / tj4d_dataset.py:588-618
        # data_processor.py:83-85 with self.point_cloud_range = np.array(POINT_CLOUD_RANGE, dtype=np.float32) (dataset.py:26)
        pts = pts[mask_by_range(pts, np.array(pc_range, dtype=np.float32))]
        out[f"real{b}"], out[f"gt{b}"], out[f"virt{b}"] = real, gt, virt
        out[f"P2_{b}"], out[f"R0_{b}"], out[f"V2C_{b}"] = cal["P2"], cal["R0"], cal["Tr_velo2cam"]
        out[f"points{b}"] = pts.astype(np.float32)                       # load_data_to_gpu: .float()
        assert pts.dtype == np.float64
    return out


def main():
    check = "--check" in sys.argv
    for i, case in enumerate(CASES):
        data = run_case(*case, seed=100 + i)
        path = os.path.join(HERE, f"hybrid_{case[0]}.npz")
        np.savez_compressed(path, **data)
        kept = [len(data[f"points{b}"]) for b in range(data["n_frames"])]
        print(case[0], "rows kept per frame", kept, "of", [sum(f) for f in case[4]])
        if check:
            from oracle import hybrid_oracle
            for b in range(data["n_frames"]):
                got = hybrid_oracle.assemble_frame_from_fixture(data, b)
                assert got.shape == data[f"points{b}"].shape and np.array_equal(got, data[f"points{b}"]), (case[0], b)
            print("   oracle == reference")


if __name__ == "__main__":
    main()
