"""ctypes binding of include/hgsfusion_b200.h.

The product path is the CUDA library and nothing else: if libhgsfusion_b200.so is missing this
module raises -- there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HGSF_LIB: experiment builds of the same library (scripts/variants.sh); the product always loads the in-tree default
LIB_PATH = os.environ.get("HGSF_LIB") or os.path.join(_HERE, "libhgsfusion_b200.so")

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_DRIVER = 0, -1, -2, -3, -4

EXPORTS = ["hgsf_abi_version", "hgsf_status_string", "hgsf_pillar_capacity", "hgsf_workspace_size",
           "hgsf_pillarize", "hgsf_points_to_bev", "hgsf_pillar_vfe", "hgsf_pillar_vfe_stacked", "hgsf_scatter_workspace_size",
           "hgsf_pointpillar_scatter", "hgsf_last_launch_count", "hgsf_emit_timing_begin", "hgsf_emit_timing_collect",
           "hgsf_pillarnet_workspace_size", "hgsf_pillarnet_indices", "hgsf_gather_feature", "hgsf_gather_feature_grad",
           "hgsf_scatter_max", "hgsf_scatter_max_grad", "hgsf_split_encode", "hgsf_pillarnet_reader",
           "hgsf_train_stats_doubles", "hgsf_train_scratch_doubles", "hgsf_pillar_vfe_batch_stats",
           "hgsf_pillar_vfe_backward", "hgsf_pointpillar_scatter_backward", "hgsf_points_to_bev_train",
           "hgsf_points_to_bev_train_backward", "hgsf_hybrid_workspace_size",
           "hgsf_assemble_hybrid_points", "hgsf_sparse_to_dense_workspace_size", "hgsf_sparse_to_dense",
           "hgsf_subm_neighbors", "hgsf_subm_conv3x3", "hgsf_sparse_conv_s2_workspace_size",
           "hgsf_sparse_conv_s2_indices"]


class Geometry(C.Structure):
    _fields_ = [("pc_range", C.c_float * 6), ("voxel_size", C.c_float * 3), ("grid", C.c_int32 * 3),
                ("centre_off", C.c_float * 3)]


class Points(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_int64), ("stride", C.c_int32), ("xyz_col", C.c_int32),
                ("num_features", C.c_int32), ("batch_col", C.c_int32), ("frame_offsets", C.c_void_p),
                ("batch_size", C.c_int32), ("flags", C.c_int32)]


POINTS_SPCONV1_BREAK, POINTS_FLIP_X, POINTS_FLIP_Y = 1, 2, 4


class Pfn(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("bn_weight", C.c_void_p), ("bn_bias", C.c_void_p),
                ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p), ("bn_eps", C.c_float),
                ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("use_absolute_xyz", C.c_int32),
                ("with_distance", C.c_int32)]


class PillarOutputs(C.Structure):
    _fields_ = [("voxel_coords", C.c_void_p), ("voxel_num_points", C.c_void_p), ("voxels", C.c_void_p),
                ("pillar_features", C.c_void_p), ("spatial_features", C.c_void_p), ("num_pillars", C.c_void_p),
                ("pillar_capacity", C.c_int64)]


class HybridInputs(C.Structure):
    _fields_ = [("real", C.c_void_p), ("gt_real", C.c_void_p), ("virt", C.c_void_p), ("real_offsets", C.c_void_p),
                ("gt_offsets", C.c_void_p), ("virt_offsets", C.c_void_p), ("n_candidates", C.c_int64),
                ("real_features", C.c_int32), ("hybrid_features", C.c_int32), ("batch_size", C.c_int32),
                ("no_dup", C.c_int32), ("dup_threshold", C.c_double)]


class SubmConv(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("weight_layout", C.c_int32), ("bias", C.c_void_p), ("bn_weight", C.c_void_p),
                ("bn_bias", C.c_void_p), ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p), ("bn_eps", C.c_float),
                ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("relu", C.c_int32)]


class HgsfError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        super().__init__(f"{where}: status {status} ({status_string(status)})")


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m hgsfusion_b200.build` (nvcc, sm_100a). "
            "hgsfusion_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.hgsf_abi_version.restype = C.c_int
    lib.hgsf_status_string.restype = C.c_char_p
    lib.hgsf_status_string.argtypes = [C.c_int]
    lib.hgsf_pillar_capacity.restype = C.c_int64
    lib.hgsf_pillar_capacity.argtypes = [C.POINTER(Geometry), C.c_int64, C.c_int32, C.c_int32]
    lib.hgsf_workspace_size.argtypes = [C.POINTER(Geometry), C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    path_args = [C.POINTER(Geometry), C.POINTER(Points)]
    tail = [C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(PillarOutputs), C.c_void_p]
    lib.hgsf_pillarize.argtypes = path_args + tail
    lib.hgsf_points_to_bev.argtypes = path_args + [C.POINTER(Pfn)] + tail
    lib.hgsf_pillar_vfe.argtypes = [C.POINTER(Geometry), C.POINTER(Pfn), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    if hasattr(lib, "hgsf_pillar_vfe_stacked"):        # (absent only from experiment builds of older trees, see below)
        lib.hgsf_pillar_vfe_stacked.argtypes = [C.POINTER(Geometry), C.POINTER(Pfn), C.POINTER(Pfn), C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.hgsf_scatter_workspace_size.argtypes = [C.POINTER(Geometry), C.c_int32, C.POINTER(C.c_size_t)]
    lib.hgsf_pointpillar_scatter.argtypes = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                             C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.hgsf_pillarnet_workspace_size.argtypes = [C.c_int64, C.POINTER(C.c_size_t)]
    lib.hgsf_pillarnet_indices.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_size_t, C.c_void_p]
    lib.hgsf_gather_feature.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.hgsf_gather_feature_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.hgsf_scatter_max.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hgsf_scatter_max_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    lib.hgsf_split_encode.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hgsf_pillarnet_reader.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_float, C.c_float, C.POINTER(Pfn), C.c_void_p, C.c_void_p]
    lib.hgsf_train_stats_doubles.restype = C.c_int64
    lib.hgsf_train_stats_doubles.argtypes = [C.c_int32, C.c_int32]
    lib.hgsf_train_scratch_doubles.restype = C.c_int64
    lib.hgsf_train_scratch_doubles.argtypes = [C.c_int32, C.c_int32]
    contract = [C.POINTER(Geometry), C.POINTER(Pfn), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                C.c_int32, C.c_int32]
    lib.hgsf_pillar_vfe_batch_stats.argtypes = contract + [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                           C.c_void_p]
    lib.hgsf_pillar_vfe_backward.argtypes = contract + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                        C.c_void_p]
    lib.hgsf_pointpillar_scatter_backward.argtypes = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                                      C.c_int32, C.c_void_p, C.c_void_p]
    if hasattr(lib, "hgsf_points_to_bev_train"):
        lib.hgsf_points_to_bev_train.argtypes = path_args + [C.POINTER(Pfn), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                                                             C.POINTER(PillarOutputs), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                                             C.c_void_p, C.c_void_p, C.c_void_p]
        lib.hgsf_points_to_bev_train_backward.argtypes = [C.POINTER(Geometry), C.POINTER(Pfn), C.c_void_p, C.c_void_p, C.c_void_p,
                                                          C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                          C.c_void_p, C.c_void_p]
    lib.hgsf_hybrid_workspace_size.argtypes = [C.c_int64, C.POINTER(C.c_size_t)]
    lib.hgsf_assemble_hybrid_points.argtypes = [C.POINTER(HybridInputs), C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_size_t,
                                                C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hgsf_sparse_to_dense_workspace_size.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    lib.hgsf_sparse_to_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_size_t, C.c_void_p, C.c_void_p]
    lib.hgsf_subm_neighbors.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p]
    lib.hgsf_subm_conv3x3.argtypes = [C.POINTER(SubmConv), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    lib.hgsf_sparse_conv_s2_workspace_size.argtypes = [C.c_int64, C.c_int32, C.POINTER(C.c_size_t)]
    lib.hgsf_sparse_conv_s2_indices.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t,
                                                C.c_void_p]
    lib.hgsf_emit_timing_begin.argtypes = [C.c_int]
    lib.hgsf_emit_timing_collect.argtypes = [C.POINTER(C.c_float), C.c_int]
    for name in EXPORTS:
        if os.environ.get("HGSF_LIB") and not hasattr(lib, name):
            continue                   # an experiment build of an older source tree (scripts/variants.sh): tolerate missing entry points
        getattr(lib, name)
    if lib.hgsf_abi_version() != 1:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.hgsf_abi_version()} != 1")
    _lib = lib
    return lib


def status_string(status: int) -> str:
    return load().hgsf_status_string(int(status)).decode()


def check(status: int, where: str):
    if status != OK:
        raise HgsfError(status, where)
