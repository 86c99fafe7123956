"""Shared helpers for the parity tests: the oracle is the checker, the CUDA library the thing checked."""
from __future__ import annotations

import numpy as np

from hgsfusion_b200 import synthetic
from oracle import oracle


def geom_for(config: str) -> oracle.Geometry:
    c = synthetic.CONFIGS[config]
    return oracle.Geometry(c["pc_range"], c["voxel_size"])


def oracle_pfn(w) -> oracle.PfnParams:
    return oracle.PfnParams(w.weight, w.gamma, w.beta, w.running_mean, w.running_var, bias=w.bias, eps=w.eps)


def device_pfn(w, device, use_absolute_xyz=True, with_distance=False):
    import torch
    from hgsfusion_b200.ops import PfnWeights
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return PfnWeights(weight=t(w.weight), bn_weight=t(w.gamma), bn_bias=t(w.beta), running_mean=t(w.running_mean),
                      running_var=t(w.running_var), bias=t(w.bias), eps=w.eps,
                      use_absolute_xyz=use_absolute_xyz, with_distance=with_distance)


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def features_close(got: np.ndarray, ref: np.ndarray, rtol: float = 1e-5, atol_frac: float = 1e-6) -> bool:
    """|got - ref| <= rtol*|ref| + atol with atol = atol_frac * max|ref|.

    rtol is BASELINE.json's bar ("within 1e-5 relative for fp32 features").  The absolute term only
    matters for outputs that are themselves the result of cancellation to ~0 in the BatchNorm (a value
    of 1e-4 produced from O(1) terms): there the reference's own MKL-sqrt invstd, 1 ulp off the IEEE
    value, already moves the result by ~1e-8.  SURVEY.md section 7 allows atol = 1e-5*max|ref|; this uses a
    ten times tighter 1e-6."""
    got = got.astype(np.float64)
    ref = ref.astype(np.float64)
    if ref.size == 0:
        return got.size == 0
    atol = atol_frac * np.abs(ref).max()
    return bool((np.abs(got - ref) <= rtol * np.abs(ref) + atol).all())
