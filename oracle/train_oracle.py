"""Numpy restatement of TRAINING through the reference's PillarVFE / PointPillarScatter (TEST INFRASTRUCTURE, like the
rest of oracle/): PFNLayer.forward in train mode (pcdet/models/backbones_3d/vfe/pillar_vfe.py:29-49) -- Linear,
BatchNorm1d with batch statistics over all M*P rows (zero-padded rows included), ReLU, max over P -- and what torch
autograd computes for it.  float64 arithmetic; pinned by tests/golden/train_*.npz, which hold the outputs and
gradients of the reference class itself (tests/golden/make_golden.py, torch CPU autograd).
"""
import numpy as np


def decorate(voxels, coords, num, pc_range, voxel_size, use_absolute_xyz=True, with_distance=False):
    """PillarVFE.forward up to the PFN input (pillar_vfe.py:94-118) -> [M, P, Cin] float32."""
    v = np.asarray(voxels, dtype=np.float32)
    M, P, F = v.shape
    num = np.asarray(num).astype(np.float32)
    c = np.asarray(coords).astype(np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    off = (np.asarray(voxel_size, dtype=np.float64) / 2 + np.asarray(pc_range[:3], dtype=np.float64)).astype(np.float32)
    mean = v[:, :, :3].sum(axis=1, keepdims=True, dtype=np.float32) / num.reshape(-1, 1, 1)
    f_cluster = v[:, :, :3] - mean
    f_center = np.empty_like(f_cluster)
    f_center[:, :, 0] = v[:, :, 0] - (c[:, 3:4] * vs[0] + off[0])
    f_center[:, :, 1] = v[:, :, 1] - (c[:, 2:3] * vs[1] + off[1])
    f_center[:, :, 2] = v[:, :, 2] - (c[:, 1:2] * vs[2] + off[2])
    parts = [v if use_absolute_xyz else v[:, :, 3:], f_cluster, f_center]
    if with_distance:
        parts.append(np.linalg.norm(v[:, :, :3], axis=2, keepdims=True).astype(np.float32))
    feats = np.concatenate(parts, axis=2)
    mask = (num.astype(np.int32).reshape(-1, 1) > np.arange(P).reshape(1, -1)).astype(np.float32)
    return feats * mask[:, :, None]


def pfn_train_forward(feats, W, gamma, beta, eps=1e-3):
    """-> out [M,C], cache.  Batch statistics are biased (what the forward normalises with)."""
    f = feats.astype(np.float64)
    x = f @ np.asarray(W, dtype=np.float64).T                        # [M,P,C]
    N = x.shape[0] * x.shape[1]
    mean = x.reshape(N, -1).mean(axis=0)
    var = x.reshape(N, -1).var(axis=0)
    invstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * invstd
    y = xhat * np.asarray(gamma, dtype=np.float64) + np.asarray(beta, dtype=np.float64)
    z = np.maximum(y, 0.0)
    arg = z.argmax(axis=1)                                           # first maximum, like torch.max on the CPU
    out = np.take_along_axis(z, arg[:, None, :], axis=1)[:, 0, :]
    return out, dict(f=f, x=x, xhat=xhat, y=y, arg=arg, mean=mean, var=var, invstd=invstd, N=N)


def running_update(running_mean, running_var, mean, var, N, momentum=0.01):
    unbiased = var * N / (N - 1) if N > 1 else var
    return ((1 - momentum) * np.asarray(running_mean, np.float64) + momentum * mean,
            (1 - momentum) * np.asarray(running_var, np.float64) + momentum * unbiased)


def pfn_backward(cache, gamma, grad_out, batch_stats=True):
    """d out [M,C] -> (dW [C,Cin], dgamma [C], dbeta [C]).  batch_stats=False: BN statistics are constants."""
    f, xhat, y, arg, invstd, N = cache["f"], cache["xhat"], cache["y"], cache["arg"], cache["invstd"], cache["N"]
    M, P, C = y.shape
    dz = np.zeros_like(y)
    np.put_along_axis(dz, arg[:, None, :], np.asarray(grad_out, dtype=np.float64)[:, None, :], axis=1)
    dy = dz * (y > 0)
    dbeta = dy.sum(axis=(0, 1))
    dgamma = (dy * xhat).sum(axis=(0, 1))
    g = np.asarray(gamma, dtype=np.float64)
    if batch_stats:
        dx = g * invstd * (dy - dbeta / N - xhat * dgamma / N)
    else:
        dx = g * invstd * dy
    dW = np.einsum("mpc,mpk->ck", dx, f)
    return dW, dgamma, dbeta


def scatter_backward(grad_canvas, coords):
    c = np.asarray(coords).astype(np.int64)
    return np.asarray(grad_canvas)[c[:, 0], :, c[:, 2], c[:, 3]]
