"""Helpers for the -m gpu parity tests: run the product ops for ONE tile on cuda:0 with halo
strips cut from the oracle's exchange, so every neighbour configuration is exercised on a
single GPU (the real transports are tested in test_gpu_peer_transport.py)."""
import numpy as np
import torch

from oracle import spatial_oracle as so


def strips_from_padded(xp, mask, hh, hw, dtype, device="cuda:0"):
    out = [None] * 9
    Hp, Wp = xp.shape[2:]
    for i in range(9):
        if i != 4 and mask[i]:
            (r0, r1), (c0, c1) = so._recv_region(i, hh, hw, Hp, Wp)
            out[i] = torch.tensor(np.ascontiguousarray(xp[:, :, r0:r1, c0:c1]), dtype=dtype, device=device)
    return out


def t(a, dtype=torch.float32, device="cuda:0", grad=False):
    x = torch.tensor(np.ascontiguousarray(a), dtype=dtype, device=device)
    return x.requires_grad_(True) if grad else x


def conv_tile(x_np, w_np, b_np, gy_np, strips, stride, dtype=torch.float32, algo=0):
    """fwd + bwd of one tile through _ConvSpatialFn (C ABI).  Returns y, dx, dw, db as fp32 numpy."""
    from mpi4dl_b200 import _lib
    from mpi4dl_b200.torchgems.spatial import _ConvSpatialFn

    x = t(x_np, dtype, grad=True)
    w = t(w_np, dtype, grad=True)
    b = t(b_np, dtype, grad=True) if b_np is not None else None
    N, C, H, W = x.shape
    K, _, R, S = w.shape
    desc = (N, C, H, W, K, R, S, stride[0], stride[1], (R - 1) // 2, (S - 1) // 2, _lib.dtype_code(dtype), algo)
    y = _ConvSpatialFn.apply(x, w, b, desc, *strips)
    out = {"y": y.detach().float().cpu().numpy()}
    if gy_np is not None:
        y.backward(t(gy_np, dtype))
        out["dx"] = x.grad.float().cpu().numpy()
        out["dw"] = w.grad.float().cpu().numpy()
        if b is not None:
            out["db"] = b.grad.float().cpu().numpy()
    return out


def pool_tile(x_np, gy_np, strips, mode, k, stride, dtype=torch.float32):
    from mpi4dl_b200 import _lib
    from mpi4dl_b200.torchgems.spatial import _PoolFn

    x = t(x_np, dtype, grad=True)
    N, C, H, W = x.shape
    desc = (N, C, H, W, k, stride, (k - 1) // 2, _lib.SPC_POOL_MAX if mode == "max" else _lib.SPC_POOL_AVG,
            _lib.dtype_code(dtype))
    y = _PoolFn.apply(x, desc, *strips)
    out = {"y": y.detach().float().cpu().numpy()}
    if gy_np is not None:
        y.backward(t(gy_np, dtype))
        out["dx"] = x.grad.float().cpu().numpy()
    return out


def bf16_round(a):
    return torch.tensor(a).to(torch.bfloat16).float().numpy()
