"""Known-answer fixtures of the reference's self-checking halo benchmarks
(benchmarks/communication/halo/benchmark_sp_halo_exchange*.py): the full image is
arange(B*C*H*W) as float32 NCHW, weights and bias are ones, so every expected value is an exact
integer for small images.  Pure numpy (no torch, no GPU) so the expectations can be unit-tested."""
import math

import numpy as np


def grid(slice_method, parts):
    if slice_method == "square":
        q = int(math.sqrt(parts))
        return q, q
    return (1, parts) if slice_method == "vertical" else (parts, 1)


def full_image(batch, channels, size):
    return np.arange(batch * channels * size * size, dtype=np.float32).reshape(batch, channels, size, size)


def tile_box(slice_method, parts, rank, size):
    rows, cols = grid(slice_method, parts)
    th, tw = size // rows, size // cols
    r, c = rank // cols, rank % cols
    return r * th, (r + 1) * th, c * tw, (c + 1) * tw


def tile(full, slice_method, parts, rank):
    y0, y1, x0, x1 = tile_box(slice_method, parts, rank, full.shape[-1])
    return np.ascontiguousarray(full[:, :, y0:y1, x0:x1])


def expected_padded_tile(full, slice_method, parts, rank, halo_h, halo_w=None):
    """The tile plus its halo cut from the zero-padded full image: what pad -> exchange -> unpack must give."""
    halo_w = halo_h if halo_w is None else halo_w
    y0, y1, x0, x1 = tile_box(slice_method, parts, rank, full.shape[-1])
    padded = np.pad(full, ((0, 0), (0, 0), (halo_h, halo_h), (halo_w, halo_w)))
    return np.ascontiguousarray(padded[:, :, y0:y1 + 2 * halo_h, x0:x1 + 2 * halo_w])


def expected_conv_tile(full, slice_method, parts, rank, kh, kw, out_channels):
    """Tile of conv2d(full, ones[K][C][kh][kw], bias = 1, "same" padding): box sums over the window and all
    input channels, + 1, replicated over the K output channels.  float64 -> exact for integers < 2^53."""
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    s = np.pad(full.astype(np.float64).sum(axis=1), ((0, 0), (ph, ph), (pw, pw)))     # [B][H+2ph][W+2pw]
    c = np.cumsum(np.cumsum(np.pad(s, ((0, 0), (1, 0), (1, 0))), axis=1), axis=2)     # summed-area table
    H, W = full.shape[-2:]
    box = c[:, kh:kh + H, kw:kw + W] - c[:, :H, kw:kw + W] - c[:, kh:kh + H, :W] + c[:, :H, :W]
    y0, y1, x0, x1 = tile_box(slice_method, parts, rank, W)
    out = box[:, y0:y1, x0:x1] + 1.0
    return np.repeat(out[:, None], out_channels, axis=1)
