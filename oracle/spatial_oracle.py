"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

A CPU (numpy) restatement of the reference's spatially-partitioned Conv2d / Pool2d /
halo-exchange algorithm (OSU-Nowlab/MPI4DL src/torchgems/spatial.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this; the product path (mpi4dl_b200/) never does and fails loudly without its CUDA library.

Parity pinning: checked against tests/golden/spatial_golden.npz, which was produced by
running the UNMODIFIED reference on CPU/gloo (tools/gen_golden.py) -- see
tests/test_oracle_golden.py.  The arithmetic itself (`torch.nn.Conv2d`, `nn.AvgPool2d`,
`nn.MaxPool2d`) lives in PyTorch (third-party, not vendored in the reference; the reference
pins it by prose only: "PyTorch 1.12.1 or 1.13.1", README.md:109); it is restated here from
its published definition (cross-correlation, `padding=0` on an explicitly padded tile).

All ranks of a spatial group are simulated in ONE process: `tiles` is the list of per-rank
NCHW arrays.  Every function cites the reference lines it follows.
"""
import math

import numpy as np

# 3x3 neighbour stencil of the reference (spatial.py:961-964):   0 1 2 / 3 4 5 / 6 7 8
DIRS = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]


def grid_shape(method, P):
    """Rank grid (rows, cols).  spatial.py:871-904, train_spatial.py:241-290."""
    if method == "square":
        q = int(math.sqrt(P))
        assert q * q == P, "square slicing needs a perfect-square part count"
        return q, q
    if method == "vertical":
        return 1, P
    if method == "horizontal":
        return P, 1
    raise ValueError(method)


def neighbour_mask(method, P, rank, kh=3, kw=3):
    """spatial.py:941-1017 get_neighbours + :921-939 set_neighbours_based_on_kernel_size."""
    rows, cols = grid_shape(method, P)
    r, c = rank // cols, rank % cols
    mask = []
    for dr, dc in DIRS:
        rr, cc = r + dr, c + dc
        ok = (dr, dc) != (0, 0) and 0 <= rr < rows and 0 <= cc < cols
        mask.append(1 if ok else 0)
    if kh == 1:  # no vertical reach: drop up/down and the corners
        for i in (0, 1, 2, 6, 7, 8):
            mask[i] = 0
    if kw == 1:
        for i in (0, 3, 6, 2, 5, 8):
            mask[i] = 0
    return mask


def neighbour_ranks(method, P, rank, mask):
    """spatial.py:868-910 get_neighbours_rank (non-GEMS-inverse case)."""
    rows, cols = grid_shape(method, P)
    out = []
    for i, (dr, dc) in enumerate(DIRS):
        out.append(rank + dr * cols + dc if mask[i] else -1)
    return out


def tile_slices(method, P, rank, H, W):
    """train_spatial.py:241-290 split_input (even partitions)."""
    rows, cols = grid_shape(method, P)
    r, c = rank // cols, rank % cols
    th, tw = H // rows, W // cols
    return slice(r * th, (r + 1) * th), slice(c * tw, (c + 1) * tw)


def split(full, method, P):
    H, W = full.shape[2:]
    return [np.ascontiguousarray(full[:, :, hs, ws])
            for hs, ws in (tile_slices(method, P, r, H, W) for r in range(P))]


def _send_region(i, hh, hw, Hp, Wp):
    """spatial.py:239-309 locations_send, as explicit index ranges in the PADDED tile:
    the first / last `halo` REAL rows and columns just inside the pad."""
    dr, dc = DIRS[i]
    rows = {-1: (hh, 2 * hh), 0: (hh, Hp - hh), 1: (Hp - 2 * hh, Hp - hh)}[dr]
    cols = {-1: (hw, 2 * hw), 0: (hw, Wp - hw), 1: (Wp - 2 * hw, Wp - hw)}[dc]
    return rows, cols


def _recv_region(i, hh, hw, Hp, Wp):
    """spatial.py:177-237 locations_recv: the pad strips themselves."""
    dr, dc = DIRS[i]
    rows = {-1: (0, hh), 0: (hh, Hp - hh), 1: (Hp - hh, Hp)}[dr]
    cols = {-1: (0, hw), 0: (hw, Wp - hw), 1: (Wp - hw, Wp)}[dc]
    return rows, cols


def exchange_halos(tiles, method, hh, hw, kh=None, kw=None):
    """spatial.py:1019-1025 (pad, start/end_halo_exchange, copy_halo_exchange_values).

    Returns the padded tiles after the exchange.  kh/kw prune directions for conv_spatial
    (1-D kernels); halo_exchange_layer passes None (no pruning, spatial.py:1329-1392).
    """
    P = len(tiles)
    padded = [np.pad(t, ((0, 0), (0, 0), (hh, hh), (hw, hw))) for t in tiles]  # ZeroPad2d :1020
    if hh == 0 and hw == 0:
        return padded
    src = [p.copy() for p in padded]  # sends are clones taken before any unpack (:340-349)
    for rank in range(P):
        mask = neighbour_mask(method, P, rank, kh if kh is not None else 3, kw if kw is not None else 3)
        nbr = neighbour_ranks(method, P, rank, mask)
        Hp, Wp = padded[rank].shape[2:]
        for i in range(9):
            if not mask[i]:
                continue
            peer = src[nbr[i]]
            # peer sends its direction (8-i) strip; tags pair send[i] with recv[8-i] (:170-172)
            (sr0, sr1), (sc0, sc1) = _send_region(8 - i, hh, hw, peer.shape[2], peer.shape[3])
            (rr0, rr1), (rc0, rc1) = _recv_region(i, hh, hw, Hp, Wp)
            padded[rank][:, :, rr0:rr1, rc0:rc1] = peer[:, :, sr0:sr1, sc0:sc1]
    return padded


# ---------------------------------------------------------------------------------------
# Arithmetic: padding=0 cross-correlation and pooling on an explicitly padded tile
# (the reference's call sites: spatial.py:1027 nn.Conv2d.forward, :1480-1498 pools).
# ---------------------------------------------------------------------------------------

def conv2d_fwd(xp, w, b=None, stride=(1, 1)):
    """y[n,k,i,j] = b[k] + sum_{c,r,s} w[k,c,r,s] * xp[n,c,i*sh+r,j*sw+s]   (spatial.py:1027)."""
    sh, sw = stride
    N, C, Hp, Wp = xp.shape
    K, C2, R, S = w.shape
    assert C == C2
    Ho, Wo = (Hp - R) // sh + 1, (Wp - S) // sw + 1
    y = np.zeros((N, K, Ho, Wo), dtype=np.float64)
    for r in range(R):
        for s in range(S):
            patch = xp[:, :, r:r + (Ho - 1) * sh + 1:sh, s:s + (Wo - 1) * sw + 1:sw]
            y += np.einsum("kc,nchw->nkhw", w[:, :, r, s].astype(np.float64), patch.astype(np.float64))
    if b is not None:
        y += b.reshape(1, K, 1, 1)
    return y.astype(np.float32)


def conv2d_bwd(xp, w, gy, stride=(1, 1), need_db=True):
    """Autograd of conv2d_fwd w.r.t. the padded tile, weight and bias."""
    sh, sw = stride
    N, C, Hp, Wp = xp.shape
    K, _, R, S = w.shape
    Ho, Wo = gy.shape[2:]
    dxp = np.zeros(xp.shape, dtype=np.float64)
    dw = np.zeros(w.shape, dtype=np.float64)
    g64 = gy.astype(np.float64)
    for r in range(R):
        for s in range(S):
            sl = (slice(None), slice(None), slice(r, r + (Ho - 1) * sh + 1, sh), slice(s, s + (Wo - 1) * sw + 1, sw))
            dxp[sl] += np.einsum("kc,nkhw->nchw", w[:, :, r, s].astype(np.float64), g64)
            dw[:, :, r, s] = np.einsum("nkhw,nchw->kc", g64, xp[sl].astype(np.float64))
    db = g64.sum(axis=(0, 2, 3)).astype(np.float32) if need_db else None
    return dxp.astype(np.float32), dw.astype(np.float32), db


def pool_fwd(xp, mode, k, stride):
    """nn.MaxPool2d / nn.AvgPool2d with padding=0 on the padded tile (spatial.py:1478-1498).
    The pad is explicit zeros, so avg always divides by k*k and max sees 0 at true borders."""
    N, C, Hp, Wp = xp.shape
    Ho, Wo = (Hp - k) // stride + 1, (Wp - k) // stride + 1
    win = np.stack([xp[:, :, r:r + (Ho - 1) * stride + 1:stride, s:s + (Wo - 1) * stride + 1:stride]
                    for r in range(k) for s in range(k)], axis=0)
    if mode == "max":
        return win.max(axis=0)
    return (win.astype(np.float64).sum(axis=0) / (k * k)).astype(np.float32)


def pool_bwd(xp, gy, mode, k, stride):
    """Autograd of pool_fwd w.r.t. the padded tile.  Max routes to the FIRST maximal element
    in row-major window order (ATen max_pool2d backward)."""
    N, C, Hp, Wp = xp.shape
    Ho, Wo = gy.shape[2:]
    dxp = np.zeros(xp.shape, dtype=np.float64)
    if mode == "avg":
        for r in range(k):
            for s in range(k):
                dxp[:, :, r:r + (Ho - 1) * stride + 1:stride, s:s + (Wo - 1) * stride + 1:stride] += gy / (k * k)
        return dxp.astype(np.float32)
    win = np.stack([xp[:, :, r:r + (Ho - 1) * stride + 1:stride, s:s + (Wo - 1) * stride + 1:stride]
                    for r in range(k) for s in range(k)], axis=0)
    arg = win.argmax(axis=0)  # first max in (r,s) row-major order
    for idx in range(k * k):
        r, s = idx // k, idx % k
        dxp[:, :, r:r + (Ho - 1) * stride + 1:stride, s:s + (Wo - 1) * stride + 1:stride] += np.where(arg == idx, gy, 0.0)
    return dxp.astype(np.float32)


def crop(dxp, hh, hw):
    """ZeroPad2d backward = crop to the tile; received halos are detached constants, so no
    gradient flows back to the neighbour (SURVEY 8a note N2: no backward halo exchange)."""
    Hp, Wp = dxp.shape[2:]
    return np.ascontiguousarray(dxp[:, :, hh:Hp - hh, hw:Wp - hw])


# ---------------------------------------------------------------------------------------
# The three reference modules, all ranks at once
# ---------------------------------------------------------------------------------------

def conv_spatial(tiles, w, b, method, stride=(1, 1), gys=None):
    """spatial.py:25-1029 conv_spatial forward (+ autograd backward if gys given).
    Returns list of dicts per rank: y, and dx/dw/db when gys is not None."""
    R, S = w.shape[2:]
    hh, hw = (R - 1) // 2, (S - 1) // 2  # :115-117
    padded = exchange_halos(tiles, method, hh, hw, kh=R, kw=S)
    out = []
    for rank, xp in enumerate(padded):
        rec = {"y": conv2d_fwd(xp, w, b, stride)}
        if gys is not None:
            dxp, dw, db = conv2d_bwd(xp, w, gys[rank], stride, need_db=b is not None)
            rec.update(dx=crop(dxp, hh, hw), dw=dw, db=db)
        out.append(rec)
    return out


def halo_exchange_layer(tiles, method, halo_len, gys=None):
    """spatial.py:1032-1413: pad by halo_len and fill from all 8 neighbours."""
    padded = exchange_halos(tiles, method, halo_len, halo_len)
    out = []
    for rank, xp in enumerate(padded):
        rec = {"y": xp}
        if gys is not None:
            rec["dx"] = crop(gys[rank], halo_len, halo_len)
        out.append(rec)
    return out


def pool_spatial(tiles, method, mode, k, stride, pad, gys=None):
    """spatial.py:1416-1509 Pool: halo_exchange_layer(halo=floor((k-1)/2)) then pool(padding=0)."""
    halo = (k - 1) // 2  # :1457
    assert halo == pad  # :1462-1464
    padded = exchange_halos(tiles, method, halo, halo) if halo else [t for t in tiles]
    out = []
    for rank, xp in enumerate(padded):
        rec = {"y": pool_fwd(xp, mode, k, stride)}
        if gys is not None:
            rec["dx"] = crop(pool_bwd(xp, gys[rank], mode, k, stride), halo, halo)
        out.append(rec)
    return out
