"""CPU-only, world_size 2 and 4 over gloo: the N>1 host path -- neighbour ranks, strip shapes and
the direction pairing (strip i is received as the neighbour's strip 8-i) -- checked against the
oracle's halo exchange.  The strips are cut with plain slicing HERE, in the test; the product's
pack kernel is CUDA-only (tests/test_gpu_*.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import spatial_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, P, method, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=P)
    from mpi4dl_b200.torchgems import halo_transport as ht
    from mpi4dl_b200.torchgems import spatial

    torch.manual_seed(0)
    full = torch.randn(2, 3, 16, 16)
    ok = True
    for hh, hw, k in [(1, 1, (3, 3)), (2, 2, (5, 5)), (0, 3, (1, 7)), (3, 0, (7, 1))]:
        layer = spatial.conv_spatial(rank, 1, P, 3, 3, k, padding=(hh, hw), slice_method=method)
        hs, ws = so.tile_slices(method, P, rank, 16, 16)
        x = full[:, :, hs, ws].contiguous()
        N, C, H, W = x.shape
        send, recv = [None] * 9, [None] * 9
        for i in range(9):
            if i != 4 and layer.neighbours[i]:
                dr, dc = so.DIRS[i]
                rs = {-1: slice(0, hh), 0: slice(0, H), 1: slice(H - hh, H)}[dr]
                cs = {-1: slice(0, hw), 0: slice(0, W), 1: slice(W - hw, W)}[dc]
                send[i] = x[:, :, rs, cs].contiguous()
                recv[i] = torch.empty(ht.strip_shape(i, N, C, H, W, hh, hw))
        ht.exchange_strips(send, recv, layer.rank_neighbours)
        tiles = so.split(full.numpy(), method, P)
        xp = so.exchange_halos(tiles, method, hh, hw, kh=k[0], kw=k[1])[rank]
        for i in range(9):
            if recv[i] is None:
                continue
            (r0, r1), (c0, c1) = so._recv_region(i, hh, hw, xp.shape[2], xp.shape[3])
            ok = ok and np.array_equal(recv[i].numpy(), xp[:, :, r0:r1, c0:c1])
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P,method,port", [(2, "vertical", 29711), (2, "horizontal", 29712), (4, "square", 29713)])
def test_exchange_strips_over_gloo(P, method, port):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, P, method, port, q)) for r in range(P)]
    for p in procs:
        p.start()
    res = [q.get() for _ in range(P)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
