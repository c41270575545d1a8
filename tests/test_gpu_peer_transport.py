"""-m gpu: the real multi-process halo path.  Two (or four) processes, one tile each, exchange halo
strips through the CUDA-IPC mailbox transport (pack kernel writes straight into the neighbour's
arena; device-side flags order the streams) and through the torch.distributed transport, and run
conv_spatial / Pool / halo_exchange_layer forward + backward.  Results are checked per rank
against the oracle.  With >= P GPUs every rank gets its own device (NVLink peer stores, NCCL for
the handle hand-shake); on a single GPU the ranks share cuda:0 (IPC mapping of the same device,
gloo hand-shake) -- the protocol and kernels are identical."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import spatial_oracle as so

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, P, method, transport, port, ngpu, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SPCONV_HALO_TRANSPORT"] = transport
    os.environ["SPCONV_ARENA_MB"] = "64"
    multi = ngpu >= P
    dev = torch.device("cuda", rank if multi else 0)
    torch.cuda.set_device(dev)
    if multi:
        dist.init_process_group("nccl", rank=rank, world_size=P, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=P)
    from mpi4dl_b200.torchgems import spatial

    errs = []
    try:
        rng = np.random.default_rng(7)
        full = rng.standard_normal((1, 64, 32, 256)).astype(np.float32)
        full = torch.tensor(full).to(torch.bfloat16).float().numpy()
        tiles = so.split(full, method, P)
        # --- conv_spatial, several kernels, repeated calls (exercises sequence numbers / parity / acks)
        for (K, R, S, st) in [(64, 3, 3, 1), (32, 1, 7, 1), (32, 7, 1, 1), (64, 3, 3, 2), (128, 1, 1, 1)]:
            w = (rng.standard_normal((K, 64, R, S)) / np.sqrt(64 * R * S)).astype(np.float32)
            w = torch.tensor(w).to(torch.bfloat16).float().numpy()
            m = spatial.conv_spatial(rank, 1, P, 64, K, (R, S), stride=st, padding=((R - 1) // 2, (S - 1) // 2),
                                     bias=False, slice_method=method).to(dev).to(torch.bfloat16)
            with torch.no_grad():
                m.weight.copy_(torch.tensor(w))
            ref = so.conv_spatial(tiles, w, None, method, (st, st), None)
            gy = rng.standard_normal(ref[rank]["y"].shape).astype(np.float32)
            gys = [np.zeros_like(r["y"]) for r in ref]
            gys[rank] = torch.tensor(gy).to(torch.bfloat16).float().numpy()
            refb = so.conv_spatial(tiles, w, None, method, (st, st), gys)
            for it in range(4):
                x = torch.tensor(tiles[rank], dtype=torch.bfloat16, device=dev, requires_grad=True)
                y = m(x)
                y.backward(torch.tensor(gys[rank], dtype=torch.bfloat16, device=dev))
                yr = ref[rank]["y"]
                e = np.abs(y.detach().float().cpu().numpy() - yr).max() / max(1e-6, np.abs(yr).max())
                edx = np.abs(x.grad.float().cpu().numpy() - refb[rank]["dx"]).max() / max(1e-6, np.abs(refb[rank]["dx"]).max())
                edw = np.abs(m.weight.grad.float().cpu().numpy() - refb[rank]["dw"]).max() / max(1e-6, np.abs(refb[rank]["dw"]).max())
                m.weight.grad = None
                if e > 2e-2 or edx > 2e-2 or edw > 2e-2:
                    errs.append(("conv", K, R, S, st, it, float(e), float(edx), float(edw)))
        # --- Pool + halo_exchange_layer (fp32)
        for (mode, k, st) in [("AvgPool2d", 3, 1), ("AvgPool2d", 3, 2), ("MaxPool2d", 3, 1)]:
            pm = spatial.Pool(rank, 1, P, k, st, 1, slice_method=method, operation=mode)
            ref = so.pool_spatial(tiles, method, "avg" if mode == "AvgPool2d" else "max", k, st, 1)
            for it in range(3):
                y = pm(torch.tensor(tiles[rank], device=dev))
                e = np.abs(y.cpu().numpy() - ref[rank]["y"]).max()
                if e > 1e-5:
                    errs.append(("pool", mode, k, st, it, float(e)))
        for h in (1, 2, 3):
            hl = spatial.halo_exchange_layer(rank, 1, P, h, slice_method=method)
            ref = so.halo_exchange_layer(tiles, method, h)
            for it in range(3):
                y = hl(torch.tensor(tiles[rank], device=dev))
                if not np.array_equal(y.cpu().numpy(), ref[rank]["y"]):
                    errs.append(("halo", h, it))
        torch.cuda.synchronize()
    except Exception as ex:  # report instead of hanging the peers
        import traceback
        errs.append(("exception", repr(ex), traceback.format_exc()[-800:]))
    q.put((rank, errs))
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass


@pytest.mark.parametrize("P,method,transport,port", [
    (2, "vertical", "peer", 29811), (2, "horizontal", "peer", 29812), (4, "square", "peer", 29813),
    (2, "vertical", "dist", 29814)])
def test_multiprocess_halo_exchange(P, method, transport, port):
    ngpu = torch.cuda.device_count()
    if transport == "dist" and ngpu < P:
        pytest.skip("torch.distributed transport on GPUs needs NCCL with one GPU per rank")
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, P, method, transport, port, ngpu, q)) for r in range(P)]
    for p in procs:
        p.start()
    res = []
    for _ in range(P):
        res.append(q.get())
    for p in procs:
        p.join(120)
        if p.is_alive():
            p.kill()
    bad = [(r, e) for r, e in res if e]
    assert not bad, bad
