"""Generate golden input/output vectors from the UNMODIFIED reference (torchgems.spatial)
running on CPU over gloo (see tools/ref_shim.py for the six import shims).

Run in the build container only (needs /root/reference):
    python tools/gen_golden.py
Writes tests/golden/spatial_golden.npz  (committed; /root/reference does not travel).

For every case, P processes each own one tile of a full image, build the reference module
(spatial.py:25 conv_spatial / :1032 halo_exchange_layer / :1416 Pool) with identical seeded
weights, run forward and autograd backward against a seeded upstream gradient, and rank 0
gathers per-tile results.  Stored per case:
   x (full image), w, b, gy (full upstream grad), and per tile t: y_t, dx_t, dw_t, db_t.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "spatial_golden.npz")

# (name, P, slice_method)
GRIDS = [("sq4", 4, "square"), ("v2", 2, "vertical"), ("v4", 4, "vertical"),
         ("h2", 2, "horizontal"), ("h4", 4, "horizontal")]

# conv cases: (tag, C, K, (R,S), (sh,sw), bias)
CONVS = [
    ("c3x3s1", 3, 4, (3, 3), (1, 1), True),
    ("c3x3s2", 2, 3, (3, 3), (2, 2), False),
    ("c1x7", 3, 2, (1, 7), (1, 1), False),
    ("c7x1", 2, 3, (7, 1), (1, 1), False),
    ("c1x1", 4, 5, (1, 1), (1, 1), True),
    ("c5x5s1", 2, 2, (5, 5), (1, 1), True),
]
# pool cases: (tag, mode, k, stride, pad)
POOLS = [
    ("avg3s1", "AvgPool2d", 3, 1, 1),
    ("avg3s2", "AvgPool2d", 3, 2, 1),
    ("max3s1", "MaxPool2d", 3, 1, 1),
    ("max2s2", "MaxPool2d", 2, 2, 0),
]
HALOS = [1, 2, 3]
IMAGE = 32  # full image edge; tiles are 16x16 (sq4), 32x16/32x8 (v), 16x32/8x32 (h)
BATCH = 2


def tile_slices(method, P, rank, H, W):
    """train_spatial.py:241-290 split_input."""
    if method == "square":
        q = int(round(P ** 0.5))
        r, c = rank // q, rank % q
        th, tw = H // q, W // q
        return slice(r * th, (r + 1) * th), slice(c * tw, (c + 1) * tw)
    if method == "vertical":
        tw = W // P
        return slice(0, H), slice(rank * tw, (rank + 1) * tw)
    th = H // P
    return slice(rank * th, (rank + 1) * th), slice(0, W)


def make_data(kind, shape, seed):
    if kind == "kat":  # benchmark_sp_halo_exchange_conv.py:421-427 arange input
        return torch.arange(int(np.prod(shape)), dtype=torch.float32).reshape(shape)
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=torch.float32)


def worker(rank, P, method, port, cases, ret):
    sys.path.insert(0, HERE)
    import ref_shim

    ref_shim.install()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=P)
    torch.set_num_threads(1)
    from torchgems import spatial as ref

    out = {}
    for case in cases:
        kind = case["kind"]
        name = case["name"]
        x_full = make_data(case["data"], (BATCH, case["C"], IMAGE, IMAGE), case["seed"])
        hs, ws = tile_slices(method, P, rank, IMAGE, IMAGE)
        x = x_full[:, :, hs, ws].clone().requires_grad_(True)
        if kind == "conv":
            R, S = case["k"]
            m = ref.conv_spatial(
                local_rank=rank, spatial_size=1, num_spatial_parts=P,
                in_channels=case["C"], out_channels=case["K"], kernel_size=(R, S),
                stride=tuple(case["stride"]), padding=((R - 1) // 2, (S - 1) // 2),
                bias=case["bias"], slice_method=method,
            )
            with torch.no_grad():
                if case["data"] == "kat":  # benchmark_sp_halo_exchange_conv.py:913-915
                    m.weight.fill_(1.0)
                    if m.bias is not None:
                        m.bias.fill_(1.0)
                else:
                    g = torch.Generator().manual_seed(case["seed"] + 1)
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.2)
                    if m.bias is not None:
                        m.bias.copy_(torch.randn(m.bias.shape, generator=g))
        elif kind == "pool":
            m = ref.Pool(
                local_rank=rank, spatial_size=1, num_spatial_parts=P,
                kernel_size=case["k"], stride=case["stride"], padding=case["pad"],
                slice_method=method, operation=case["mode"], count_include_pad=False,
            )
        else:
            m = ref.halo_exchange_layer(
                local_rank=rank, spatial_size=1, num_spatial_parts=P,
                halo_len=case["halo"], slice_method=method,
            )
        y = m(x)
        # upstream gradient: a seeded full-size tensor sliced like the output tile
        if kind == "halo":
            gy = make_data("randn", tuple(y.shape), case["seed"] + 7 + rank)
        else:
            oh = IMAGE // (case["stride"][0] if kind == "conv" else case["stride"])
            ow = IMAGE // (case["stride"][1] if kind == "conv" else case["stride"])
            gy_full = make_data("randn", (BATCH, y.shape[1], oh, ow), case["seed"] + 7)
            ohs, ows = tile_slices(method, P, rank, oh, ow)
            gy = gy_full[:, :, ohs, ows]
            assert gy.shape == y.shape, (name, gy.shape, y.shape)
        y.backward(gy)
        rec = {"y": y.detach().numpy(), "dx": x.grad.numpy(), "gy": gy.numpy()}
        if kind == "conv":
            rec["dw"] = m.weight.grad.numpy()
            rec["w"] = m.weight.detach().numpy()
            if m.bias is not None:
                rec["db"] = m.bias.grad.numpy()
                rec["b"] = m.bias.detach().numpy()
        if rank == 0:
            rec["x"] = x_full.numpy()
        out[name] = rec
        dist.barrier()
    gathered = [None] * P if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    if rank == 0:
        ret.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def build_cases():
    cases = []
    seed = 100
    for tag, C, K, k, st, bias in CONVS:
        for data in ("kat", "randn"):
            seed += 1
            cases.append(dict(kind="conv", name=f"conv_{tag}_{data}", C=C, K=K, k=list(k),
                              stride=list(st), bias=bias, data=data, seed=seed))
    for tag, mode, k, st, pad in POOLS:
        seed += 1
        cases.append(dict(kind="pool", name=f"pool_{tag}", C=3, k=k, stride=st, pad=pad,
                          mode=mode, data="randn", seed=seed))
    for h in HALOS:
        seed += 1
        cases.append(dict(kind="halo", name=f"halo_{h}", C=2, halo=h, data="kat", seed=seed))
    return cases


def main():
    cases = build_cases()
    arrays = {}
    meta = {"image": IMAGE, "batch": BATCH, "grids": [], "cases": cases,
            "source": "tools/gen_golden.py: unmodified /root/reference src/torchgems/spatial.py on CPU/gloo"}
    port = 29610
    ctx = mp.get_context("spawn")
    for gname, P, method in GRIDS:
        port += 1
        q = ctx.SimpleQueue()
        procs = [ctx.Process(target=worker, args=(r, P, method, port, cases, q)) for r in range(P)]
        for p in procs:
            p.start()
        gathered = q.get()
        for p in procs:
            p.join()
            assert p.exitcode == 0
        meta["grids"].append({"name": gname, "P": P, "method": method})
        for r, out in enumerate(gathered):
            for cname, rec in out.items():
                for k, v in rec.items():
                    if k in ("x", "w", "b"):
                        if r == 0:
                            arrays.setdefault(f"{cname}/{k}", v)  # same across grids
                        continue
                    arrays[f"{gname}/{cname}/{k}/{r}"] = v
        print("done", gname, flush=True)
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB;", len(arrays), "arrays")


if __name__ == "__main__":
    main()
