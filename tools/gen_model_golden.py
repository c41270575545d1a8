"""Structure + forward fixtures of the UNMODIFIED reference's model builders (src/models/resnet.py,
resnet_spatial.py, amoebanet.py) on CPU -> tests/golden/model_golden.json.  Build container only.

Per configuration: sha256 of the state-dict signature [(key, shape)...], the parameter count, sha256
of the ordered list of conv / pool module kinds (which layers are conv_spatial / Pool vs ordinary),
and for the sequential builders the output of a forward pass with every tensor of the state dict
filled from a generator seeded by its position in key order (tests fill the same way)."""
import hashlib
import json
import os
import sys
import warnings

import torch
import torch.distributed as dist

warnings.simplefilter("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "model_golden.json")


def sig(m):
    return hashlib.sha256(repr([(k, tuple(v.shape)) for k, v in m.state_dict().items()]).encode()).hexdigest()


def kinds(m):
    ks = []
    for n, x in m.named_modules():
        t = type(x).__name__
        if n.endswith(".halo_len_layer"):                  # inner layer of the reference's Pool
            continue
        if t in ("local_conv2d",):                         # this repo's stand-in for the D2 cells' plain convs / pools
            t = "Conv2d"
        if t in ("local_pool2d",):
            t = "AvgPool2d"
        if t in ("Conv2d", "conv_spatial", "Pool", "halo_exchange_layer") or (t in ("AvgPool2d", "MaxPool2d") and not n.endswith(".pool")):
            ks.append((n, t))
    return hashlib.sha256(repr(ks).encode()).hexdigest(), sum(t == "conv_spatial" for _, t in ks), sum(t == "Pool" for _, t in ks)


def fill(m):
    for i, (k, v) in enumerate(sorted(m.state_dict().items())):
        g = torch.Generator().manual_seed(i)
        if v.dtype.is_floating_point:
            v.copy_(torch.rand(v.shape, generator=g) + 0.5 if "running_var" in k else torch.randn(v.shape, generator=g) * 0.1)


def forward(m, size):
    fill(m)
    x = torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(77))
    out = {}
    for mode in ("train", "eval"):
        getattr(m, mode)()
        with torch.no_grad():
            out[mode] = m(x).double().flatten().tolist()
    return out


def entry(m, fwd_size=None):
    kh, nconv, npool = kinds(m)
    e = dict(state_sig=sig(m), params=sum(p.numel() for p in m.parameters()), kinds_sig=kh, spatial_convs=nconv, spatial_pools=npool)
    if fwd_size:
        e["forward"] = forward(m, fwd_size)
    return e


def main():
    sys.path.insert(0, HERE)
    import ref_shim
    ref_shim.install()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29779")
    dist.init_process_group("gloo", rank=0, world_size=1)
    from models import amoebanet, amoebanet_d2, resnet, resnet_spatial
    res = {"resnet": [], "resnet_spatial": [], "amoebanet": [], "amoebanet_spatial": [], "amoebanet_d2_spatial": []}
    for ver, depth in ((1, 20), (2, 29), (2, 101)):
        m = getattr(resnet, "get_resnet_v%d" % ver)((2, 3, 32, 32), depth)
        res["resnet"].append(dict(version=ver, depth=depth, **entry(m, 32)))
        for kw in (dict(local_rank=0, mp_size=2, spatial_size=1, num_spatial_parts=4, balance=None, slice_method="square"),
                   dict(local_rank=0, mp_size=4, spatial_size=2, num_spatial_parts=[2, 2], balance=None, slice_method="vertical"),
                   dict(local_rank=0, mp_size=3, spatial_size=1, num_spatial_parts=2, balance=[3, 4, len(m) - 7], slice_method="horizontal")):
            ms = getattr(resnet_spatial, "get_resnet_v%d" % ver)((2, 3, 64, 64), depth, **kw)
            res["resnet_spatial"].append(dict(version=ver, depth=depth, kw=kw, **entry(ms)))
    for nl, nf in ((3, 64), (6, 128), (18, 416)):
        m = amoebanet.amoebanetd(num_classes=10, num_layers=nl, num_filters=nf)
        res["amoebanet"].append(dict(num_layers=nl, num_filters=nf, **entry(m, 64 if nl < 18 else None)))
        for kw in (dict(mp_size=2, balance=None), dict(mp_size=4, balance=None), dict(mp_size=3, balance=[5, nl + 6 - 7, 2])):
            if kw["balance"] is None and (nl + 6) // kw["mp_size"] <= 3:
                continue
            ms = amoebanet.amoebanetd_spatial(local_rank=0, spatial_size=1, num_spatial_parts=4, slice_method="square",
                                              num_classes=10, num_layers=nl, num_filters=nf, **kw)
            res["amoebanet_spatial"].append(dict(num_layers=nl, num_filters=nf, kw=kw, **entry(ms)))
            md = amoebanet_d2.amoebanetd_spatial(local_rank=0, spatial_size=1, num_spatial_parts=4, slice_method="square",
                                                 num_classes=10, num_layers=nl, num_filters=nf, **kw)
            res["amoebanet_d2_spatial"].append(dict(num_layers=nl, num_filters=nf, kw=kw, **entry(md)))
    json.dump({"source": "tools/gen_model_golden.py on unmodified /root/reference/src/models (CPU)", **res}, open(OUT, "w"), indent=1)
    print({k: len(v) for k, v in res.items()})


if __name__ == "__main__":
    main()
