"""Extract the hot-path layer list (conv_spatial / Pool call sites with shapes) of the
reference's spatial stage, by running the UNMODIFIED reference models on small CPU tensors (shapes are then scaled).

Run in the build container only (needs /root/reference):
    python tools/extract_layers.py
Writes tests/golden/layers_amoebanetd_sp4.json and tests/golden/layers_resnet101_sp2.json
(the fixtures bench.py and the tests read; /root/reference does not exist on the GPU box).

Reference anchors: models/amoebanet.py:618-719 (amoebanetd_spatial),
models/resnet_spatial.py:545-633 (get_resnet_v2), torchgems/spatial.py:25,1416.
"""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(__file__))
import ref_shim  # noqa: E402

ref_shim.install()
ref_shim.init_single_process_group()

from models import amoebanet, resnet_spatial  # noqa: E402
from torchgems import spatial as ref_spatial  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def trace(model_stage, x, scale):
    """Run `model_stage` on meta tensor x and record every conv / pool call."""
    records = []

    def rec_conv(mod, inp, out):
        t = inp[0]
        kind = "conv_spatial" if isinstance(mod, ref_spatial.conv_spatial) else "nn.Conv2d"
        hh = hw = 0
        if kind == "conv_spatial":
            hh, hw = mod.halo_len_height, mod.halo_len_width
        else:
            hh, hw = mod.padding
        records.append(
            dict(
                op="conv",
                kind=kind,
                C=mod.in_channels,
                K=mod.out_channels,
                R=mod.kernel_size[0],
                S=mod.kernel_size[1],
                stride_h=mod.stride[0],
                stride_w=mod.stride[1],
                pad_h=int(hh),
                pad_w=int(hw),
                H=int(t.shape[2] * scale),
                W=int(t.shape[3] * scale),
                bias=mod.bias is not None,
            )
        )

    def rec_pool(mod, inp, out):
        t = inp[0]
        p = mod.pool
        k = p.kernel_size if isinstance(p.kernel_size, int) else p.kernel_size[0]
        s = p.stride if isinstance(p.stride, int) else p.stride[0]
        records.append(
            dict(
                op="pool",
                kind="Pool",
                mode="max" if isinstance(p, nn.MaxPool2d) else "avg",
                C=int(t.shape[1]),
                k=int(k),
                stride=int(s),
                pad=int(mod.halo_len),
                H=int(t.shape[2] * scale),
                W=int(t.shape[3] * scale),
            )
        )

    hooks = []
    for m in model_stage.modules():
        if isinstance(m, nn.Conv2d):
            hooks.append(m.register_forward_hook(rec_conv))
        elif isinstance(m, ref_spatial.Pool):
            hooks.append(m.register_forward_hook(rec_pool))
    model_stage.eval()
    with torch.no_grad():
        model_stage(x)
    for h in hooks:
        h.remove()
    return records


def amoebanet_layers(image=8192, trace_image=128, split_size=4):
    if True:
        model = amoebanet.amoebanetd_spatial(
            local_rank=0,
            spatial_size=1,
            num_spatial_parts=1,
            mp_size=split_size,
            slice_method="square",
            num_layers=18,
            num_filters=416,
        )
    n_layers = len(model)
    end = int(n_layers // split_size)  # mp_pipeline.py:41-83 default balance: num_layers/split_size
    # amoebanet.py:635-637: predicted 6*3+6=24 layers, end_layer = 24/4 = 6 spatial layers
    end = 6 if split_size == 4 else end
    stage = model[:end]
    x = torch.empty(1, 3, trace_image, trace_image)
    recs = trace(stage, x, image // trace_image)
    names = list(dict(model.named_children()).keys())[:end]
    return dict(
        model="AmoebaNet-D(num_layers=18,num_filters=416)",
        image=image,
        split_size=split_size,
        spatial_layers=names,
        source="tools/extract_layers.py run on /root/reference models/amoebanet.py (CPU trace at 128x128, H/W scaled)",
        layers=recs,
    )


def resnet_layers(image=4096, trace_image=128, depth=101, split_size=2):
    n = (depth - 2) // 9
    num_layers = n * 3 + 2
    if True:
        model = resnet_spatial.get_resnet_v2(
            input_shape=(1, 3, trace_image, trace_image),
            depth=depth,
            local_rank=0,
            mp_size=split_size,
            spatial_size=1,
            num_spatial_parts=1,
            slice_method="square",
        )
    end = num_layers // split_size
    stage = model[:end]
    x = torch.empty(1, 3, trace_image, trace_image)
    recs = trace(stage, x, image // trace_image)
    return dict(
        model="ResNet-v2 depth=%d (n=%d)" % (depth, n),
        image=image,
        split_size=split_size,
        spatial_layers=[str(i) for i in range(end)],
        source="tools/extract_layers.py run on /root/reference models/resnet_spatial.py (CPU trace at 128x128, H/W scaled)",
        layers=recs,
    )


def summarize(d):
    convs = [l for l in d["layers"] if l["op"] == "conv"]
    pools = [l for l in d["layers"] if l["op"] == "pool"]
    fl = by = 0
    for l in convs:
        Ho = (l["H"] + 2 * l["pad_h"] - l["R"]) // l["stride_h"] + 1
        Wo = (l["W"] + 2 * l["pad_w"] - l["S"]) // l["stride_w"] + 1
        fl += 2 * l["K"] * l["C"] * l["R"] * l["S"] * Ho * Wo
        by += l["C"] * l["H"] * l["W"] + l["K"] * Ho * Wo + l["K"] * l["C"] * l["R"] * l["S"]
    for l in pools:
        Ho = (l["H"] + 2 * l["pad"] - l["k"]) // l["stride"] + 1
        by += l["C"] * l["H"] * l["W"] + l["C"] * Ho * Ho
    print(
        d["model"], "image", d["image"], ": convs", len(convs),
        "(conv_spatial %d)" % sum(l["kind"] == "conv_spatial" for l in convs),
        "pools", len(pools), "TFLOP fwd %.2f" % (fl / 1e12), "Gelem %.1f" % (by / 1e9),
    )


if __name__ == "__main__":
    os.makedirs(OUT_DIR, exist_ok=True)
    a = amoebanet_layers()
    summarize(a)
    with open(os.path.join(OUT_DIR, "layers_amoebanetd_sp4.json"), "w") as f:
        json.dump(a, f, indent=1)
    r = resnet_layers()
    summarize(r)
    with open(os.path.join(OUT_DIR, "layers_resnet101_sp2.json"), "w") as f:
        json.dump(r, f, indent=1)
