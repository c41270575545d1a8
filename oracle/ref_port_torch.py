"""ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference's CPU execution of the hot path, restated
with the same PyTorch CPU operators the reference itself calls when run on CPU tensors:

    conv_spatial.forward   = nn.ZeroPad2d -> (halo exchange) -> nn.Conv2d.forward(padding=0)   spatial.py:1019-1029
    Pool.forward           = nn.ZeroPad2d -> (halo exchange) -> nn.{Avg,Max}Pool2d(padding=0)   spatial.py:1503-1509
    backward               = torch.autograd of the above

With one spatial part there are no neighbours, so the halo exchange is the zero pad only.  Used
ONLY by bench.py (cpu_baseline and `--impl reference`) to time the reference's own CPU path on
the GPU box's host cores: /root/reference does not travel to the GPU box, so the reference's
three-line forward is restated here against the PyTorch ops it calls.  Never imported by the
product.
"""
import time

import torch
import torch.nn.functional as F


def run_layer(layer, scale, dtype=torch.float32, first=False):
    """One fwd+bwd of a conv / pool layer description (tests/golden/layers_*.json) with H, W
    divided by `scale`.  Returns (seconds, flops)."""
    H, W = max(layer["H"] // scale, 8), max(layer["W"] // scale, 8)
    if layer["op"] == "conv":
        C, K, R, S = layer["C"], layer["K"], layer["R"], layer["S"]
        x = torch.randn(1, C, H, W, dtype=dtype, requires_grad=not first)
        w = torch.randn(K, C, R, S, dtype=dtype, requires_grad=True)
        t0 = time.perf_counter()
        xp = F.pad(x, (layer["pad_w"], layer["pad_w"], layer["pad_h"], layer["pad_h"]))   # ZeroPad2d, spatial.py:1020
        y = F.conv2d(xp, w, None, (layer["stride_h"], layer["stride_w"]), 0)              # spatial.py:1027
        y.backward(torch.ones_like(y))
        dt = time.perf_counter() - t0
        flops = 2.0 * K * C * R * S * y.shape[2] * y.shape[3] * (2 if first else 3)
        return dt, flops
    x = torch.randn(1, layer["C"], H, W, dtype=dtype, requires_grad=True)
    t0 = time.perf_counter()
    p = layer["pad"]
    xp = F.pad(x, (p, p, p, p))
    if layer["mode"] == "max":
        y = F.max_pool2d(xp, layer["k"], layer["stride"], 0)
    else:
        y = F.avg_pool2d(xp, layer["k"], layer["stride"], 0)
    y.backward(torch.ones_like(y))
    return time.perf_counter() - t0, 0.0


def run_workload(layers, scale, threads=None, warm=True):
    """One pass (fwd+bwd of every layer) at 1/scale linear size.  Each distinct layer is run once
    untimed first (oneDNN primitive creation / JIT per new shape is not part of the steady-state
    step), then timed.  Returns seconds of the timed pass."""
    if threads:
        torch.set_num_threads(threads)
    total = 0.0
    seen = set()
    for i, l in enumerate(layers):
        key = tuple(sorted((k, str(v)) for k, v in l.items()))
        if warm and key not in seen:
            run_layer(l, scale, first=(i == 0))
            seen.add(key)
        dt, _ = run_layer(l, scale, first=(i == 0))
        total += dt
    return total
