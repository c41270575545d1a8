"""Cost of the halo fix-up next to the interior pass, per halo layer of the AmoebaNet-D spatial stage, for the tile
shapes of an N-GPU run (one GPU, strips handed in directly like the parity tests do):
    python tools/halo_cost_probe.py [N]        N = 8 (vertical strips, default) | 4 (square) | 2 (vertical)
Prints per layer fprop / wgrad time without and with the neighbours' strips, and the sum over the layer list."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402

from mpi4dl_b200 import _lib  # noqa: E402

dev = "cuda:0"
L = _lib.lib()
sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731


def ev(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    gr, gc = (2, 2) if n == 4 else (1, n)
    d = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "layers_amoebanetd_sp4.json")))
    seen, tot = {}, [0.0, 0.0, 0.0, 0.0]
    for l in d["layers"]:
        if l["op"] != "conv" or l["R"] * l["S"] == 1:
            continue
        key = (l["C"], l["K"], l["R"], l["S"], l["stride_h"], l["H"])
        if key not in seen:
            Cc, K, R, S, st = l["C"], l["K"], l["R"], l["S"], l["stride_h"]
            H, W = l["H"] // gr, l["W"] // gc
            hh, hw = (R - 1) // 2, (S - 1) // 2
            x = torch.randn(1, Cc, H, W, device=dev).to(torch.bfloat16)
            w = torch.randn(K, Cc, R, S, device=dev).to(torch.bfloat16)
            y = torch.empty(1, K, H // st, W // st, device=dev, dtype=torch.bfloat16)
            gy = torch.randn_like(y)
            dw = torch.zeros(K, Cc, R, S, device=dev)
            dsc = _lib.ConvDesc(1, Cc, H, W, K, R, S, st, st, hh, hw, _lib.SPC_BF16, 0)
            nb = max(L.spc_conv_workspace_bytes(C.byref(dsc), i) for i in range(3))
            ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
            strips = [None] * 9
            # an interior tile of the grid: every direction the kernel shape exchanges on
            for i, (dr, dc) in enumerate([(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]):
                if i == 4 or (dr and (hh == 0 or gr == 1)) or (dc and (hw == 0 or gc == 1)):
                    continue
                strips[i] = torch.randn(1, Cc, H if dr == 0 else hh, W if dc == 0 else hw, device=dev).to(torch.bfloat16)
            halo = _lib.make_halo(strips)
            f0 = ev(lambda: _lib.check(L.spc_conv2d_fwd(C.byref(dsc), vp(x), None, vp(w), None, vp(y), vp(ws), nb, sp()), "f"))
            f1 = ev(lambda: _lib.check(L.spc_conv2d_fwd(C.byref(dsc), vp(x), C.byref(halo), vp(w), None, vp(y), vp(ws), nb, sp()), "f"))
            w0 = ev(lambda: _lib.check(L.spc_conv2d_wgrad(C.byref(dsc), vp(x), None, vp(gy), vp(dw), None, 0, vp(ws), nb, sp()), "w"))
            w1 = ev(lambda: _lib.check(L.spc_conv2d_wgrad(C.byref(dsc), vp(x), C.byref(halo), vp(gy), vp(dw), None, 0, vp(ws), nb, sp()), "w"))
            seen[key] = (f0, f1, w0, w1, "%d->%d %dx%d s%d tile %dx%d" % (Cc, K, R, S, st, H, W), any(s is not None for s in strips))
        f0, f1, w0, w1, name, has = seen[key]
        for i, v in enumerate((f0, f1, w0, w1)):
            tot[i] += v
    for f0, f1, w0, w1, name, has in seen.values():
        print("%-36s fprop %.3f -> %.3f ms   wgrad %.3f -> %.3f ms %s" % (name, f0, f1, w0, w1, "" if has else "(no neighbours on the exchanged axis)"),
              flush=True)
    print("N=%d, sum over the layer list: fprop %.2f -> %.2f ms, wgrad %.2f -> %.2f ms" % (n, tot[0], tot[1], tot[2], tot[3]))


if __name__ == "__main__":
    main()
