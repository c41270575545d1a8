"""Dev probe (GPU): conv_tap.cu vs the round-1 shifted-copy path (SPC_TAP_V1=1) on the BASELINE tap shapes:
correctness against each other and CUDA-event time of fprop / dgrad.
    python tools/tap_probe.py [--quick]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpi4dl_b200 import _lib  # noqa: E402

SHAPES = [  # C, K, R, S, H, W
    (104, 104, 1, 7, 1024, 1024), (104, 104, 7, 1, 1024, 1024), (52, 52, 1, 7, 2048, 2048), (52, 52, 7, 1, 2048, 2048),
    (16, 16, 3, 3, 4096, 4096), (64, 16, 3, 3, 4096, 4096), (64, 64, 3, 3, 2048, 2048), (128, 64, 3, 3, 2048, 2048),
    (3, 16, 3, 3, 4096, 4096),
]


def main():
    quick = "--quick" in sys.argv
    only = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--only=")]      # shape index, v2 only (for ncu)
    L = _lib.lib()
    dev = "cuda:0"
    sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    for (Cc, K, R, S, H, W) in ([SHAPES[i] for i in only] if only else (SHAPES[:2] if quick else SHAPES)):
        torch.manual_seed(0)
        x = torch.randn(1, Cc, H, W, device=dev).to(torch.bfloat16)
        gy = torch.randn(1, K, H, W, device=dev).to(torch.bfloat16)
        w = (torch.randn(K, Cc, R, S, device=dev) / (Cc * R * S) ** 0.5).to(torch.bfloat16)
        y = torch.empty(1, K, H, W, device=dev, dtype=torch.bfloat16)
        dx = torch.empty_like(x)
        dw = torch.empty(K, Cc, R, S, device=dev, dtype=torch.float32)
        d = _lib.ConvDesc(1, Cc, H, W, K, R, S, 1, 1, (R - 1) // 2, (S - 1) // 2, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
        res = {}
        for mode in (("v2",) if only else ("v2", "v1")):
            if mode == "v1":
                os.environ["SPC_TAP_V1"] = "1"
            else:
                os.environ.pop("SPC_TAP_V1", None)
            L.spc_reload_env()
            nb = max(L.spc_conv_workspace_bytes(C.byref(d), i) for i in range(3))
            ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
            fns = {"fprop": lambda: _lib.check(L.spc_conv2d_fwd(C.byref(d), x.data_ptr(), None, w.data_ptr(), None, y.data_ptr(),
                                                                ws.data_ptr(), nb, sp()), "fwd"),
                   "dgrad": lambda: _lib.check(L.spc_conv2d_dgrad(C.byref(d), gy.data_ptr(), w.data_ptr(), dx.data_ptr(),
                                                                  ws.data_ptr(), nb, sp()), "dgrad"),
                   "wgrad": lambda: _lib.check(L.spc_conv2d_wgrad(C.byref(d), x.data_ptr(), None, gy.data_ptr(), dw.data_ptr(), None,
                                                                  0, ws.data_ptr(), nb, sp()), "wgrad")}
            for nm, fn in fns.items():
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[(mode, nm)] = (e0.elapsed_time(e1) / 5, {"fprop": y, "dgrad": dx, "wgrad": dw}[nm].float().clone())
            del ws
        gb = (Cc + K) * H * W * 2 / 1e9
        tf = 2.0 * Cc * K * R * S * H * W / 1e12
        for nm in ("fprop", "dgrad", "wgrad"):
            if only:
                print("%4d->%-4d %dx%d @%dx%d %-5s  v2 %7.3f ms" % (Cc, K, R, S, H, W, nm, res[("v2", nm)][0]))
                continue
            a, b = res[("v2", nm)], res[("v1", nm)]
            err = float((a[1] - b[1]).abs().max())
            ref = float(b[1].abs().max())
            print("%4d->%-4d %dx%d @%dx%d %-5s  v2 %7.3f ms (%5.2f TB/s %6.1f TF/s)   v1 %7.3f ms   x%.2f   maxdiff %.3g / %.3g %s" % (
                Cc, K, R, S, H, W, nm, a[0], gb / a[0], tf / a[0] * 1e3, b[0], b[0] / a[0], err, ref,
                "" if err <= 0.02 * ref else "MISMATCH"))
    os.environ.pop("SPC_TAP_V1", None)


def breakdown():
    """Timing experiments with parts of the kernel switched off (SPC_TAP_DBG; results are garbage)."""
    L = _lib.lib()
    dev = "cuda:0"
    sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    for idx in (1, 0, 4, 6):
        Cc, K, R, S, H, W = SHAPES[idx]
        x = torch.randn(1, Cc, H, W, device=dev).to(torch.bfloat16)
        w = (torch.randn(K, Cc, R, S, device=dev) / (Cc * R * S) ** 0.5).to(torch.bfloat16)
        y = torch.empty(1, K, H, W, device=dev, dtype=torch.bfloat16)
        d = _lib.ConvDesc(1, Cc, H, W, K, R, S, 1, 1, (R - 1) // 2, (S - 1) // 2, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
        nb = L.spc_conv_workspace_bytes(C.byref(d), 0)
        ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
        out = []
        for dbg in (0, 1, 2, 4, 8, 3, 7, 15, 11):
            os.environ["SPC_TAP_DBG"] = str(dbg)
            fn = lambda: _lib.check(L.spc_conv2d_fwd(C.byref(d), x.data_ptr(), None, w.data_ptr(), None, y.data_ptr(),  # noqa: E731
                                                     ws.data_ptr(), nb, sp()), "fwd")
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out.append("dbg%-2d %.3f" % (dbg, e0.elapsed_time(e1) / 5))
        os.environ.pop("SPC_TAP_DBG", None)
        print("%4d->%-4d %dx%d @%dx%d  (1=noW 2=noX 4=noStore 8=noShift)  " % (Cc, K, R, S, H, W) + "  ".join(out))


if __name__ == "__main__":
    breakdown() if "--breakdown" in sys.argv else main()
