"""Pointwise (1x1) fprop / dgrad epilogue experiments: TMA-store epilogue against the coalesced per-thread-store one
(SPC_PW_EPI_STG=1), each with two or one staging buffers (SPC_PW_OUTBUFS).  One GPU, < 1 minute:
    timeout 120 python tools/pw_probe.py [--quick]
Outputs of every configuration are compared bit for bit with the default configuration's."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from mpi4dl_b200 import _lib  # noqa: E402

SHAPES = [(104, 208, 4096, 4096), (208, 52, 4096, 4096), (52, 208, 2048, 2048), (416, 104, 1024, 1024), (104, 416, 1024, 1024),
          (416, 416, 1024, 1024), (1664, 416, 1024, 1024), (624, 416, 2048, 2048), (208, 208, 2048, 2048), (52, 52, 1000, 1096)]
CONFIGS = [{}, dict(SPC_PW_EPI_STG="1"), dict(SPC_PW_EPI_STG="1", SPC_PW_OUTBUFS="1"), dict(SPC_PW_OUTBUFS="1")]
KNOBS = ["SPC_PW_EPI_STG", "SPC_PW_OUTBUFS", "SPC_PW_XBOX", "SPC_PW_YBOX", "SPC_PW_BOX5", "SPC_PW_STATIONARY", "SPC_PW_MB2", "SPC_PW_N256"]
if "--stationary" in sys.argv:
    # weights resident per CTA with several groups of output channels (first line = the previous behaviour)
    CONFIGS = [dict(SPC_PW_N256="0"), {}, dict(SPC_PW_BOX5="3"), dict(SPC_PW_N256="1"), dict(SPC_PW_N256="0", SPC_PW_STATIONARY="1")]
    SHAPES = [(416, 416, 1024, 1024), (104, 416, 1024, 1024), (416, 104, 1024, 1024), (1664, 416, 1024, 1024), (1248, 416, 1024, 1024),
              (624, 416, 2048, 2048), (416, 416, 360, 1024), (104, 208, 4096, 4096)]
if "--boxes" in sys.argv:
    # channel rows per TMA box: fewer planes (2 MB pages) walked between the two 64-pixel blocks of a tile
    CONFIGS = [{}, dict(SPC_PW_BOX5="1"), dict(SPC_PW_BOX5="2"), dict(SPC_PW_BOX5="3"), dict(SPC_PW_XBOX="16"), dict(SPC_PW_YBOX="16")]
    SHAPES = [(104, 208, 4096, 4096), (208, 52, 4096, 4096), (52, 208, 2048, 2048), (208, 208, 2048, 2048), (624, 416, 2048, 2048),
              (416, 104, 1024, 1024), (104, 416, 1024, 1024), (416, 416, 1024, 1024), (52, 52, 1000, 1096)]


def main():
    L = _lib.lib()
    dev = "cuda:0"
    sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    shapes = SHAPES[:4] + SHAPES[-1:] if "--quick" in sys.argv else SHAPES
    for (Cc, K, H, W) in shapes:
        torch.manual_seed(0)
        x = torch.randn(1, Cc, H, W, device=dev).to(torch.bfloat16)
        gy = torch.randn(1, K, H, W, device=dev).to(torch.bfloat16)
        w = (torch.randn(K, Cc, 1, 1, device=dev) / Cc ** 0.5).to(torch.bfloat16)
        b = torch.randn(K, device=dev).to(torch.bfloat16)
        y = torch.empty(1, K, H, W, device=dev, dtype=torch.bfloat16)
        dx = torch.empty_like(x)
        d = _lib.ConvDesc(1, Cc, H, W, K, 1, 1, 1, 1, 0, 0, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
        nb = max(L.spc_conv_workspace_bytes(C.byref(d), i) for i in range(3))
        ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
        gb = (Cc + K) * H * W * 2 / 1e9
        print("== %d->%d 1x1 @%dx%d  (%.2f GB algorithmic)" % (Cc, K, H, W, gb), flush=True)
        ref = {}
        for cfg in CONFIGS:
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(cfg)
            L.spc_reload_env()
            fns = {"fprop": (lambda: _lib.check(L.spc_conv2d_fwd(C.byref(d), x.data_ptr(), None, w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                                                 ws.data_ptr(), nb, sp()), "fwd"), y),
                   "dgrad": (lambda: _lib.check(L.spc_conv2d_dgrad(C.byref(d), gy.data_ptr(), w.data_ptr(), dx.data_ptr(),
                                                                   ws.data_ptr(), nb, sp()), "dgrad"), dx)}
            line = "   %-44s" % (" ".join("%s=%s" % kv for kv in cfg.items()) or "default (TMA store, 2 buffers)")
            for nm, (fn, out) in fns.items():
                out.zero_()
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                if nm not in ref:
                    ref[nm] = out.clone()
                    ok = ""
                else:
                    ok = "" if torch.equal(ref[nm], out) else " MISMATCH"
                line += "  %s %7.3f ms %5.2f TB/s%s" % (nm, ms, gb / ms, ok)
            print(line, flush=True)
    for k in KNOBS:
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
