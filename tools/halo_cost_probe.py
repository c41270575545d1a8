"""Cost of the boundary fix-up (direct kernel on thin strips) next to the tcgen05 interior pass."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from mpi4dl_b200 import _lib
dev = "cuda:0"
L = _lib.lib()
sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# tiles of a square-4 split of the 8192^2 workload
for (Cc, K, R, S, st, H, W) in [(104, 104, 1, 7, 1, 512, 512), (104, 104, 7, 1, 1, 512, 512), (52, 52, 3, 3, 2, 2048, 2048), (104, 104, 3, 3, 2, 1024, 1024), (3, 104, 3, 3, 2, 4096, 4096)]:
    hh, hw = (R - 1) // 2, (S - 1) // 2
    x = torch.randn(1, Cc, H, W, device=dev).to(torch.bfloat16)
    w = torch.randn(K, Cc, R, S, device=dev).to(torch.bfloat16)
    Ho, Wo = H // st, W // st
    y = torch.empty(1, K, Ho, Wo, device=dev, dtype=torch.bfloat16)
    gy = torch.randn_like(y)
    dw = torch.zeros(K, Cc, R, S, device=dev)
    d = _lib.ConvDesc(1, Cc, H, W, K, R, S, st, st, hh, hw, _lib.SPC_BF16, 0)
    nb = max(L.spc_conv_workspace_bytes(C.byref(d), i) for i in range(3)); ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
    strips = [None] * 9
    # rank 0 of a 2x2 grid: right (5), bottom (7), bottom-right (8) neighbours (pruned by kernel shape)
    if hw: strips[5] = torch.randn(1, Cc, H, hw, device=dev).to(torch.bfloat16)
    if hh: strips[7] = torch.randn(1, Cc, hh, W, device=dev).to(torch.bfloat16)
    if hh and hw: strips[8] = torch.randn(1, Cc, hh, hw, device=dev).to(torch.bfloat16)
    halo = _lib.make_halo(strips)
    f0 = ev(lambda: _lib.check(L.spc_conv2d_fwd(C.byref(d), vp(x), None, vp(w), None, vp(y), vp(ws), nb, sp()), "f"))
    f1 = ev(lambda: _lib.check(L.spc_conv2d_fwd(C.byref(d), vp(x), C.byref(halo), vp(w), None, vp(y), vp(ws), nb, sp()), "f"))
    w0 = ev(lambda: _lib.check(L.spc_conv2d_wgrad(C.byref(d), vp(x), None, vp(gy), vp(dw), None, 0, vp(ws), nb, sp()), "w"))
    w1 = ev(lambda: _lib.check(L.spc_conv2d_wgrad(C.byref(d), vp(x), C.byref(halo), vp(gy), vp(dw), None, 0, vp(ws), nb, sp()), "w"))
    print("%d->%d %dx%d s%d tile %dx%d: fprop %.3f -> %.3f ms with halo; wgrad %.3f -> %.3f ms" % (Cc, K, R, S, st, H, W, f0, f1, w0, w1), flush=True)
