"""Sweep the wgrad tiling knobs (SPC_WG_NBLK / _MG / _STAGES / _SPLITS, SPC_WG_GROUP_MAJOR,
SPC_WG_ROWS128) over the heavy AmoebaNet-D wgrad shapes and print ms / TB/s / TFLOP/s per
configuration.  One GPU, ~1 minute:   python tools/wgrad_probe.py [--quick]
                                   timeout 90 python tools/wgrad_probe.py --pair --quick   (CTA-pair kernel)

What it is for: the 1664->416 @1024^2 wgrad runs at 2.5 ms with DRAM 44 %, tensor pipe 44 % and
L2->SM 5.8 TB/s (profiles/r1b_ncu_full_summary.csv) -- nothing saturated.  The sweep separates the
candidates: bytes in flight (stages x stage size via NBLK/MG), HBM re-reads (group order, splits),
accumulator shape (NBLK x MG)."""
import ctypes as C
import itertools
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from mpi4dl_b200 import _lib  # noqa: E402

SHAPES = [(1664, 416, 1024), (416, 416, 1024), (624, 416, 2048), (104, 208, 4096), (1248, 416, 1024), (104, 416, 1024),
          (416, 104, 1024), (208, 208, 2048), (208, 52, 4096), (52, 208, 2048)]
KNOBS = ["SPC_WG_NBLK", "SPC_WG_MG", "SPC_WG_STAGES", "SPC_WG_SPLITS", "SPC_WG_GROUP_MAJOR", "SPC_WG_ROWS128"]
CONFIGS = [{}] + [dict(SPC_WG_NBLK=str(n), SPC_WG_MG=str(m)) for n, m in itertools.product((128, 192, 256), (1, 2, 4))] + [
    dict(SPC_WG_STAGES="2"), dict(SPC_WG_SPLITS="10"), dict(SPC_WG_SPLITS="42"), dict(SPC_WG_GROUP_MAJOR="1"),
    dict(SPC_WG_ROWS128="1")]


PAIR_CONFIGS = [{}, dict(SPC_WG_2CTA="1"), dict(SPC_WG_2CTA="1", SPC_WG_STAGES="3"), dict(SPC_WG_2CTA="1", SPC_WG_NBLK="128"),
                dict(SPC_WG_2CTA="1", SPC_WG_NBLK="192")]


def main():
    quick = "--quick" in sys.argv
    global CONFIGS
    global SHAPES
    if "--splits" in sys.argv:
        # flush cost: every work item adds its K x C accumulators to dw with fp32 atomics; sweep the item count
        CONFIGS = [{}] + [dict(SPC_WG_SPLITS=str(v)) for v in (148, 74, 37, 18, 296)]
        SHAPES = [(416, 416, 1024), (1664, 416, 1024), (104, 416, 1024), (104, 416, 360), (416, 416, 360), (416, 104, 360),
                  (104, 208, 4096), (1248, 416, 360)]
    if "--wide" in sys.argv:
        # wide stages (two 64-pixel blocks per stage, 5-d TMA boxes) for layers with multi-page channel planes
        CONFIGS = [dict(SPC_WG_WIDE="0"), {}, dict(SPC_WG_WIDE="1", SPC_WG_BOX5="0"), dict(SPC_WG_WIDE="1", SPC_WG_BOX5="1"),
                   dict(SPC_WG_WIDE="1", SPC_WG_BOX5="2"), dict(SPC_WG_WIDE="1", SPC_WG_MG="1")]
        SHAPES = [(104, 208, 4096), (208, 52, 4096), (52, 208, 2048), (208, 208, 2048), (104, 416, 1024), (208, 104, 1024)]
        KNOBS.extend(["SPC_WG_WIDE", "SPC_WG_BOX5"])
    if "--pairwide" in sys.argv:
        CONFIGS = [{}, dict(SPC_WG_PAIR_WIDE="1"), dict(SPC_WG_PAIR_WIDE="1", SPC_WG_PAIR_MP1="1")]
        SHAPES = [(416, 416, 1024), (1664, 416, 1024), (624, 416, 2048), (1248, 416, 1024), (624, 208, 2048), (416, 104, 2048),
                  (416, 104, 1024), (104, 416, 1024)]
        KNOBS.extend(["SPC_WG_PAIR_WIDE", "SPC_WG_PAIR_MP1"])
    if "--pair" in sys.argv:
        # first run of the cta_group::2 kernel (never executed on hardware in round 1): ALWAYS under an outer
        # `timeout 90`, results are compared with the single-CTA kernel's dw (MISMATCH flag)
        CONFIGS = PAIR_CONFIGS
        KNOBS.append("SPC_WG_2CTA")
    L = _lib.lib()
    dev = "cuda:0"
    sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    for (Cc, K, H) in (SHAPES[:2] if quick else SHAPES):
        x = torch.randn(1, Cc, H, H, device=dev).to(torch.bfloat16)
        gy = torch.randn(1, K, H, H, device=dev).to(torch.bfloat16)
        dw = torch.zeros(K, Cc, 1, 1, device=dev)
        d = _lib.ConvDesc(1, Cc, H, H, K, 1, 1, 1, 1, 0, 0, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
        nb = L.spc_conv_workspace_bytes(C.byref(d), 2)
        ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
        gbytes = (Cc + K) * H * H * 2 / 1e9
        tflop = 2.0 * Cc * K * H * H / 1e12
        print("== wgrad %d->%d 1x1 @%d^2   (%.2f GB algorithmic, %.2f TFLOP)" % (Cc, K, H, gbytes, tflop))
        ref = None
        for cfg in CONFIGS:
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(cfg)
            L.spc_reload_env()

            def call():
                _lib.check(L.spc_conv2d_wgrad(C.byref(d), x.data_ptr(), None, gy.data_ptr(), dw.data_ptr(), None, 0,
                                              ws.data_ptr(), nb, sp()), "wgrad")
            try:
                for _ in range(2):
                    call()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    call()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
            except Exception as ex:  # a configuration the launcher rejects (smem budget, TMEM columns)
                print("   %-46s rejected: %s" % (cfg or "default", str(ex)[:80]))
                continue
            if ref is None:
                ref, chk = ms, dw.clone()
            ok = torch.allclose(dw, chk, rtol=2e-2, atol=2e-2 * chk.abs().max().item())
            print("   %-46s %7.3f ms  %5.2f TB/s  %6.1f TFLOP/s  x%.2f %s" % (
                cfg or "default", ms, gbytes / ms, tflop / ms * 1e3, ref / ms, "" if ok else "MISMATCH"))
    for k in KNOBS:
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
