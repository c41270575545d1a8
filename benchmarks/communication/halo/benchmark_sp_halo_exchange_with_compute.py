"""Halo exchange followed by the convolution it feeds (1 -> 256 channels, ones), timed -- the reference's
benchmark_sp_halo_exchange_with_compute.py (its halo_bench_pt2pt.run = start/end_halo_exchange + nn.Conv2d.forward,
:392-397).  Here that pair IS conv_spatial.forward, so this is benchmark_sp_halo_exchange_conv.py with the
reference script's fixed layer (in_channels 1, out_channels 256) and its flags."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from benchmark_sp_halo_exchange_conv import main  # noqa: E402

if __name__ == "__main__":
    sys.argv += ["--in-channels", "1", "--out-channels", "256"]
    sys.exit(main())
