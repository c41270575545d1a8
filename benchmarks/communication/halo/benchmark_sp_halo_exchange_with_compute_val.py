"""benchmark_sp_halo_exchange_with_compute.py with the reference's validation switched on
(benchmark_sp_halo_exchange_with_compute_val.py: received halos and the convolution output are compared with
the arange known answers, :572-700)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from benchmark_sp_halo_exchange_conv import main  # noqa: E402

if __name__ == "__main__":
    sys.argv += ["--in-channels", "1", "--out-channels", "256", "--enable-val-recv-tensors", "--enable-val-conv"]
    sys.exit(main())
