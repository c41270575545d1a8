"""Halo exchange alone (pad -> exchange -> unpack), timed and validated against the zero-padded arange
image -- the reference's benchmark_sp_halo_exchange.py; see benchmark_sp_halo_exchange_conv.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from benchmark_sp_halo_exchange_conv import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main(exchange_only_default=True))
