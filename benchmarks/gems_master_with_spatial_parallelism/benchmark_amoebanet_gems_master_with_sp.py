"""amoebanet variant of benchmark_gems_master_with_sp.py (same flags as the reference's script of this name)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from benchmark_gems_master_with_sp import main  # noqa: E402

if __name__ == "__main__":
    main("amoebanet")
