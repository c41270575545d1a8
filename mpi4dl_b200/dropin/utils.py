"""`from utils import isPowerTwo, get_depth` (the reference puts src/torchgems on sys.path)."""
from mpi4dl_b200.torchgems.utils import *  # noqa: F401,F403
from mpi4dl_b200.torchgems.utils import get_depth, isPowerTwo  # noqa: F401
