"""Import shim: with `mpi4dl_b200/dropin` on PYTHONPATH (ahead of the reference's `src/`), the
reference's own scripts -- `from torchgems import parser`, `from torchgems.train_spatial import ...`,
`import torchgems.comm as gems_comm` -- resolve to this repository's modules, unmodified."""
import importlib
import sys

for _n in ("spatial", "comm", "mp_pipeline", "train_spatial", "train_spatial_master", "gems_master", "parser", "utils",
           "halo_transport"):
    _m = importlib.import_module("mpi4dl_b200.torchgems." + _n)
    sys.modules[__name__ + "." + _n] = _m
    globals()[_n] = _m
