"""Import shim for `from models import resnet, resnet_spatial, resnet_spatial_d2, amoebanet, amoebanet_d2` (see ../torchgems)."""
import importlib
import sys

for _n in ("resnet", "resnet_spatial", "resnet_spatial_d2", "amoebanet", "amoebanet_d2"):
    _m = importlib.import_module("mpi4dl_b200.models." + _n)
    sys.modules[__name__ + "." + _n] = _m
    globals()[_n] = _m
