"""Import shims that let the UNMODIFIED reference (/root/reference/src) run on CPU over gloo.

Used ONLY by the fixture generators in tools/ (run in the build container, where
/root/reference exists).  Nothing in the product, the tests' run-time path, smoke() or
bench.py imports this file.  The six shims are the ones listed in SURVEY.md section 8c:

  1. sys.path += reference src/ and src/torchgems/   (train_spatial.py does `from utils import ...`)
  2. dist.init_process_group forced to backend="gloo" (reference hard-codes "mpi", comm.py:73)
  3. torch.cuda.{synchronize,init,ipc_collect,empty_cache} -> no-ops
  4. torch.zeros(..., device="cuda") -> cpu          (spatial.py:368-375)
  5. Module.to("cuda:0") / Tensor.to("cuda:0") / Tensor.cuda() -> identity
  6. (caller) ready_model(..., GET_SHAPES_ON_CUDA=False)
"""
import os
import sys

import torch
import torch.distributed as dist

REF_ROOT = os.environ.get("MPI4DL_REFERENCE", "/root/reference")


def install():
    for p in (os.path.join(REF_ROOT, "src"), os.path.join(REF_ROOT, "src", "torchgems")):
        if p not in sys.path:
            sys.path.insert(0, p)

    if getattr(torch, "_mpi4dl_shimmed", False):
        return
    torch._mpi4dl_shimmed = True

    _init = dist.init_process_group

    def init_pg(backend=None, *a, **k):
        return _init("gloo", *a, **k)

    dist.init_process_group = init_pg

    noop = lambda *a, **k: None
    torch.cuda.synchronize = noop
    torch.cuda.init = noop
    torch.cuda.ipc_collect = noop
    torch.cuda.empty_cache = noop

    def _is_cuda(dev):
        return dev is not None and str(dev).startswith("cuda")

    def _wrap_factory(name):
        orig = getattr(torch, name)

        def f(*a, **k):
            if _is_cuda(k.get("device")):
                k["device"] = "cpu"
            return orig(*a, **k)

        setattr(torch, name, f)

    for n in ("zeros", "ones", "empty", "rand", "randn", "tensor"):
        _wrap_factory(n)

    _t_to = torch.Tensor.to

    def t_to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, (str, torch.device)) and _is_cuda(x)) else x for x in a)
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        return _t_to(self, *a, **k)

    torch.Tensor.to = t_to
    torch.Tensor.cuda = lambda self, *a, **k: self

    _m_to = torch.nn.Module.to

    def m_to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, (str, torch.device)) and _is_cuda(x)) else x for x in a)
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        return _m_to(self, *a, **k)

    torch.nn.Module.to = m_to
    torch.nn.Module.cuda = lambda self, *a, **k: self


def init_single_process_group(port=29533):
    """world_size-1 gloo group: enough for num_spatial_parts=1 model construction."""
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(port))
        dist.init_process_group("gloo", rank=0, world_size=1)
