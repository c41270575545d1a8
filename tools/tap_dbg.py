import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from mpi4dl_b200 import _lib
from mpi4dl_b200.torchgems.spatial import _ConvSpatialFn
dev = "cuda:0"
C, K, R, S, H, W = [int(a) for a in sys.argv[1:7]]
x = torch.randn(1, C, H, W, device=dev).to(torch.bfloat16)
w = (torch.randn(K, C, R, S, device=dev) / (C * R * S) ** 0.5).to(torch.bfloat16)
desc = (1, C, H, W, K, R, S, 1, 1, (R - 1) // 2, (S - 1) // 2, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
with torch.no_grad():
    y = _ConvSpatialFn.apply(x, w, None, desc, *([None] * 9))
torch.cuda.synchronize()
import torch.nn.functional as F
ref = F.conv2d(x.float(), w.float(), None, 1, ((R - 1) // 2, (S - 1) // 2))
print("err", (y.float() - ref).abs().max().item(), ref.abs().max().item())
