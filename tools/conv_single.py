"""A few launches of K1 (conv_tc) on one layer for `ncu --set full -k regex:conv_tc`:
    python tools/conv_single.py B H Cin Cout [ksize]      default: 8 16 512 512 3 (cluster split-K, N tile 256, split 4)"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fidm_b200 as F  # noqa: F401
from fidm_b200 import ops

dev = "cuda:0"
a = [int(v) for v in sys.argv[1:]]
B, H, Cin, Cout, ks = (a + [8, 16, 512, 512, 3][len(a):])[:5]
x = torch.randn(B, H, H, Cin, device=dev).half()
w = ops.repack_weight(torch.randn(Cout, Cin, ks, ks, device=dev) / math.sqrt(Cin * ks * ks), torch.float16)
b = torch.zeros(Cout, device=dev)
res = torch.randn(B, H, H, Cout, device=dev).bfloat16()
emb = torch.randn(B, Cout, device=dev)
y = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
for _ in range(6):
    ops.conv2d(x, w, b, row_add=emb, residual=res, out=y, impl="tc")
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
