"""A few launches of K1h (conv 3x3 256->256 @ 256x256, batch 8, GroupNorm+SiLU in the operand path) for
`ncu --set full -k regex:conv_halo`."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fidm_b200 as F  # noqa: F401
from fidm_b200 import ops

dev = "cuda:0"
B, H, W = 8, 256, 256
Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 256          # 128: the swapped-role kernel (conv_halo_swap.cu)
x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
FP8 = len(sys.argv) > 3 and sys.argv[3] == "fp8"          # the opt-in e4m3 operand path (kind::f8f6f4)
w32 = torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(Cin * 9)
w, scale = (ops.quantize_weight_e4m3(ops.repack_weight(w32, torch.float32)) if FP8
            else (ops.repack_weight(w32, torch.float16), None))
b = torch.zeros(Cout, device=dev)
y = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
coef = ops.groupnorm_silu_coeff(x, torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev))
for _ in range(6):
    ops.conv2d(x, w, b, out=y, impl="tc", gn_coef=coef, want_chansum=True, w_scale=scale)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
