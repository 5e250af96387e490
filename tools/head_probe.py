"""Time the fused head (GroupNorm+SiLU + conv3x3 256->6, fp32 NCHW out) alone: K1h BLOCK_N = 16."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fidm_b200 as F  # noqa: F401
from fidm_b200 import ops
dev = "cuda:0"
B, H, W, Cin = 8, 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
w = ops.repack_weight(torch.randn(6, Cin, 3, 3, device=dev) / math.sqrt(Cin * 9), torch.float16, cout_pad=16)
b = torch.zeros(16, device=dev)
coef = ops.groupnorm_silu_coeff(x, torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev))
for _ in range(5):
    ops.conv2d(x, w, b, nchw_out_channels=6, impl="tc", gn_coef=coef)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.conv2d(x, w, b, nchw_out_channels=6, impl="tc", gn_coef=coef)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 50 * 1e3
print(f"head K1h<16> Cin={Cin} B={B} {H}x{W}: {us:.1f} us/launch, {x.numel() * 2 / us / 1e6:.2f} TB/s of raw input")
