"""Smallest K1h cases (one CTA pair; up / head variants) for `compute-sanitizer --tool memcheck`."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as Fn

import fidm_b200 as F  # noqa: F401
from fidm_b200 import ops

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)


def act(x, gamma, beta):
    return Fn.silu(Fn.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5))


for (B, H, W, Cin, Cout, up, head) in [(1, 16, 16, 64, 128, False, False), (2, 32, 32, 128, 256, False, False),
                                       (1, 16, 16, 64, 128, True, False), (1, 16, 16, 64, 6, False, True)]:
    hs, ws = (H // 2, W // 2) if up else (H, W)
    x = (torch.randn(B, hs, ws, Cin, device=dev, generator=g) * 1.5).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) / math.sqrt(Cin * 9)).half()
    gamma, beta = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
    coef = ops.groupnorm_silu_coeff(x, gamma, beta)
    a = act(x, gamma, beta)
    if up:
        a = Fn.interpolate(a, scale_factor=2, mode="nearest")
    want = Fn.conv2d(a, w.float(), None, padding=1)
    if head:
        y = ops.conv2d(x, ops.repack_weight(w.float(), torch.float16, cout_pad=16), torch.zeros(16, device=dev),
                       nchw_out_channels=6, impl="tc", gn_coef=coef)
    else:
        y, _ = ops.conv2d(x, ops.repack_weight(w.float(), torch.float16), None, impl="tc", gn_coef=coef, x_half_res=up,
                          want_chansum=True)
        y = y.float().permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    rel = ((y - want).norm() / want.norm()).item()
    print(f"B={B} {H}x{W} {Cin}->{Cout} up={up} head={head}: rel-L2 {rel:.2e}", flush=True)
    assert rel < 4e-3
print("ok")
