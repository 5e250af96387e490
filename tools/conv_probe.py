"""Sustained TFLOP/s of conv shapes (tensor-core kernel) under the power cap."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fidm_b200 as F
from fidm_b200 import ops
dev = "cuda:0"

def run(name, B, H, W, Cin, Cout, ks, dtype=torch.bfloat16, seconds=2.0):
    x = torch.randn(B, H, W, Cin, device=dev).to(dtype)
    w = ops.repack_weight(torch.randn(Cout, Cin, ks, ks, device=dev) / math.sqrt(Cin * ks * ks), dtype)
    b = torch.zeros(Cout, device=dev)
    y = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv2d(x, w, b, out=y, impl="tc")
    torch.cuda.synchronize()
    n = 0
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            ops.conv2d(x, w, b, out=y, impl="tc")
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * W * Cin * Cout * ks * ks
    print(f"{name:46s} {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)

print("CTA_PAIR =", os.environ.get("FIDM_CONV_CTA_PAIR", "1"))
run("3x3 256->256 @256^2 B8 (K=2304)", 8, 256, 256, 256, 256, 3)
run("1x1 2304->256 @256^2 B8 (pure GEMM, K=2304)", 8, 256, 256, 2304, 256, 1)
run("3x3 512->256 @256^2 B8 (K=4608)", 8, 256, 256, 512, 256, 3)
run("3x3 512->512 @128^2 B8", 8, 128, 128, 512, 512, 3)
run("3x3 512->512 @64^2 B8", 8, 64, 64, 512, 512, 3)
run("3x3 1024->1024 @16^2 B8", 8, 16, 16, 1024, 1024, 3)
run("3x3 1024->1024 @8^2 B8", 8, 8, 8, 1024, 1024, 3)
