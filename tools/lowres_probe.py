"""True per-launch time of the low-resolution layers (CUDA-graph replay of 50 back-to-back launches, no profiler):
where does a batch-1 evaluation spend its time -- kernel bodies or launch / dependency gaps?"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fidm_b200 as F  # noqa: F401
from fidm_b200 import ops

dev = "cuda:0"


def graph_time(fn, n=50, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3          # us per launch


def conv(B, H, Cin, Cout, ks=3):
    x = torch.randn(B, H, H, Cin, device=dev).half()
    w = ops.repack_weight(torch.randn(Cout, Cin, ks, ks, device=dev) / math.sqrt(Cin * ks * ks), torch.float16)
    b = torch.zeros(Cout, device=dev)
    y = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
    us = graph_time(lambda: ops.conv2d(x, w, b, out=y, impl="tc"))
    fl = 2.0 * B * H * H * Cin * Cout * ks * ks
    wb = Cout * Cin * ks * ks * 2
    print(f"conv{ks}x{ks} {Cin:5d}->{Cout:5d} @{H:3d}^2 B{B}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  weights {wb / 1e6:5.1f} MB -> {wb / us / 1e3:7.1f} GB/s",
          flush=True)


def gn(B, H, C):
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    y = torch.empty(B, H, H, C, device=dev, dtype=torch.float16)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    us = graph_time(lambda: ops.groupnorm_silu(x, gamma, beta, out=y))
    print(f"groupnorm+silu {C:5d} ch @{H:3d}^2 B{B}: {us:7.1f} us", flush=True)


for B in (1, 8):
    conv(B, 8, 1024, 1024)
    conv(B, 16, 1024, 1024)
    conv(B, 16, 2048, 1024)
    conv(B, 32, 512, 512)
    conv(B, 32, 1024, 512)
    conv(B, 64, 512, 512)
    conv(B, 32, 512, 1536, 1)
    conv(B, 8, 1024, 1024, 1)
    gn(B, 8, 1024)
    gn(B, 32, 512)
    gn(B, 64, 512)
# an empty-ish kernel for the launch floor inside a graph
t = torch.zeros(1, device=dev)
print(f"tiny torch kernel in a graph: {graph_time(lambda: t.add_(1.0)):.2f} us per launch")
