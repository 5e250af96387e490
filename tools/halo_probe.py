"""K1h (GroupNorm fused into the conv operand path) vs K1 + GroupNorm apply: sustained time, power, SM clock, and the
per-role cycle counters of the K1h kernel (fidm_conv_set_profile_buffer)."""
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pynvml
import torch

import fidm_b200 as F  # noqa: F401
from fidm_b200 import _lib as L
from fidm_b200 import ops

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda:0"


class Sampler:
    def __init__(self):
        self.p, self.c, self.stop = [], [], threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            self.c.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.02)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join()


def sustained(name, fn, seconds, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 0
    with Sampler() as s:
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < seconds:
            for _ in range(20):
                fn()
            n += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    half = len(s.p) // 2
    pw = sum(s.p[half:]) / max(1, len(s.p) - half)
    ck = sum(s.c[half:]) / max(1, len(s.c) - half)
    print(f"{name:52s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s  {pw:5.0f} W  sm {ck:5.0f} MHz", flush=True)
    return ms


def layer(B, H, W, Cin, Cout, seconds):
    xb = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    w16 = ops.repack_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(Cin * 9), torch.float16)
    bias = torch.zeros(Cout, device=dev)
    y = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    a16 = torch.empty(B, H, W, Cin, device=dev, dtype=torch.float16)
    gamma, beta = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
    flops = 2.0 * B * H * W * Cin * Cout * 9
    tag = f"{Cin}->{Cout} @{H}x{W} B{B}"

    def two_pass():
        ops.groupnorm_silu(xb, gamma, beta, out=a16)
        ops.conv2d(a16, w16, bias, out=y, impl="tc")

    coef = ops.groupnorm_silu_coeff(xb, gamma, beta)
    sustained(f"K1 conv only (normalized fp16 in)   {tag}", lambda: ops.conv2d(a16, w16, bias, out=y, impl="tc"), seconds, flops)
    sustained(f"GN stats+apply then K1              {tag}", two_pass, seconds, flops)
    sustained(f"K1h conv (GN in operand path)       {tag}", lambda: ops.conv2d(xb, w16, bias, out=y, impl="tc", gn_coef=coef),
              seconds, flops)
    # role counters of one launch
    prof = torch.zeros(16 * 148, device=dev, dtype=torch.int64)
    L.lib().fidm_conv_set_profile_buffer(L.ptr(prof))
    ops.conv2d(xb, w16, bias, out=y, impl="tc", gn_coef=coef)
    torch.cuda.synchronize()
    L.lib().fidm_conv_set_profile_buffer(None)
    pr = prof.view(148, 16).cpu().double()
    swapped = Cout % 256 != 0          # K1s: one CTA per SM, every CTA issues MMAs; epilogue slots: total, wait acc, wait buffer, tcgen05.ld
    lead = pr if swapped else pr[0::2]
    m = lead.mean(0)
    print(f"   MMA issuer   total {m[0]:9.0f} clk | wait accumulator {m[1] / m[0] * 100:5.1f}%  wait operand copy "
          f"{m[2] / m[0] * 100:5.1f}%  wait weight stage {m[3] / m[0] * 100:5.1f}%")
    t = pr.mean(0)
    print(f"   transform    total {t[4]:9.0f} clk | wait halo tile {t[5] / t[4] * 100:5.1f}%  wait free copy "
          f"{t[6] / t[4] * 100:5.1f}%  load+activate {t[7] / t[4] * 100:5.1f}%")
    extra = f"  wait staging buffer {t[10] / t[8] * 100:5.1f}%  tcgen05.ld {t[11] / t[8] * 100:5.1f}%" if swapped else ""
    print(f"   epilogue     total {t[8]:9.0f} clk | wait accumulator {t[9] / t[8] * 100:5.1f}%{extra}", flush=True)


secs = float(os.environ.get("PROBE_SECONDS", "2.0"))
if "--ref" in sys.argv:       # the 128-wide layers of REF-FFHQ256
    layer(8, 256, 256, 128, 128, secs)
    layer(8, 256, 256, 256, 128, secs)
    layer(8, 128, 128, 128, 128, secs)
    sys.exit(0)
layer(8, 256, 256, 256, 256, secs)
layer(8, 256, 256, 512, 256, secs)
layer(8, 128, 128, 512, 512, secs)
layer(8, 64, 64, 512, 512, secs)
