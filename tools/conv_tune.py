"""Times every (N tile, split) candidate of K1 on the low-resolution layer shapes against the cost model's own pick
(graph replay of 50 launches, no profiler).  One subprocess per candidate: FIDM_CONV_FORCE is read per call but the
cluster-capacity cache and CUDA context are per process.  Usage: python tools/conv_tune.py [B ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import json, math, os, sys
sys.path.insert(0, sys.argv[1])
import torch
import fidm_b200
from fidm_b200 import ops
shapes = json.loads(sys.argv[2])
dev = "cuda:0"
out = {}
for (B, H, Cin, Cout, ks) in shapes:
    x = torch.randn(B, H, H, Cin, device=dev).half()
    w = ops.repack_weight(torch.randn(Cout, Cin, ks, ks, device=dev) / math.sqrt(Cin * ks * ks), torch.float16)
    b = torch.zeros(Cout, device=dev)
    y = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.conv2d(x, w, b, out=y, impl="tc")
    try:
        fn(); torch.cuda.synchronize()
    except ValueError:
        out[str((B, H, Cin, Cout, ks))] = float("inf")
        continue
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(50): fn()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    out[str((B, H, Cin, Cout, ks))] = e0.elapsed_time(e1) / 250 * 1e3
print("RESULT " + json.dumps(out))
"""

Bs = [int(v) for v in sys.argv[1:]] or [1, 8]
shapes = []
for B in Bs:
    shapes += [(B, 8, 512, 512, 3), (B, 8, 1024, 1024, 3), (B, 16, 512, 512, 3), (B, 16, 1024, 1024, 3), (B, 16, 1024, 512, 3),
               (B, 32, 256, 256, 3), (B, 32, 512, 512, 3), (B, 32, 512, 1536, 1), (B, 64, 512, 512, 3) if B <= 2 else (B, 16, 512, 1536, 1)]


def run(force):
    env = dict(os.environ)
    if force:
        env["FIDM_CONV_FORCE"] = force
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, json.dumps(shapes)], env=env, capture_output=True, text=True, timeout=600)
    for line in r.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    raise RuntimeError(r.stderr[-1500:])


base = run(None)
cands = [f"{n},{s}" for n in (64, 128, 256) for s in (1, 2, 4, 8)]
res = {c: run(c) for c in cands}
for sh in base:
    best = min(cands, key=lambda c: res[c][sh])
    row = "  ".join(f"{c}:{res[c][sh]:5.1f}" for c in cands)
    flag = "" if base[sh] <= res[best][sh] * 1.05 else f"   <-- model pick is {base[sh] / res[best][sh]:.2f}x the best ({best})"
    print(f"{sh:28s} model {base[sh]:5.1f} us | {row}{flag}", flush=True)
