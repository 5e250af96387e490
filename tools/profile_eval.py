"""One UNet evaluation (ADM256, batch 8 by default) between cudaProfilerStart/Stop, without CUDA graphs,
so that `ncu --profile-from-start off` lists exactly the kernels of one evaluation."""
import os
import sys

os.environ.setdefault("FIDM_CUDA_GRAPH", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fidm_b200 as F
from fidm_b200.utils.synth import synth_batch, synth_state_dict

name = sys.argv[1] if len(sys.argv) > 1 else "ADM256"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda:0"
cfg = F.CONFIGS[name]
S = cfg["image_size"]
m = F.DiffusionInpaintingModel(F.UNetModel(**dict(cfg, in_channels=3)))
m.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
m.to(dev)
data = synth_batch(B, S, seed=1, device=dev)
x = torch.randn(B, 3, S, S, device=dev)
t = torch.full((B,), 50, device=dev)
d = F.create_gaussian_diffusion(steps=100, learn_sigma=True, noise_schedule="cosine")
from fidm_b200 import _lib as L
for _ in range(2):
    out = m(x, t, masked_image=data["masked_image"], mask=data["mask"])
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = m(x, t, masked_image=data["masked_image"], mask=data["mask"])
r = d._step(L.STEP_UPDATE_INJECT, x, t=50, t_inject=49, model_out=out, gt=data["gt"], keep=data["gt_keep_mask"],
            inject_noise=torch.randn_like(x), ddim=True, want_next=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out.abs().mean()))
