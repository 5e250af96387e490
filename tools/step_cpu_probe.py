"""Is a small-batch sampling step bound by the host?  Times (a) the CPU cost of issuing one DDIM step (graph replay +
randn + K4 through ctypes) with the GPU idle-waiting excluded, (b) the device time per step, at batch 1 / 2 / 8."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fidm_b200 as F
from fidm_b200.utils.synth import synth_batch, synth_state_dict

dev = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "ADM256"
cfg = F.CONFIGS[name]
S = cfg["image_size"]
m = F.DiffusionInpaintingModel(F.UNetModel(**dict(cfg, in_channels=3)))
m.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
m.to(dev)
d = F.create_gaussian_diffusion(steps=100, learn_sigma=True, noise_schedule="cosine")
for B in (1, 2, 8):
    data = synth_batch(B, S, seed=1, device=dev)
    fn = F.InpaintingModelFn(m)
    kw = dict(model_kwargs={"gt": data["gt"], "gt_keep_mask": data["gt_keep_mask"]}, device=dev, eta=0.0,
              use_inpainting_injection=True)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = d.ddim_sample_loop(fn, (B, 3, S, S), **kw)
        e1.record()
        t_issue = time.perf_counter() - t0          # the host has issued all 100 steps (the device may lag behind)
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
    print(f"{name} B={B}: host issue {t_issue * 10:.3f} ms/step, wall {t_all * 10:.3f} ms/step, device {e0.elapsed_time(e1) / 100:.3f} ms/step",
          flush=True)
