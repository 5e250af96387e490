"""GPU probe: rel-L2 of one UNet evaluation (bf16 tensor-core path, fp32 FFMA path) against the CPU
oracle, next to what plain PyTorch bf16 (same oracle code, bf16 tensors on the GPU) achieves."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fidm_b200 as F
from fidm_b200.utils.synth import synth_batch, synth_state_dict
from oracle import unet_oracle as uor

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
names = sys.argv[1:] or ["T64", "REF_FFHQ256", "ADM256"]
for name in names:
    cfg = F.CONFIGS[name]
    S = cfg["image_size"]
    sd = synth_state_dict(cfg, seed=7)
    data = synth_batch(1, S, seed=2)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 3, S, S, generator=g)
    for tval in (5, 61):
        t = torch.tensor([tval])
        sdg = {k: v.to(dev) for k, v in sd.items()}
        with torch.no_grad():
            t0 = time.time()
            want = uor.inpaint_forward(sdg, cfg, x.to(dev), t.to(dev), data["masked_image"].to(dev), data["mask"].to(dev)).cpu()
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                pt16 = uor.inpaint_forward(sdg, cfg, x.to(dev), t.to(dev), data["masked_image"].to(dev),
                                           data["mask"].to(dev)).float().cpu()
        res = {}
        for prec in ("bf16", "fp32"):
            m = F.DiffusionInpaintingModel(F.UNetModel(**dict(cfg, in_channels=3)))
            m.load_state_dict(sd, strict=True)
            m.to(dev)
            m.base_model.set_precision(prec)
            out = m(x.to(dev), t.to(dev), masked_image=data["masked_image"].to(dev), mask=data["mask"].to(dev)).cpu()
            res[prec] = out
            del m
            torch.cuda.empty_cache()
        rl = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
        print(f"{name} t={tval}: ours bf16 all6 {rl(res['bf16'], want):.3e} eps {rl(res['bf16'][:, :3], want[:, :3]):.3e} | "
              f"ours fp32 {rl(res['fp32'], want):.3e} | torch-autocast-bf16 {rl(pt16, want):.3e}", flush=True)
