"""BASELINE config #4 evidence: ADM256 UNet eval time vs per-GPU batch (1..32) and DDIM-50 images/s implied."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fidm_b200 as F
from fidm_b200.utils.synth import synth_batch, synth_state_dict

dev = "cuda:0"
cfg = F.CONFIGS["ADM256"]
m = F.DiffusionInpaintingModel(F.UNetModel(**dict(cfg, in_channels=3)))
m.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
m.to(dev)
fn = F.InpaintingModelFn(m)
rows = []
for B in [int(b) for b in (sys.argv[1:] or [1, 2, 4, 8, 16, 32])]:
    data = synth_batch(B, 256, seed=B, device=dev)       # procedural masks, 5-60 % holes
    x = torch.randn(B, 3, 256, 256, device=dev)
    t = torch.full((B,), 25, device=dev)
    for _ in range(4):
        fn(x, t, gt=data["gt"], gt_keep_mask=data["gt_keep_mask"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(8, 64 // B)
    e0.record()
    for _ in range(n):
        fn(x, t, gt=data["gt"], gt_keep_mask=data["gt_keep_mask"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    rows.append({"batch": B, "ms_per_eval": ms, "ms_per_image_eval": ms / B, "tflops": 2241.48 * B / ms,
                 "ddim50_images_per_s": B / (50 * ms / 1e3), "hole_fraction": float(data["mask"].mean())})
    print(json.dumps(rows[-1]), flush=True)
    m.base_model._plans.clear()
    torch.cuda.empty_cache()
with open(os.path.join(ROOT, "gpurun_out", "batch_sweep_adm256.json"), "w") as f:
    json.dump(rows, f, indent=1)
