"""GPU evidence run: full DDIM loop at 256x256 (ADM256 or REF_FFHQ256) on the B200 vs the CPU oracle with
identical weights, inputs and noise.  Writes PSNR of the final image, known-region exactness and per-eval
eps rel-L2 along the trajectory to gpurun_out/parity_<cfg>_ddim<T>.json."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import fidm_b200 as F
from fidm_b200.utils.synth import synth_batch, synth_state_dict
from helpers import PatchedRandn, psnr, rel_l2, seeded_noise
from oracle import diffusion_oracle as dor
from oracle import unet_oracle as uor

name = sys.argv[1] if len(sys.argv) > 1 else "ADM256"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sched = sys.argv[3] if len(sys.argv) > 3 else "cosine"
dev = "cuda:0"
cfg = F.CONFIGS[name]
S = cfg["image_size"]
sd = synth_state_dict(cfg, seed=11)
data = synth_batch(1, S, seed=12)
gt, keep = data["gt"], data["gt_keep_mask"]
seed, shape = 5, (1, 3, S, S)
torch.set_num_threads(os.cpu_count())

m = F.DiffusionInpaintingModel(F.UNetModel(**dict(cfg, in_channels=3)))
m.load_state_dict(sd, strict=True)
m.to(dev)
fn = F.InpaintingModelFn(m)
d = F.create_gaussian_diffusion(steps=T, learn_sigma=True, noise_schedule=sched)
t0 = time.time()
gpu_trace = []
with PatchedRandn(T, seed, device=dev):
    for o in d.ddim_sample_loop_progressive(fn, shape, model_kwargs={"gt": gt.to(dev), "gt_keep_mask": keep.to(dev)},
                                            device=dev, eta=0.0, use_inpainting_injection=True):
        gpu_trace.append({k: v.cpu() for k, v in o.items()})
torch.cuda.synchronize()
t_gpu = time.time() - t0

sdg = sd
tab = dor.Tables(F.get_named_beta_schedule(sched, T))
trace = []
t0 = time.time()
with torch.no_grad():
    ref = dor.sample_loop(tab, lambda x, t, **k: uor.inpaint_forward(sdg, cfg, x, t, gt * keep, 1 - keep), shape,
                          ddim=True, x_T=seeded_noise("xT", 0, shape, seed), gt=gt, keep=keep,
                          noise_fn=lambda kind, t: seeded_noise(kind, t, shape, seed), trace=trace)
t_cpu = time.time() - t0
final = gpu_trace[-1]["sample"]
hole = (1 - keep).expand_as(ref) == 1
res = {
    "config": name, "ddim_steps": T, "schedule": sched,
    "psnr_final_db": psnr(final, ref), "psnr_hole_only_db": psnr(final[hole], ref[hole]),
    "psnr_pred_xstart_by_step": {str(i): psnr(gpu_trace[i]["pred_xstart"], trace[i]["pred_xstart"])
                                 for i in (0, T // 4, T // 2, 3 * T // 4, T - 1)},
    "psnr_blended_db": psnr(final * (1 - keep) + gt * keep, ref * (1 - keep) + gt * keep),
    "seconds_gpu_loop": t_gpu, "seconds_cpu_oracle_loop": t_cpu, "cpu_cores": os.cpu_count(),
}
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"parity_{name}_ddim{T}.json"), "w") as f:
    json.dump(res, f, indent=1)
