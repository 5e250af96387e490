"""The evaluation scripts' sampling loops.

The reference's CLIs (`test_inp_ddim_100.py`, `test_inp_ddim_50.py`, `test_ddim_30_cos.py`, `tes_ddpm.py`)
do not call `GaussianDiffusion.ddim_sample_loop`; their `InpaintingSampler` runs its own loops
(`test_inp_ddim_100.py:387-576`) whose semantics differ from the class path:

  * DDIM over a STRIDED subset of a long schedule (`create_ddim_timestep_sequence`, :387-400 -- 101 evals
    for "100 steps"), x0 = (x - sqrt(1-ab_t) eps) / sqrt(ab_t), update with the RAW eps (:539-557);
  * the known region is injected AFTER the update, at the noise level of the next timestep, with FRESH
    noise every step (:560-574), and x_T itself is not injected;
  * the caller blends `result * mask + gt * (1 - mask)` at the end (:693-696).

Here each step is the model call plus ONE fused K4 launch (`FIDM_SAMPLER_DDIM_SCRIPT`, or the DDPM update,
followed by the injection), with a per-step coefficient table computed in the scripts' own arithmetic
(float64 scalars rounded to fp32 when they meet the fp32 tensors).
"""
import numpy as np
import torch

from . import _lib as L
from .train_inpainting import InpaintingModelFn


def create_ddim_timestep_sequence(total_timesteps, ddim_timesteps):
    """Evenly strided timesteps plus the last one, high to low (test_inp_ddim_100.py:387-400)."""
    stride = total_timesteps // ddim_timesteps
    seq = np.arange(0, total_timesteps, stride)
    if seq[-1] != total_timesteps - 1:
        seq = np.append(seq, total_timesteps - 1)
    return seq[::-1]


def script_ddim_table(diffusion, seq, eta):
    """[len(seq), FIDM_COEF_COLS] fp32 rows for FIDM_SAMPLER_DDIM_SCRIPT (float64 math, then fp32)."""
    ac = diffusion.alphas_cumprod
    a_t = torch.tensor([ac[int(t)] for t in seq], dtype=torch.float64)
    a_p = torch.tensor([ac[int(seq[k + 1])] if k + 1 < len(seq) else 1.0 for k in range(len(seq))],
                       dtype=torch.float64)
    sigma = eta * torch.sqrt((1 - a_p) / (1 - a_t)) * torch.sqrt(1 - a_t / a_p)
    tab = torch.zeros(len(seq), L.COEF_COLS, dtype=torch.float32)
    tab[:, 0] = torch.sqrt(a_p).float()                       # injection at the NEXT level (:562-570)
    tab[:, 1] = torch.sqrt(1 - a_p).float()
    tab[:, 4] = torch.sqrt(a_t).float()                       # x0 = (x - c5 eps) / c4 (:536-539)
    tab[:, 5] = torch.sqrt(1 - a_t).float()
    tab[:, 11] = torch.sqrt(a_p).float()
    tab[:, 12] = torch.sqrt(1 - a_p - sigma ** 2).float()
    tab[:, 13] = sigma.float()
    tab[:, 14] = 1.0
    return tab


class InpaintingSampler:
    """The sampling methods of the scripts' `InpaintingSampler` (test_inp_ddim_100.py:288-698); dataset,
    metric and PNG plumbing of that class are out of scope."""

    def __init__(self, model, diffusion, ddim_timesteps=100, use_ddim=True, eta=0.0, clip_denoised=True,
                 blend_output=True):
        self.model, self.diffusion = model, diffusion
        self.ddim_timesteps, self.use_ddim, self.eta = ddim_timesteps, use_ddim, eta
        self.clip_denoised, self.blend_output = clip_denoised, blend_output
        self._fn = InpaintingModelFn(model)

    def model_fn(self, x, t, gt=None, gt_keep_mask=None, **kwargs):
        """:373-385."""
        return self._fn(x, t, gt=gt, gt_keep_mask=gt_keep_mask, **kwargs)

    create_ddim_timestep_sequence = staticmethod(create_ddim_timestep_sequence)

    def inpainting_ddim_sample_loop(self, model_fn, shape, gt_images, masks, clip_denoised=True, device=None,
                                    progress=False, eta=0.0):
        """:470-576.  masks: 1 = inpaint."""
        d = self.diffusion
        if device is None:
            device = next(self.model.parameters()).device
        img = torch.randn(*shape, device=device)
        L.require_cuda(img)
        seq = create_ddim_timestep_sequence(d.num_timesteps, self.ddim_timesteps)
        table = script_ddim_table(d, seq, eta).to(device)
        keep = 1 - masks
        kw = {"gt": gt_images, "gt_keep_mask": keep}
        B = shape[0]
        with torch.no_grad():
            for k, timestep in enumerate(seq):
                timestep = int(timestep)
                t = torch.full((B,), timestep, device=device, dtype=torch.long)
                out = model_fn(img, t, **kw)
                if out.shape[1] not in (3, 6):
                    raise ValueError(f"Unexpected model output shape: {out.shape}")
                z = torch.randn_like(img) if (timestep > 0 and eta > 0) else None
                inj = timestep > 0
                r = d._step(L.STEP_UPDATE_INJECT if inj else L.STEP_UPDATE_ONLY, img, t=k, t_inject=k, model_out=out,
                            z=z, gt=gt_images if inj else None, keep=keep if inj else None,
                            inject_noise=torch.randn_like(gt_images) if inj else None, clip=clip_denoised,
                            cumulative=True, script_table=table, want_next=inj, want_sample=not inj)
                img = r["x_next"] if inj else r["sample"]
        return img

    def inpainting_p_sample_loop(self, model_fn, shape, gt_images, masks, clip_denoised=True, device=None,
                                 progress=False):
        """:402-468: DDPM update at i, then injection at level i-1 with fresh noise."""
        d = self.diffusion
        if device is None:
            device = next(self.model.parameters()).device
        img = torch.randn(*shape, device=device)
        L.require_cuda(img)
        keep = 1 - masks
        kw = {"gt": gt_images, "gt_keep_mask": keep}
        B = shape[0]
        with torch.no_grad():
            for i in range(d.num_timesteps - 1, -1, -1):
                t = torch.full((B,), i, device=device, dtype=torch.long)
                out = model_fn(img, t, **kw)
                z = torch.randn_like(img)
                inj = i > 0
                r = d._step(L.STEP_UPDATE_INJECT if inj else L.STEP_UPDATE_ONLY, img, t=i, t_inject=i - 1, model_out=out,
                            z=z, gt=gt_images if inj else None, keep=keep if inj else None,
                            inject_noise=torch.randn_like(gt_images) if inj else None, ddim=False, clip=clip_denoised,
                            cumulative=True, want_next=inj, want_sample=not inj)
                img = r["x_next"] if inj else r["sample"]
        return img

    def sample_batch(self, gt_images, masks):
        """The sampling core of :578-698 for tensors already on the device: returns
        (result, gt_images, masked_images, masks) with the final blend applied when `blend_output`."""
        masked = gt_images * (1 - masks)
        shape = tuple(gt_images.shape)
        if self.use_ddim:
            res = self.inpainting_ddim_sample_loop(self.model_fn, shape, gt_images, masks, self.clip_denoised,
                                                   gt_images.device, eta=self.eta)
        else:
            res = self.inpainting_p_sample_loop(self.model_fn, shape, gt_images, masks, self.clip_denoised,
                                                gt_images.device)
        if self.blend_output:
            res = res * masks + gt_images * (1 - masks)      # :693-696
        return res, gt_images, masked, masks
