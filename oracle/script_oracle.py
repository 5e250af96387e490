"""CPU ORACLE (test infrastructure only) -- the evaluation scripts' sampling loops.

Restates `InpaintingSampler.inpainting_ddim_sample_loop` / `inpainting_p_sample_loop` /
`create_ddim_timestep_sequence` of /root/reference/code/test_inp_ddim_100.py:387-576 with explicit noise.
Pinned by oracle/make_golden.py, which executes the reference's own method bodies (extracted from the
script's source; the script itself cannot be imported: lpips / skimage / pytorch_fid are not installed).
"""
import numpy as np
import torch

from .diffusion_oracle import ddpm_update


def ddim_timestep_sequence(total, n):
    """:387-400."""
    c = total // n
    seq = np.asarray(list(range(0, total, c)))
    if seq[-1] != total - 1:
        seq = np.append(seq, total - 1)
    return seq[::-1]


def script_ddim_loop(tab, model_fn, shape, gt, masks, n_steps, *, eta=0.0, clip=True, noise_fn=None, x_T=None):
    """:470-576.  masks: 1 = inpaint.  noise_fn(kind, t): kind in {"step", "inject"}."""
    img = x_T if x_T is not None else torch.randn(*shape)
    seq = ddim_timestep_sequence(tab.T, n_steps)
    keep = 1 - masks
    for k, timestep in enumerate(seq):
        t = torch.tensor([timestep] * shape[0])
        out = model_fn(img, t, gt=gt, gt_keep_mask=keep)
        eps = out[:, :3] if out.shape[1] == 6 else out
        a_t = torch.tensor(tab.alphas_cumprod[timestep])
        a_p = torch.tensor(tab.alphas_cumprod[seq[k + 1]] if k < len(seq) - 1 else 1.0)
        x0 = (img - torch.sqrt(1 - a_t) * eps) / torch.sqrt(a_t)
        if clip:
            x0 = torch.clamp(x0, -1, 1)
        sigma = eta * torch.sqrt((1 - a_p) / (1 - a_t)) * torch.sqrt(1 - a_t / a_p)
        direction = torch.sqrt(1 - a_p - sigma ** 2) * eps
        z = (noise_fn("step", timestep) if noise_fn else torch.randn_like(img)) if (timestep > 0 and eta > 0) \
            else torch.zeros_like(img)
        img = torch.sqrt(a_p) * x0 + direction + sigma * z
        if timestep > 0:
            n = noise_fn("inject", timestep) if noise_fn else torch.randn_like(gt)
            noised = torch.sqrt(a_p) * gt + torch.sqrt(1 - a_p) * n
            img = img * masks + noised * keep
    return img


def script_ddpm_loop(tab, model_fn, shape, gt, masks, *, clip=True, var_type="learned_range", noise_fn=None, x_T=None):
    """:402-468."""
    img = x_T if x_T is not None else torch.randn(*shape)
    keep = 1 - masks
    for i in range(tab.T - 1, -1, -1):
        t = torch.tensor([i] * shape[0])
        out = model_fn(img, t, gt=gt, gt_keep_mask=keep)
        z = noise_fn("step", i) if noise_fn else torch.randn_like(img)
        img, _ = ddpm_update(tab, out, img, i, z, var_type, clip)
        if i > 0:
            a = torch.tensor(tab.alphas_cumprod[i - 1])
            n = noise_fn("inject", i) if noise_fn else torch.randn_like(gt)
            noised = torch.sqrt(a) * gt + torch.sqrt(1 - a) * n
            img = img * masks + noised * keep
    return img
