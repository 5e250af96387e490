"""CPU ORACLE (test infrastructure only) -- masked-inpainting DDIM / DDPM reverse process.

Checker only: imported by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs,
never by the product path.

Restates, as flat per-step arithmetic on fp32 CPU tensors (one torch op per rounding, in the
reference's op order -- SURVEY.md Appendix A), these pieces of
/root/reference/code/gaussian_diffusion.py:

  * float64 tables ................. :41-83
  * coefficient lookup ............. :12-24   (float64 entry rounded to fp32)
  * known-region injection ......... :85-157, q_sample :172-189
  * p_mean_variance ................ :213-298 (+ :191-211, :300-319)
  * p_sample / ddim_sample ......... :357-388 / :447-485
  * *_sample_loop .................. :390-445 / :487-538

Unlike the reference, every noise tensor is an explicit argument (`noise_fn(kind, t)`),
so the same numbers can be fed to the CUDA path.  `noise_fn=None` draws from torch's
global generator in the reference's order (randn(shape); per step randn_like(gt) on
cache miss, then randn_like(x)).

Parity status: pinned against the reference itself by `oracle/make_golden.py` / `make_golden_r2.py`
(bit-exact on the DDIM path for every mean / variance type, tests/test_oracle_pinned.py); the reference has no tests.
"""
import numpy as np
import torch


class Tables:
    """gaussian_diffusion.py:41-83, same names."""

    def __init__(self, betas):
        b = np.array(betas, dtype=np.float64)
        assert b.ndim == 1 and (b > 0).all() and (b <= 1).all()
        self.betas = b
        self.T = int(b.shape[0])
        a = 1.0 - b
        self.alphas_cumprod = np.cumprod(a, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = b * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = b * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(a) / (1.0 - self.alphas_cumprod)


def _c(arr, t):
    """:12-24 for a scalar timestep: float64 table entry -> fp32 0-dim tensor."""
    return torch.tensor(float(arr[t]), dtype=torch.float64).float()


def inject(tab, x, t, gt, keep, noise, cumulative=True):
    """apply_inpainting_injection (:114-157) with the noise passed in."""
    if cumulative:
        wg = _c(tab.sqrt_alphas_cumprod, t) * gt + _c(tab.sqrt_one_minus_alphas_cumprod, t) * noise
    else:
        ab = _c(tab.alphas_cumprod, t)
        wg = torch.sqrt(ab) * gt + torch.sqrt(1 - ab) * noise
    m = keep.repeat(1, x.shape[1], 1, 1) if keep.shape[1] == 1 and x.shape[1] > 1 else keep
    return m * wg + (1 - m) * x


def mean_variance(tab, out, x, t, var_type="learned_range", clip=True, mean_type="epsilon"):
    """p_mean_variance (:213-298); returns mean, log_variance, x0.  mean_type: "epsilon" | "start_x" | "previous_x"
    (ModelMeanType, :273-286); var_type: "learned_range" | "learned" | "fixed_large" | "fixed_small" (:241-265)."""
    C = x.shape[1]
    if var_type == "learned_range":
        mo, v = torch.split(out, C, dim=1)
        min_log = _c(tab.posterior_log_variance_clipped, t)
        max_log = _c(np.log(tab.betas), t)
        frac = (v + 1) / 2
        logvar = frac * max_log + (1 - frac) * min_log
    elif var_type == "learned":
        mo, logvar = torch.split(out, C, dim=1)
    elif var_type == "fixed_large":
        mo = out
        logvar = _c(np.log(np.append(tab.posterior_variance[1], tab.betas[1:])), t).expand(x.shape)
    elif var_type == "fixed_small":
        mo = out
        logvar = _c(tab.posterior_log_variance_clipped, t).expand(x.shape)
    else:
        raise NotImplementedError(var_type)

    def process(v):
        return v.clamp(-1, 1) if clip else v

    if mean_type == "previous_x":                      # :274-278, _predict_xstart_from_xprev :307-314
        x0 = process(_c(1.0 / tab.posterior_mean_coef1, t) * mo
                     - _c(tab.posterior_mean_coef2 / tab.posterior_mean_coef1, t) * x)
        return mo, logvar, x0
    if mean_type == "start_x":
        x0 = process(mo)
    elif mean_type == "epsilon":
        x0 = process(_c(tab.sqrt_recip_alphas_cumprod, t) * x - _c(tab.sqrt_recipm1_alphas_cumprod, t) * mo)
    else:
        raise NotImplementedError(mean_type)
    mean = _c(tab.posterior_mean_coef1, t) * x0 + _c(tab.posterior_mean_coef2, t) * x
    return mean, logvar, x0


def ddim_update(tab, out, x, t, z, eta=0.0, var_type="learned_range", clip=True, mean_type="epsilon"):
    """ddim_sample (:464-485) after the injection."""
    _, _, x0 = mean_variance(tab, out, x, t, var_type, clip, mean_type)
    eps = (_c(tab.sqrt_recip_alphas_cumprod, t) * x - x0) / _c(tab.sqrt_recipm1_alphas_cumprod, t)
    ab = _c(tab.alphas_cumprod, t)
    abp = _c(tab.alphas_cumprod_prev, t)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    mean_pred = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
    nz = torch.tensor(float(t != 0))
    return mean_pred + nz * sigma * z, x0


def ddpm_update(tab, out, x, t, z, var_type="learned_range", clip=True, mean_type="epsilon"):
    """p_sample (:378-388) after the injection."""
    mean, logvar, x0 = mean_variance(tab, out, x, t, var_type, clip, mean_type)
    nz = torch.tensor(float(t != 0))
    return mean + nz * torch.exp(0.5 * logvar) * z, x0


def sample_loop(tab, model, shape, *, ddim=True, eta=0.0, x_T=None, gt=None, keep=None,
                model_kwargs=None, var_type="learned_range", clip=True, inject_on=True,
                schedule="all", cumulative=True, noise_fn=None, trace=None, mean_type="epsilon", rescale_timesteps=False):
    """ddim_sample_loop / p_sample_loop (:390-445, :487-538).

    model(x, t_int64[B], **model_kwargs) -> [B, C or 2C, H, W].
    noise_fn(kind, t) with kind in {"inject", "step"} supplies noise; None -> torch.randn.
    """
    model_kwargs = dict(model_kwargs or {})
    if gt is not None:
        model_kwargs.setdefault("gt", gt)
        model_kwargs.setdefault("gt_keep_mask", keep)
    x = x_T if x_T is not None else torch.randn(*shape)
    cache = {}
    x0 = None
    for t in range(tab.T - 1, -1, -1):
        if inject_on and gt is not None:
            gated = (schedule == "high" and t < tab.T // 2) or (schedule == "low" and t >= tab.T // 2)
            if not gated:
                if cumulative:
                    if t not in cache:
                        cache[t] = noise_fn("inject", t) if noise_fn else torch.randn_like(gt)
                    n = cache[t]
                else:
                    n = noise_fn("inject", t) if noise_fn else torch.randn_like(gt)
                x = inject(tab, x, t, gt, keep, n, cumulative)
        tt = torch.full((shape[0],), t, dtype=torch.int64)
        if rescale_timesteps:                              # _scale_timesteps (:321-324)
            tt = tt.float() * (1000.0 / tab.T)
        with torch.no_grad():
            out = model(x, tt, **model_kwargs)
        z = noise_fn("step", t) if noise_fn else torch.randn_like(x)
        if ddim:
            x, x0 = ddim_update(tab, out, x, t, z, eta, var_type, clip, mean_type)
        else:
            x, x0 = ddpm_update(tab, out, x, t, z, var_type, clip, mean_type)
        if trace is not None:
            trace.append({"t": t, "sample": x.clone(), "pred_xstart": x0.clone(), "model_out": out.clone()})
    return x
