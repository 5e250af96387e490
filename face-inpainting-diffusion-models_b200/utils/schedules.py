"""Beta tables and the GaussianDiffusion factory (host side, float64).

Drop-in for the reference's `utils/schedules.py:9-106`: same function names, argument
names and defaults.  `timestep_respacing` is accepted and ignored exactly as the
reference does (`utils/schedules.py:79`).
"""
import math

import numpy as np


def _scaled_endpoints(n):
    # Ho et al. endpoints, stretched so that the total noise is independent of n
    # (reference: utils/schedules.py:19-22 and :31-34).
    k = 1000 / n
    return k * 0.0001, k * 0.02


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """beta_i = min(1 - abar((i+1)/n) / abar(i/n), max_beta)  (utils/schedules.py:49-66)."""
    n = num_diffusion_timesteps
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        out[i] = min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta)
    return out


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """Named schedules of the reference (utils/schedules.py:9-46)."""
    n = num_diffusion_timesteps
    if schedule_name == "linear":
        lo, hi = _scaled_endpoints(n)
        return np.linspace(lo, hi, n, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(
            n, lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2)
    if schedule_name == "quadratic":
        lo, hi = _scaled_endpoints(n)
        u = np.linspace(0, 1, n, dtype=np.float64)
        return lo + (hi - lo) * (u ** 2)
    if schedule_name in ("sqrt_linear", "sqrt"):
        return np.sqrt(np.linspace(0.0001, 0.02, n, dtype=np.float64))
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False,
                              noise_schedule="linear", use_kl=False, predict_xstart=False,
                              rescale_timesteps=False, rescale_learned_sigmas=False,
                              timestep_respacing=""):
    """Factory with the reference's keyword surface (utils/schedules.py:69-106)."""
    from ..gaussian_diffusion import GaussianDiffusion
    from ..losses import LossType, ModelMeanType, ModelVarType

    if use_kl:
        loss_type = LossType.RESCALED_KL if rescale_learned_sigmas else LossType.KL
    else:
        loss_type = LossType.RESCALED_MSE if rescale_learned_sigmas else LossType.MSE
    if learn_sigma:
        var_type = ModelVarType.LEARNED_RANGE
    else:
        var_type = ModelVarType.FIXED_SMALL if sigma_small else ModelVarType.FIXED_LARGE
    mean_type = ModelMeanType.START_X if predict_xstart else ModelMeanType.EPSILON
    return GaussianDiffusion(betas=get_named_beta_schedule(noise_schedule, steps),
                             model_mean_type=mean_type, model_var_type=var_type,
                             loss_type=loss_type, rescale_timesteps=rescale_timesteps)
