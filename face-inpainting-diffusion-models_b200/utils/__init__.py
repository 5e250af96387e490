from .schedules import get_named_beta_schedule, betas_for_alpha_bar, create_gaussian_diffusion  # noqa: F401
