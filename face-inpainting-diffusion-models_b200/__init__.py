"""B200-native masked-inpainting diffusion sampler: drop-in for the sampling path of
Sayzal28/Face-Inpainting-Diffusion-Models (UNetModel / DiffusionInpaintingModel /
create_model_and_diffusion / GaussianDiffusion.*_sample_loop), executed by hand-written sm_100a
kernels behind the C ABI of include/fidm_b200.h.  No CPU fallback exists."""
from .arch import CONFIGS, param_shapes, unet_topology  # noqa: F401
from .gaussian_diffusion import GaussianDiffusion  # noqa: F401
from .losses import LossType, ModelMeanType, ModelVarType  # noqa: F401
from .nn import timestep_embedding  # noqa: F401
from .train_inpainting import (InpaintingModelFn, create_model_and_diffusion,  # noqa: F401
                               sample_with_advanced_inpainting)
from .unet import DiffusionInpaintingModel, UNetModel  # noqa: F401
from .utils.schedules import create_gaussian_diffusion, get_named_beta_schedule  # noqa: F401
from .script_sampler import InpaintingSampler, create_ddim_timestep_sequence  # noqa: F401
