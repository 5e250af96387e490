"""B200-native masked-inpainting diffusion sampler (drop-in for the reference's sampling path)."""
from .arch import CONFIGS, unet_topology, param_shapes  # noqa: F401
