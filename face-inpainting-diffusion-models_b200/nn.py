"""Parameter containers for the UNet layers.

These modules own the learnable tensors under exactly the attribute names of the reference's
layers (`nn.py:136-265`), so that `state_dict()` / `load_state_dict()` are interchangeable with
the reference checkpoints.  They carry NO forward arithmetic: the whole network is executed
by `engine.Plan` through the C-ABI kernels (K1 conv, K2 GroupNorm, K3 attention, K5 timestep path).
"""
import math

import torch
import torch.nn as nn


def _zero_(module):
    # the reference zero-initialises out_layers.3 / proj_out / out.2 (nn.py:39-43,176,254; unet.py:151)
    for p in module.parameters():
        p.detach().zero_()
    return module


def timestep_embedding(timesteps, dim, max_period=10000):
    """Sinusoidal embedding with the reference's signature (nn.py:51-61), evaluated by K5.

    `timesteps` must be a CUDA tensor; returns [N, dim] fp32 (cos half first, then sin).
    """
    from . import _lib as L
    L.require_cuda(timesteps)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half
                      ).to(timesteps.device)
    t = timesteps.float().contiguous()
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=torch.float32)
    L.check(L.lib().fidm_timestep_embedding(L.ptr(t), L.ptr(freqs), L.ptr(out), t.shape[0], dim, L.stream()),
            "timestep_embedding")
    return out


class _Params(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("layers are executed by the fused engine, not module by module")


class ResBlock(_Params):
    """Holds in_layers.{0,2}, emb_layers.1, out_layers.{0,3}, skip_connection (nn.py:149-184)."""

    def __init__(self, spec, emb_channels, scale_shift, dropout=0.0):
        super().__init__()
        cin, cout = spec.cin, spec.cout
        self.in_layers = nn.Sequential(nn.GroupNorm(32, cin), nn.SiLU(), nn.Conv2d(cin, cout, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * cout if scale_shift else cout))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, cout), nn.SiLU(), nn.Dropout(p=dropout),
                                        _zero_(nn.Conv2d(cout, cout, 3, padding=1)))
        if spec.skip == "identity":
            self.skip_connection = nn.Identity()
        elif spec.skip == "conv1x1":
            self.skip_connection = nn.Conv2d(cin, cout, 1)
        else:
            self.skip_connection = nn.Conv2d(cin, cout, 3, padding=1)


class AttentionBlock(_Params):
    """Holds norm, qkv, proj_out (nn.py:251-254)."""

    def __init__(self, spec):
        super().__init__()
        c = spec.channels
        self.num_heads = spec.heads
        self.norm = nn.GroupNorm(32, c)
        self.qkv = nn.Conv1d(c, 3 * c, 1)
        self.proj_out = _zero_(nn.Conv1d(c, c, 1))


class Downsample(_Params):
    def __init__(self, spec):
        super().__init__()
        if spec.use_conv:
            self.op = nn.Conv2d(spec.channels, spec.channels, 3, stride=2, padding=1)
        else:
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)


class Upsample(_Params):
    def __init__(self, spec):
        super().__init__()
        if spec.use_conv:
            self.conv = nn.Conv2d(spec.channels, spec.channels, 3, padding=1)


class TimestepEmbedSequential(nn.Sequential):
    """Container with the reference's name (nn.py:80-89); indexable like the reference's."""
