"""Deterministic synthetic weights, images and procedural masks.

There is no network for checkpoints or datasets, so benchmarks and parity tests use
random-init weights of the named architecture.  The reference zero-initialises every
`out_layers.3`, `proj_out`, `out.2` and the stem's channels 3..8 (`nn.py:176,254`,
`unet.py:151,195`), which makes a freshly constructed model output exactly 0; the
generator below therefore draws *every* tensor, from a seed, independent of any module
construction order, so that the oracle and the CUDA path can regenerate identical
weights on different machines without shipping them.
"""
import zlib

import numpy as np
import torch

from ..arch import param_shapes, unet_topology


def synth_state_dict(cfg, seed=0, prefix="base_model.", dtype=torch.float32):
    """state_dict with the reference's keys/shapes for ctor kwargs `cfg`."""
    topo = unet_topology(**cfg)
    sd = {}
    for name, shape in param_shapes(topo):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        leaf = name.rsplit(".", 1)[1]
        is_norm = len(shape) == 1 and leaf == "weight"
        if leaf == "bias":
            is_gn_bias = name.rsplit(".", 1)[0] + ".weight" in sd and sd[name.rsplit(".", 1)[0] + ".weight"].dim() == 1
            v = torch.randn(shape, generator=g) * (0.1 if is_gn_bias else 0.02)
        elif is_norm:
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 1.0
            # residual-branch outputs are kept small so that deep stacks stay O(1)
            if any(s in name for s in ("out_layers.3", "proj_out")):
                gain = 0.5
            v = torch.randn(shape, generator=g) * (gain / np.sqrt(fan_in))
        sd[prefix + name] = v.to(dtype)
    return sd


def synth_batch(batch, size, seed=0, device="cpu"):
    """gt in U[-1,1], binary procedural masks with 5-60 % hole coverage (1 = hole).

    Returns gt [B,3,H,W], mask [B,1,H,W] (1 = inpaint), gt_keep_mask = 1 - mask,
    masked_image = gt * (1 - mask)   (conventions: data/dataset.py:136-142, unet.py:199).
    Always generated on the CPU from `seed`, then moved, so every machine sees the same data.
    """
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    rng = np.random.RandomState(seed + 17)
    mask = np.zeros((batch, 1, size, size), dtype=np.float32)
    for b in range(batch):
        target = rng.uniform(0.05, 0.60)
        while mask[b].mean() < target:
            if rng.rand() < 0.6:       # rectangle
                h, w = rng.randint(size // 16 + 1, size // 3 + 2, size=2)
                y, x = rng.randint(0, size - h + 1), rng.randint(0, size - w + 1)
                mask[b, 0, y:y + h, x:x + w] = 1.0
            else:                      # thick axis-aligned stroke
                th = rng.randint(max(1, size // 64), size // 16 + 2)
                if rng.rand() < 0.5:
                    y = rng.randint(0, size - th + 1)
                    x0, x1 = sorted(rng.randint(0, size, size=2))
                    mask[b, 0, y:y + th, x0:x1 + 1] = 1.0
                else:
                    x = rng.randint(0, size - th + 1)
                    y0, y1 = sorted(rng.randint(0, size, size=2))
                    mask[b, 0, y0:y1 + 1, x:x + th] = 1.0
    mask = torch.from_numpy(mask)
    keep = 1 - mask
    return {"gt": gt.to(device), "mask": mask.to(device), "gt_keep_mask": keep.to(device),
            "masked_image": (gt * keep).to(device)}


def merge_lora(sd, rank=8, alpha=16.0, seed=0, prefix="base_model."):
    """Fold a random rank-`rank` LoRA update into every attention `qkv` / `proj_out` weight:
    W <- W + (alpha/rank) * B @ A.  Keys and shapes are unchanged."""
    out = dict(sd)
    for k, w in sd.items():
        if k.endswith(("qkv.weight", "proj_out.weight")):
            g = torch.Generator().manual_seed((seed * 7919 + zlib.crc32(k.encode())) & 0x7FFFFFFF)
            o, i = w.shape[0], w.shape[1]
            A = torch.randn(rank, i, generator=g) / np.sqrt(i)
            B = torch.randn(o, rank, generator=g) * 0.02
            out[k] = (w.reshape(o, i) + (alpha / rank) * (B @ A)).reshape(w.shape).to(w.dtype)
    return out
