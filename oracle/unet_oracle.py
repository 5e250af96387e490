"""CPU ORACLE (test infrastructure only) -- 9-channel inpainting UNet, restated functionally.

This file is a checker.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it; the product path
(`face-inpainting-diffusion-models_b200/`) never does and has no CPU fallback.

It restates, in plain fp32 PyTorch-on-CPU calls driven directly by a *state_dict*, what
the reference builds out of nn.Modules:

  * topology walk ............ /root/reference/code/unet.py:41-152 (ctor) and :154-173 (forward)
  * 9-channel wrapper ........ unet.py:197-200  (cat [x, masked_image, mask x3])
  * timestep_embedding ....... nn.py:51-61      (cos first, then sin)
  * ResBlock ................. nn.py:189-212    (GN-SiLU-[resample]-conv, emb scale/shift, skip)
  * Up/Downsample ............ nn.py:92-133
  * AttentionBlock/QKV ....... nn.py:222-235, :259-265

Parity status: pinned against the *reference itself* run in the build container
(`oracle/make_golden.py` imports /root/reference/code and stores its outputs under
tests/golden/); the reference ships no tests or golden vectors of its own.
"""
import math

import torch
import torch.nn.functional as F


def sinusoid(t, dim, max_period=10000):
    """nn.py:51-61."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    ang = t[:, None].float() * freqs[None].to(t.device)
    out = torch.cat([ang.cos(), ang.sin()], dim=-1)
    if dim % 2:
        out = torch.cat([out, torch.zeros_like(out[:, :1])], dim=-1)
    return out


def _gn(sd, key, x):
    return F.group_norm(x, 32, sd[key + ".weight"], sd[key + ".bias"], eps=1e-5)


def _conv(sd, key, x, stride=1):
    w = sd[key + ".weight"]
    return F.conv2d(x, w, sd[key + ".bias"], stride=stride, padding=w.shape[-1] // 2)


def resblock(sd, p, x, emb, *, up=False, down=False, scale_shift=True):
    """nn.py:189-212."""
    h = F.silu(_gn(sd, p + ".in_layers.0", x))
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif down:
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    h = _conv(sd, p + ".in_layers.2", h)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    e = e[:, :, None, None]
    if scale_shift:
        scale, shift = e.chunk(2, dim=1)
        h = _gn(sd, p + ".out_layers.0", h) * (1 + scale) + shift
        h = _conv(sd, p + ".out_layers.3", F.silu(h))
    else:
        h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h + e)))
    if p + ".skip_connection.weight" in sd:
        x = _conv(sd, p + ".skip_connection", x)
    return x + h


def attention(sd, p, x, heads):
    """nn.py:222-235 and :259-265."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, p + ".norm", xf), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    ch = c // heads
    q, k, v = qkv.chunk(3, dim=1)
    s = 1.0 / math.sqrt(math.sqrt(ch))
    t = xf.shape[-1]
    w = torch.einsum("bct,bcs->bts", (q * s).reshape(b * heads, ch, t), (k * s).reshape(b * heads, ch, t))
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(b * heads, ch, t)).reshape(b, c, t)
    a = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + a).reshape(b, c, hh, ww)


def _heads(cfg, c, upsample=False):
    if cfg.get("num_head_channels", -1) != -1:
        return c // cfg["num_head_channels"]
    if upsample and cfg.get("num_heads_upsample", -1) != -1:
        return cfg["num_heads_upsample"]
    return cfg.get("num_heads", 1)


def unet_forward(sd, cfg, x, t, prefix="base_model."):
    """UNetModel.forward (unet.py:154-173) evaluated from a state_dict.

    cfg: the ctor kwargs of unet.py:17-21 (dict).  x: [B,in_ch,H,W] fp32.  t: [B].
    """
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    mc = cfg["model_channels"]
    mult = tuple(cfg.get("channel_mult", (1, 2, 4, 8)))
    nrb = cfg["num_res_blocks"]
    attn_ds = set(cfg["attention_resolutions"])
    ssn = cfg.get("use_scale_shift_norm", False)
    updown = cfg.get("resblock_updown", False)
    conv_resample = cfg.get("conv_resample", True)

    emb = F.linear(sinusoid(t, mc), sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    hs = []
    h = _conv(sd, "input_blocks.0.0", x)
    hs.append(h)
    idx, ds, ch = 1, 1, int(mult[0] * mc)
    for level, m in enumerate(mult):
        for _ in range(nrb):
            h = resblock(sd, f"input_blocks.{idx}.0", h, emb, scale_shift=ssn)
            ch = int(m * mc)
            if ds in attn_ds:
                h = attention(sd, f"input_blocks.{idx}.1", h, _heads(cfg, ch))
            hs.append(h)
            idx += 1
        if level != len(mult) - 1:
            if updown:
                h = resblock(sd, f"input_blocks.{idx}.0", h, emb, down=True, scale_shift=ssn)
            elif conv_resample:
                h = _conv(sd, f"input_blocks.{idx}.0.op", h, stride=2)
            else:
                h = F.avg_pool2d(h, 2)
            hs.append(h)
            idx += 1
            ds *= 2

    h = resblock(sd, "middle_block.0", h, emb, scale_shift=ssn)
    h = attention(sd, "middle_block.1", h, _heads(cfg, ch))
    h = resblock(sd, "middle_block.2", h, emb, scale_shift=ssn)

    idx = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            h = torch.cat([h, hs.pop()], dim=1)
            h = resblock(sd, f"output_blocks.{idx}.0", h, emb, scale_shift=ssn)
            ch = int(mc * m)
            j = 1
            if ds in attn_ds:
                h = attention(sd, f"output_blocks.{idx}.{j}", h, _heads(cfg, ch, upsample=True))
                j += 1
            if level and i == nrb:
                if updown:
                    h = resblock(sd, f"output_blocks.{idx}.{j}", h, emb, up=True, scale_shift=ssn)
                else:
                    h = F.interpolate(h, scale_factor=2, mode="nearest")
                    if conv_resample:
                        h = _conv(sd, f"output_blocks.{idx}.{j}.conv", h)
                ds //= 2
            idx += 1

    return _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h)))


def inpaint_forward(sd, cfg, x, t, masked_image, mask):
    """DiffusionInpaintingModel.forward (unet.py:197-200)."""
    return unet_forward(sd, cfg, torch.cat([x, masked_image, mask.repeat(1, 3, 1, 1)], dim=1), t)
