"""Topology of the reference UNet as plain data.

`unet_topology(**ctor_kwargs)` turns the constructor arguments of the reference's
`UNetModel` (`unet.py:17-21`) into a list of layer descriptors whose names are the
reference's state_dict prefixes (`unet.py:52-152`).  Both the parameter container
(`unet.py` of this package) and the execution plan (`engine.py`) are derived from it, so
the checkpoint layout and the kernel schedule cannot drift apart.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple


@dataclass
class Res:
    name: str
    cin: int
    cout: int
    up: bool = False
    down: bool = False
    skip: str = "identity"            # "identity" | "conv1x1" | "conv3x3"
    kind: str = "res"


@dataclass
class Attn:
    name: str
    channels: int
    heads: int
    kind: str = "attn"


@dataclass
class Down:                           # nn.py:115-133 (only when resblock_updown=False)
    name: str
    channels: int
    use_conv: bool
    kind: str = "down"


@dataclass
class Up:                             # nn.py:92-112
    name: str
    channels: int
    use_conv: bool
    kind: str = "up"


@dataclass
class Stem:
    name: str
    cin: int
    cout: int
    kind: str = "stem"


@dataclass
class Block:
    """One `TimestepEmbedSequential`; `skip_ch` is the width of the `hs.pop()` tensor that is
    concatenated in front of an output block (`unet.py:170`)."""
    layers: List[object]
    cin: int
    cout: int
    ds: int                           # downsample factor of the block's OUTPUT
    skip_ch: int = 0


@dataclass
class Topology:
    cfg: dict
    input_blocks: List[Block] = field(default_factory=list)
    middle: Optional[Block] = None
    output_blocks: List[Block] = field(default_factory=list)
    head_ch: int = 0
    out_channels: int = 0
    time_embed_dim: int = 0


UNET_DEFAULTS = dict(dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                     num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=1,
                     num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                     resblock_updown=False, use_new_attention_order=False)


def _n_heads(ch, num_heads, num_head_channels):
    # nn.py:245-249
    if num_head_channels == -1:
        return num_heads
    assert ch % num_head_channels == 0
    return ch // num_head_channels


def unet_topology(image_size, in_channels, model_channels, out_channels, num_res_blocks,
                  attention_resolutions, **kw) -> Topology:
    cfg = dict(UNET_DEFAULTS)
    unknown = set(kw) - set(cfg)
    if unknown:
        raise TypeError(f"unexpected UNetModel arguments: {sorted(unknown)}")
    cfg.update(kw)
    cfg.update(image_size=image_size, in_channels=in_channels, model_channels=model_channels,
               out_channels=out_channels, num_res_blocks=num_res_blocks,
               attention_resolutions=tuple(attention_resolutions))
    if cfg["dims"] != 2:
        raise NotImplementedError("only dims=2 is on the sampling path (SURVEY.md 8-b)")
    if cfg["num_classes"] is not None:
        raise NotImplementedError("class conditioning is not used by the inpainting sampler")
    heads_up = cfg["num_heads"] if cfg["num_heads_upsample"] == -1 else cfg["num_heads_upsample"]
    mult = tuple(cfg["channel_mult"])
    mc, nrb = model_channels, num_res_blocks
    updown, conv_rs = cfg["resblock_updown"], cfg["conv_resample"]
    attn_at = set(cfg["attention_resolutions"])

    def res(name, cin, cout, **f):
        return Res(name, cin, cout, skip="identity" if cin == cout else "conv1x1", **f)

    topo = Topology(cfg=cfg, out_channels=out_channels, time_embed_dim=4 * mc)
    ch = int(mult[0] * mc)
    topo.head_ch = ch
    topo.input_blocks.append(Block([Stem("input_blocks.0.0", in_channels, ch)], in_channels, ch, 1))
    widths = [ch]
    ds = 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            i = len(topo.input_blocks)
            cout = int(m * mc)
            layers = [res(f"input_blocks.{i}.0", ch, cout)]
            if ds in attn_at:
                layers.append(Attn(f"input_blocks.{i}.1", cout,
                                   _n_heads(cout, cfg["num_heads"], cfg["num_head_channels"])))
            topo.input_blocks.append(Block(layers, ch, cout, ds))
            ch = cout
            widths.append(ch)
        if level != len(mult) - 1:
            i = len(topo.input_blocks)
            layer = (res(f"input_blocks.{i}.0", ch, ch, down=True) if updown
                     else Down(f"input_blocks.{i}.0", ch, conv_rs))
            ds *= 2
            topo.input_blocks.append(Block([layer], ch, ch, ds))
            widths.append(ch)

    topo.middle = Block([
        res("middle_block.0", ch, ch),
        Attn("middle_block.1", ch, _n_heads(ch, cfg["num_heads"], cfg["num_head_channels"])),
        res("middle_block.2", ch, ch)], ch, ch, ds)

    for level, m in list(enumerate(mult))[::-1]:
        for k in range(nrb + 1):
            i = len(topo.output_blocks)
            sk = widths.pop()
            cout = int(mc * m)
            layers = [res(f"output_blocks.{i}.0", ch + sk, cout)]
            j = 1
            if ds in attn_at:
                layers.append(Attn(f"output_blocks.{i}.{j}", cout,
                                   _n_heads(cout, heads_up, cfg["num_head_channels"])))
                j += 1
            blk_ds = ds
            if level and k == nrb:
                layers.append(res(f"output_blocks.{i}.{j}", cout, cout, up=True) if updown
                              else Up(f"output_blocks.{i}.{j}", cout, conv_rs))
                ds //= 2
                blk_ds = ds
            topo.output_blocks.append(Block(layers, ch + sk, cout, blk_ds, skip_ch=sk))
            ch = cout
    assert not widths
    return topo


def all_layers(topo: Topology):
    for blk in topo.input_blocks + [topo.middle] + topo.output_blocks:
        for layer in blk.layers:
            yield layer


def param_shapes(topo: Topology) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every tensor of the reference's `UNetModel.state_dict()`, in the
    reference's registration order (`unet.py:44-152`, `nn.py:149-184`, `nn.py:251-254`)."""
    cfg = topo.cfg
    mc, ted = cfg["model_channels"], topo.time_embed_dim
    ssn = cfg["use_scale_shift_norm"]
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def wb(name, *wshape):
        out.append((name + ".weight", tuple(wshape)))
        out.append((name + ".bias", (wshape[0],)))

    wb("time_embed.0", ted, mc)
    wb("time_embed.2", ted, ted)

    def emit(layer):
        n = layer.name
        if layer.kind == "stem":
            wb(n, layer.cout, layer.cin, 3, 3)
        elif layer.kind == "res":
            wb(n + ".in_layers.0", layer.cin)
            wb(n + ".in_layers.2", layer.cout, layer.cin, 3, 3)
            wb(n + ".emb_layers.1", (2 if ssn else 1) * layer.cout, ted)
            wb(n + ".out_layers.0", layer.cout)
            wb(n + ".out_layers.3", layer.cout, layer.cout, 3, 3)
            if layer.skip == "conv1x1":
                wb(n + ".skip_connection", layer.cout, layer.cin, 1, 1)
            elif layer.skip == "conv3x3":
                wb(n + ".skip_connection", layer.cout, layer.cin, 3, 3)
        elif layer.kind == "attn":
            wb(n + ".norm", layer.channels)
            wb(n + ".qkv", 3 * layer.channels, layer.channels, 1)
            wb(n + ".proj_out", layer.channels, layer.channels, 1)
        elif layer.kind == "down" and layer.use_conv:
            wb(n + ".op", layer.channels, layer.channels, 3, 3)
        elif layer.kind == "up" and layer.use_conv:
            wb(n + ".conv", layer.channels, layer.channels, 3, 3)

    for layer in all_layers(topo):
        emit(layer)
    wb("out.0", topo.head_ch)
    wb("out.2", topo.out_channels, topo.head_ch, 3, 3)
    return out


# Named configurations of SURVEY.md section 8 (ctor kwargs of unet.py:17-21, with in_channels=9).
CONFIGS = {
    "T64": dict(image_size=64, in_channels=9, model_channels=64, out_channels=6, num_res_blocks=1,
                attention_resolutions=(4,), channel_mult=(1, 2, 4, 8), num_heads=4,
                use_scale_shift_norm=True, resblock_updown=True),
    # exactly train_inpainting.py:208-224 after the 9-channel stem swap (unet.py:184-195)
    "REF_FFHQ256": dict(image_size=256, in_channels=9, model_channels=128, out_channels=6,
                        num_res_blocks=1, attention_resolutions=(16,),
                        channel_mult=(1, 1, 2, 2, 4, 4), num_heads=4, num_head_channels=64,
                        use_scale_shift_norm=True, resblock_updown=True),
    "ADM256": dict(image_size=256, in_channels=9, model_channels=256, out_channels=6,
                   num_res_blocks=2, attention_resolutions=(8, 16, 32),
                   channel_mult=(1, 1, 2, 2, 4, 4), num_heads=4, num_head_channels=64,
                   use_scale_shift_norm=True, resblock_updown=True),
}
