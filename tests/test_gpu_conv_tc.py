"""GPU: K1 tcgen05 implicit-GEMM convolution against torch fp32 on the same bf16-rounded operands."""
import math

import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu


def _ref(x, w, b, row_add=None, residual=None, x2=None, w2=None):
    y = Fn.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=w.shape[-1] // 2)
    if x2 is not None:
        y = y + Fn.conv2d(x2.float().permute(0, 3, 1, 2), w2.float())
    if row_add is not None:
        y = y + row_add[:, :, None, None]
    if residual is not None:
        y = y + residual.float().permute(0, 3, 1, 2)
    return y


def _mk(B, H, W, Cin, Cout, ks, seed):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, H, W, Cin, device=dev, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, ks, ks, device=dev, generator=g) / math.sqrt(Cin * ks * ks)).bfloat16()
    b = torch.randn(Cout, device=dev, generator=g)
    return x, w, b, g


def _check(y_nhwc, want, what):
    got = y_nhwc.float().permute(0, 3, 1, 2)
    err = (got - want).abs().max().item()
    rel = ((got - want).norm() / want.norm()).item()
    assert rel < 4e-3 and err < 6e-2, (what, rel, err)     # output is rounded to bf16 (2^-9 relative)


# every pixel-box geometry (W = 8 .. 256), every N tile (256 / 128 / 64), 3x3 and 1x1
SHAPES = [
    (2, 8, 8, 64, 64, 3), (3, 8, 8, 128, 256, 3), (1, 8, 8, 64, 64, 1),        # TN = 2 tiles, odd batch
    (2, 16, 16, 64, 128, 3), (1, 16, 16, 192, 64, 3), (2, 32, 32, 128, 128, 3),
    (1, 64, 64, 64, 256, 3), (1, 64, 64, 256, 512, 1), (1, 128, 128, 64, 64, 3),
    (1, 256, 256, 64, 256, 3), (2, 32, 32, 512, 1536, 1), (4, 8, 8, 1024, 1024, 3),
    (1, 16, 16, 2048, 1024, 3), (8, 64, 64, 128, 256, 3),
    (37, 16, 24, 64, 256, 3),      # CTA-pair kernel with an ODD number of M tiles (phantom half-tile), W = 24
    (2, 128, 128, 256, 512, 3),    # CTA-pair kernel, two N blocks
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks", SHAPES)
def test_conv_tc_plain(cuda_lib, B, H, W, Cin, Cout, ks):
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, _ = _mk(B, H, W, Cin, Cout, ks, seed=H * 7 + Cin)
    y = ops.conv2d(x, ops.repack_weight(w.float()), b, impl="tc")
    torch.cuda.synchronize()
    _check(y, _ref(x, w, b), (B, H, W, Cin, Cout, ks))


@pytest.mark.parametrize("B,H,W,Cin,Cout,Cin2", [(2, 16, 16, 128, 64, 64), (1, 64, 64, 64, 256, 128),
                                                 (3, 8, 8, 256, 128, 192), (1, 256, 256, 64, 256, 64)])
def test_conv_tc_fused_epilogue_and_skip(cuda_lib, B, H, W, Cin, Cout, Cin2):
    """bias + timestep row-add + residual, the 1x1 skip as extra K slices, and zero-copy writes into /
    reads from channel slices of wider (concat) buffers."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, g = _mk(B, H, W, Cin, Cout, 3, seed=Cin2)
    dev = "cuda"
    wk = ops.repack_weight(w.float())
    emb = torch.randn(B, 3 * Cout, device=dev, generator=g)
    row = emb[:, Cout:2 * Cout]
    res_buf = torch.randn(B, H, W, Cout + 64, device=dev, generator=g).bfloat16()
    res = res_buf[..., 64:]
    y = ops.conv2d(x, wk, b, row_add=row, residual=res, impl="tc")
    _check(y, _ref(x, w, b, row, res), "epilogue")
    x2_buf = torch.randn(B, H, W, Cin2 + 128, device=dev, generator=g).bfloat16()
    x2 = x2_buf[..., 64:64 + Cin2]
    w2 = (torch.randn(Cout, Cin2, 1, 1, device=dev, generator=g) / math.sqrt(Cin2)).bfloat16()
    out_buf = torch.zeros(B, H, W, Cout + 128, device=dev, dtype=torch.bfloat16)
    out = out_buf[..., 64:64 + Cout]
    ops.conv2d(x, wk, b, x2=x2, w2=ops.repack_weight(w2.float()), out=out, impl="tc")
    _check(out, _ref(x, w, b, x2=x2, w2=w2), "skip")
    assert not out_buf[..., :64].any() and not out_buf[..., 64 + Cout:].any()      # neighbours untouched
    # input read through a channel slice
    xin_buf = torch.randn(B, H, W, Cin + 64, device=dev, generator=g).bfloat16()
    xs = xin_buf[..., 64:]
    _check(ops.conv2d(xs, wk, b, impl="tc"), _ref(xs, w, b), "sliced input")


@pytest.mark.parametrize("B,H,W,Cin", [(2, 64, 64, 64), (1, 256, 256, 128), (3, 8, 8, 256)])
def test_conv_tc_head_nchw_f32(cuda_lib, B, H, W, Cin):
    """out.2: Cout 6 padded to 16, fp32 NCHW output (unet.py:151)."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, _ = _mk(B, H, W, Cin, 6, 3, seed=Cin)
    wk = ops.repack_weight(w.float(), cout_pad=16)
    b16 = torch.zeros(16, device="cuda")
    b16[:6] = b
    y = ops.conv2d(x, wk, b16, nchw_out_channels=6, impl="tc")
    want = _ref(x, w, b)
    assert y.shape == want.shape
    assert torch.allclose(y, want, atol=2e-3, rtol=2e-3)


def test_conv_tc_matches_simt_bf16(cuda_lib):
    from fidm_b200 import ops
    x, w, b, _ = _mk(2, 32, 32, 128, 128, 3, seed=5)
    wk = ops.repack_weight(w.float())
    a = ops.conv2d(x, wk, b, impl="tc").float()
    s = ops.conv2d(x, wk, b, impl="simt").float()
    assert torch.allclose(a, s, atol=3e-2, rtol=2e-2)


def test_conv_tc_rejects_unsupported(cuda_lib):
    from fidm_b200 import ops
    x = torch.zeros(1, 8, 8, 48, device="cuda", dtype=torch.bfloat16)
    wk = torch.zeros(64, 3, 3, 48, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.conv2d(x, wk, None, impl="tc")


@pytest.mark.parametrize("B,H,W,Cin,Cout,Cin2", [(2, 16, 16, 128, 64, 64), (1, 64, 64, 64, 256, 128), (2, 8, 8, 256, 512, 0)])
def test_conv_tc_fp16_operands(cuda_lib, B, H, W, Cin, Cout, Cin2):
    """Normalized operands are fp16 (x, w); the fused skip source, residual and output stay bf16."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(Cin + Cout)
    x = torch.randn(B, H, W, Cin, device=dev, generator=g).half()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) / math.sqrt(Cin * 9)).half()
    b = torch.randn(Cout, device=dev, generator=g)
    res = torch.randn(B, H, W, Cout, device=dev, generator=g).bfloat16()
    kw, x2, w2 = {}, None, None
    if Cin2:
        x2 = torch.randn(B, H, W, Cin2, device=dev, generator=g).bfloat16()
        w2 = (torch.randn(Cout, Cin2, 1, 1, device=dev, generator=g) / math.sqrt(Cin2)).bfloat16()
        kw = dict(x2=x2, w2=ops.repack_weight(w2.float(), torch.bfloat16))
    y = ops.conv2d(x, ops.repack_weight(w.float(), torch.float16), b, residual=res, impl="tc", **kw)
    assert y.dtype == torch.bfloat16
    _check(y, _ref(x, w, b, residual=res, x2=x2, w2=w2), "fp16 operands")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 64, 256), (3, 8, 8, 128, 128), (1, 256, 256, 64, 256),
                                            (2, 16, 16, 64, 192), (5, 32, 32, 64, 64)])
def test_conv_tc_fused_groupnorm_statistics(cuda_lib, B, H, W, Cin, Cout):
    """The conv epilogue's per-channel sums == sums of the stored bf16 output, and GroupNorm driven by them
    == GroupNorm with its own statistics pass."""
    from fidm_b200 import ops
    x, w, b, g = _mk(B, H, W, Cin, Cout, 3, seed=H + Cout)
    y, cs = ops.conv2d(x, ops.repack_weight(w.float()), b, impl="tc", want_chansum=True)
    yf = y.float()
    want = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], dim=-1)
    assert torch.allclose(cs, want, rtol=2e-4, atol=2e-2)
    gamma = 1 + 0.1 * torch.randn(Cout, device="cuda", generator=g)
    beta = 0.1 * torch.randn(Cout, device="cuda", generator=g)
    a = ops.groupnorm_silu(y, gamma, beta, silu=True, out_dtype=torch.float16, chansum=cs)
    r = ops.groupnorm_silu(y, gamma, beta, silu=True, out_dtype=torch.float16)
    assert torch.allclose(a.float(), r.float(), atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout,Cin2", [(8, 8, 8, 1024, 1024, 0), (8, 8, 8, 2048, 1024, 2048), (2, 16, 16, 512, 512, 0),
                                                 (1, 8, 8, 1024, 3072, 0), (8, 16, 16, 1024, 1024, 0)])
def test_conv_tc_split_k_layers(cuda_lib, B, H, W, Cin, Cout, Cin2):
    """Low-resolution layers take the split-K path (fixed-order fold): correct, bit-reproducible, and the
    workspace counters re-arm themselves (back-to-back launches)."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, g = _mk(B, H, W, Cin, Cout, 3, seed=Cin + H)
    wk = ops.repack_weight(w.float())
    res = torch.randn(B, H, W, Cout, device="cuda", generator=g).bfloat16()
    kw, x2, w2 = {}, None, None
    if Cin2:
        x2 = torch.randn(B, H, W, Cin2, device="cuda", generator=g).bfloat16()
        w2 = (torch.randn(Cout, Cin2, 1, 1, device="cuda", generator=g) / math.sqrt(Cin2)).bfloat16()
        kw = dict(x2=x2, w2=ops.repack_weight(w2.float()))
    outs = [ops.conv2d(x, wk, b, residual=res, impl="tc", **kw).clone() for _ in range(3)]
    _check(outs[0], _ref(x, w, b, residual=res, x2=x2, w2=w2), "split-K")
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 8, 512, 512), (3, 8, 8, 512, 256), (1, 16, 16, 512, 512), (8, 16, 16, 512, 512),
                                            (1, 32, 32, 512, 512), (5, 8, 16, 1024, 1024), (8, 8, 8, 512, 512),
                                            (16, 6, 6, 512, 512), (40, 2, 2, 512, 256), (4, 12, 12, 512, 512)])
def test_conv_tc_split_k_cluster_epilogue(cuda_lib, B, H, W, Cin, Cout):
    """The cluster split-K fold (partials parked in the workspace, one cluster barrier, each CTA of the cluster finishing
    128 / S rows of the tile): timestep-embedding rows, residual, and output into a channel slice of a wider concat
    buffer whose other channels must stay untouched; odd batches leave phantom rows in the last tile.  Widths 6 and 2
    give pixel boxes narrower than the fold's 4-pixel quads: those layers must take the workspace fold instead."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, g = _mk(B, H, W, Cin, Cout, 3, seed=Cin + H + B)
    wk = ops.repack_weight(w.float())
    res = torch.randn(B, H, W, Cout, device="cuda", generator=g).bfloat16()
    emb = torch.randn(B, Cout + 64, device="cuda", generator=g)[:, 64:]           # strided rows
    buf = torch.full((B, H, W, Cout + 128), 7.0, device="cuda", dtype=torch.bfloat16)
    out = buf[..., 64:64 + Cout]
    outs = []
    for _ in range(2):
        ops.conv2d(x, wk, b, row_add=emb, residual=res, out=out, impl="tc")
        outs.append(out.clone())
    _check(outs[0], _ref(x, w, b, row_add=emb, residual=res), "cluster split-K")
    assert torch.equal(outs[0], outs[1])
    assert (buf[..., :64] == 7.0).all() and (buf[..., 64 + Cout:] == 7.0).all()


_WORKSPACE_FOLD_SCRIPT = r"""
import math, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.nn.functional as Fn
import fidm_b200
from fidm_b200 import ops
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(5)
for B, H, Cin, Cout in ((1, 8, 512, 512), (8, 8, 1024, 1024), (3, 16, 512, 512), (8, 16, 1024, 512)):
    x = torch.randn(B, H, H, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Cin)).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(B, H, H, Cout, device="cuda", generator=g).bfloat16()
    wk = ops.repack_weight(w.float())
    ys = [ops.conv2d(x, wk, b, residual=res, impl="tc").clone() for _ in range(3)]
    want = Fn.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1) + res.float().permute(0, 3, 1, 2)
    got = ys[0].float().permute(0, 3, 1, 2)
    rel = ((got - want).norm() / want.norm()).item()
    assert rel < 4e-3, (B, H, Cin, Cout, rel)
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[1], ys[2])
print("workspace fold ok")
"""


def test_conv_tc_workspace_fold_fallback(cuda_lib):
    """FIDM_CONV_SPLIT_CLUSTER=0 keeps the round-1 split-K fold (arrival counter in the workspace, the last CTA of a tile
    folds): still correct, bit-reproducible, counters re-armed.  The switch is read once per process -> subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FIDM_CONV_SPLIT_CLUSTER="0")
    r = subprocess.run([sys.executable, "-c", _WORKSPACE_FOLD_SCRIPT, root], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "workspace fold ok" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 64, 64, 128, 128), (3, 32, 48, 64, 256),
                                            (8, 256, 256, 128, 128), (2, 128, 128, 256, 256), (1, 8, 8, 512, 512)])
def test_conv_tc_stride2(cuda_lib, B, H, W, Cin, Cout):
    """Downsample(use_conv=True) (nn.py:115-133, :126): 3x3 stride 2 pad 1 on the tensor-core kernel -- the A tile of
    tap (r,s) is a TMA box that loads every other pixel (element strides 2), out-of-image taps zero-filled."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, g = _mk(B, H, W, Cin, Cout, 3, seed=H + Cin + 1)
    res = torch.randn(B, H // 2, W // 2, Cout, device="cuda", generator=g).bfloat16()
    y = ops.conv2d(x, ops.repack_weight(w.float()), b, stride=2, residual=res, impl="tc")
    torch.cuda.synchronize()
    want = Fn.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=2, padding=1) + res.float().permute(0, 3, 1, 2)
    assert y.shape == (B, H // 2, W // 2, Cout)
    _check(y, want, ("stride2", B, H, W, Cin, Cout))
    # the SIMT kernel (fp32 verification mode / odd shapes) gives the same numbers
    y2 = ops.conv2d(x, ops.repack_weight(w.float()), b, stride=2, residual=res, impl="simt")
    _check(y2, want, ("stride2 simt", B, H, W, Cin, Cout))
