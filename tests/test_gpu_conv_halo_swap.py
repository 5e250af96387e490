"""GPU: K1s (conv_halo_swap.cu) -- the fused GroupNorm+SiLU operand path with the operand roles swapped (weights as the
A operand, 256 pixels as N) that the 128-channel layers of REF-FFHQ256 and the 6-channel head take -- against torch fp32
(group_norm -> silu -> conv2d on the same bf16 input).  Same contract as tests/test_gpu_conv_halo.py."""
import math

import pytest
import torch
import torch.nn.functional as Fn

from test_gpu_conv_halo import _act, _check, _mk

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wdtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,Cin,Cout,ld_extra", [
    (1, 16, 16, 64, 128, 0),          # one tile
    (2, 32, 48, 128, 128, 0),         # non-square, image borders on every side of some tile
    (3, 64, 64, 256, 384, 64),        # three output-channel blocks, input is a channel slice
    (2, 256, 256, 128, 128, 0),       # 512 tiles on 148 SMs: persistent loop, TMEM double buffering, ring wrap-around
])
def test_conv_swap_plain(cuda_lib, B, H, W, Cin, Cout, ld_extra, wdtype):
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, gamma, beta, _ = _mk(B, H, W, Cin, Cout, seed=H + Cin + Cout, ld_extra=ld_extra)
    wq = w.to(wdtype)
    coef = ops.groupnorm_silu_coeff(x, gamma, beta)
    y = ops.conv2d(x, ops.repack_weight(wq.float(), wdtype), b, impl="tc", gn_coef=coef)
    torch.cuda.synchronize()
    want = Fn.conv2d(_act(x, gamma, beta), wq.float(), b, padding=1)
    _check(y, want, (B, H, W, Cin, Cout), rel_tol=4e-3 if wdtype == torch.float16 else 8e-3)
    y2 = ops.conv2d(x, ops.repack_weight(wq.float(), wdtype), b, impl="tc", gn_coef=coef)
    assert torch.equal(y, y2)                       # bit-reproducible


@pytest.mark.parametrize("B,H,W,Cin,Cout,Cin2", [(2, 32, 32, 128, 128, 64), (1, 64, 64, 64, 128, 256),
                                                 (2, 128, 128, 128, 128, 256), (8, 64, 64, 256, 128, 384)])
def test_conv_swap_epilogue_skip_and_stats(cuda_lib, B, H, W, Cin, Cout, Cin2):
    """scale/shift coefficients; bias + timestep row + residual (TMA-loaded, updated in place); the fused 1x1 skip
    source; output into a channel slice of a concat buffer; fused output statistics."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    x, w, b, gamma, beta, g = _mk(B, H, W, Cin, Cout, seed=Cin2 + H)
    wq = w.half()
    wk = ops.repack_weight(wq.float(), torch.float16)
    ss = torch.randn(B, 2 * Cin, device=dev, generator=g) * 0.3
    coef = ops.groupnorm_silu_coeff(x, gamma, beta, scale_shift=ss)
    act = _act(x, gamma, beta, ss)
    row = torch.randn(B, Cout, device=dev, generator=g)
    res_buf = torch.randn(B, H, W, Cout + 64, device=dev, generator=g).bfloat16()
    res = res_buf[..., 64:]
    y, cs = ops.conv2d(x, wk, b, row_add=row, residual=res, impl="tc", gn_coef=coef, want_chansum=True)
    want = Fn.conv2d(act, wq.float(), b, padding=1) + row[:, :, None, None] + res.float().permute(0, 3, 1, 2)
    _check(y, want, "epilogue")
    yf = y.float()
    assert torch.allclose(cs[..., 0], yf.sum(dim=(1, 2)), rtol=1e-3, atol=0.5)
    assert torch.allclose(cs[..., 1], (yf * yf).sum(dim=(1, 2)), rtol=1e-3, atol=0.5)
    # 1x1 skip source as extra K iterations, output into a slice, fused statistics of the output
    x2 = torch.randn(B, H, W, Cin2 + 64, device=dev, generator=g).bfloat16()[..., 64:]
    w2 = (torch.randn(Cout, Cin2, 1, 1, device=dev, generator=g) / math.sqrt(Cin2)).bfloat16()
    out_buf = torch.zeros(B, H, W, Cout + 128, device=dev, dtype=torch.bfloat16)
    out = out_buf[..., 64:64 + Cout]
    _, chansum = ops.conv2d(x, wk, b, x2=x2, w2=ops.repack_weight(w2.float()), out=out, impl="tc", gn_coef=coef,
                            want_chansum=True)
    want = Fn.conv2d(act, wq.float(), b, padding=1) + Fn.conv2d(x2.float().permute(0, 3, 1, 2), w2.float())
    _check(out, want, "skip")
    assert not out_buf[..., :64].any() and not out_buf[..., 64 + Cout:].any()
    of = out.float()
    assert torch.allclose(chansum[..., 0], of.sum(dim=(1, 2)), rtol=1e-3, atol=0.5)
    assert torch.allclose(chansum[..., 1], (of * of).sum(dim=(1, 2)), rtol=1e-3, atol=0.5)


def test_conv_swap_matches_two_pass_path(cuda_lib):
    """The same layer through GroupNorm apply -> fp16 tensor -> K1: the fused operand path must agree to the fp16
    rounding of the operand."""
    from fidm_b200 import ops
    x, w, b, gamma, beta, _ = _mk(2, 128, 128, 128, 128, seed=17)
    wk = ops.repack_weight(w.half().float(), torch.float16)
    a = ops.groupnorm_silu(x, gamma, beta, silu=True, out_dtype=torch.float16)
    y0 = ops.conv2d(a, wk, b, impl="tc").float()
    y1 = ops.conv2d(x, wk, b, impl="tc", gn_coef=ops.groupnorm_silu_coeff(x, gamma, beta)).float()
    assert ((y0 - y1).norm() / y0.norm()).item() < 3e-3


@pytest.mark.parametrize("B,H,W,Cin", [(1, 16, 16, 64), (8, 256, 256, 128), (3, 32, 64, 256)])
def test_conv_swap_head(cuda_lib, B, H, W, Cin):
    """The 6-channel head (unet.py:148-152) through K1s: 16 weight rows, fp32 NCHW stores from the accumulator,
    bias and an additive per-image row."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    x, w, b, gamma, beta, g = _mk(B, H, W, Cin, 6, seed=Cin + W + B)
    wq = w.half()
    wk = ops.repack_weight(wq.float(), torch.float16, cout_pad=16)
    b16 = torch.zeros(16, device="cuda")
    b16[:6] = b
    coef = ops.groupnorm_silu_coeff(x, gamma, beta)
    y = ops.conv2d(x, wk, b16, nchw_out_channels=6, impl="tc", gn_coef=coef)
    want = Fn.conv2d(_act(x, gamma, beta), wq.float(), b, padding=1)
    assert y.shape == want.shape
    rel = ((y - want).norm() / want.norm()).item()
    assert rel < 3e-3, rel


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 16, 16, 64, 128), (2, 64, 48, 64, 256), (8, 256, 256, 64, 128)])
def test_conv_swap_halo_copy_stem(cuda_lib, B, H, W, Cin, Cout):
    """halo_copy: the stem convolution (unet.py:55: no GroupNorm in front) through the swapped-role kernel with the
    transform reduced to a shifted copy -- plain conv3x3 of an fp16 input, bias, fused output statistics."""
    from fidm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B + H + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).half()
    x[..., 9:] = 0                                     # the packed 9-channel network input, zero-padded to 64
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * 9)).half()
    b = torch.randn(Cout, device="cuda", generator=g)
    y, cs = ops.conv2d(x, ops.repack_weight(w.float(), torch.float16), b, impl="tc", halo_copy=True, want_chansum=True)
    y0 = ops.conv2d(x, ops.repack_weight(w.float(), torch.float16), b, impl="tc")
    want = Fn.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1)
    _check(y, want, ("halo_copy", B, H, W, Cin, Cout))
    _check(y0, want, ("K1", B, H, W, Cin, Cout))
    yf = y.float()
    assert torch.allclose(cs[..., 0], yf.sum(dim=(1, 2)), rtol=1e-3, atol=0.5)
    assert torch.allclose(cs[..., 1], (yf * yf).sum(dim=(1, 2)), rtol=1e-3, atol=0.5)
