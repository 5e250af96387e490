"""GPU: K2 GroupNorm(+scale/shift)(+SiLU)(+resample) against torch fp32, NHWC with pixel strides."""
import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu


def _ref(x_nchw, gamma, beta, ss, silu, resample):
    y = Fn.group_norm(x_nchw, 32, gamma, beta, eps=1e-5)
    if ss is not None:
        C = x_nchw.shape[1]
        y = y * (1 + ss[:, :C, None, None]) + ss[:, C:2 * C, None, None]
    if silu:
        y = Fn.silu(y)
    raw = x_nchw
    if resample == "down":
        y, raw = Fn.avg_pool2d(y, 2), Fn.avg_pool2d(raw, 2)
    elif resample == "up":
        y, raw = Fn.interpolate(y, scale_factor=2, mode="nearest"), Fn.interpolate(raw, scale_factor=2, mode="nearest")
    return y, raw


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,H,W,ld_extra", [(64, 16, 16, 0), (128, 8, 8, 64), (192, 16, 8, 0), (256, 64, 64, 256),
                                            (768, 16, 16, 0), (1024, 8, 8, 0), (2048, 8, 8, 0), (96, 4, 4, 0)])
def test_groupnorm_modes(cuda_lib, dtype, C, H, W, ld_extra):
    from fidm_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(C + H)
    B = 3
    buf = torch.randn(B, H, W, C + ld_extra, device=dev, generator=g) * 2 + 0.5
    buf = buf.to(dtype)
    x = buf[..., ld_extra // 2: ld_extra // 2 + C] if ld_extra else buf
    gamma = 1 + 0.2 * torch.randn(C, device=dev, generator=g)
    beta = 0.2 * torch.randn(C, device=dev, generator=g)
    emb = torch.randn(B, 2 * C + 32, device=dev, generator=g) * 0.3
    ss = emb[:, 16:16 + 2 * C]
    xr = x.float().permute(0, 3, 1, 2).contiguous()
    tol = dict(atol=2e-5, rtol=2e-5) if dtype == torch.float32 else dict(atol=2.5e-2, rtol=1.6e-2)
    for silu in (True, False):
        for use_ss in (False, True):
            for resample in ("none", "down", "up"):
                if resample == "down" and (H % 2 or W % 2):
                    continue
                y, raw = ops.groupnorm_silu(x, gamma, beta, scale_shift=ss if use_ss else None, silu=silu,
                                            resample=resample, want_raw=True)
                wy, wraw = _ref(xr, gamma, beta, ss if use_ss else None, silu, resample)
                assert torch.allclose(y.float().permute(0, 3, 1, 2), wy, **tol), (silu, use_ss, resample)
                if resample != "none":
                    assert torch.allclose(raw.float().permute(0, 3, 1, 2), wraw, **tol)


def test_groupnorm_fp16_output(cuda_lib):
    """bf16 stream in, fp16 normalized operand out (8x finer rounding than bf16)."""
    from fidm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(2, 16, 16, 256, device="cuda", generator=g) * 3).bfloat16()
    gamma = 1 + 0.2 * torch.randn(256, device="cuda", generator=g)
    beta = 0.2 * torch.randn(256, device="cuda", generator=g)
    for resample in ("none", "down", "up"):
        y, raw = ops.groupnorm_silu(x, gamma, beta, silu=True, resample=resample, want_raw=True, out_dtype=torch.float16)
        wy, wraw = _ref(x.float().permute(0, 3, 1, 2), gamma, beta, None, True, resample)
        assert y.dtype == torch.float16 and raw.dtype == torch.bfloat16
        assert torch.allclose(y.float().permute(0, 3, 1, 2), wy, atol=3e-3, rtol=2e-3)
        if resample != "none":
            assert torch.allclose(raw.float().permute(0, 3, 1, 2), wraw, atol=3e-2, rtol=1.6e-2)


@pytest.mark.parametrize("B,H,W,C", [(1, 64, 64, 512), (2, 64, 64, 512), (3, 64, 64, 512), (1, 64, 32, 1024), (1, 32, 32, 512), (1, 128, 128, 256),
                                     (2, 24, 40, 256)])
def test_groupnorm_one_launch_small_batch(cuda_lib, B, H, W, C):
    """Batch-1/2 tensors of 8192 vectors per (image, group) take the one-launch kernel as 4-block clusters (batch x
    groups <= 64: the (sum, sum of squares) partials cross the cluster through distributed shared memory); batch 3 of the
    same shape keeps the two-launch path -- same numbers either way."""
    import ctypes
    from fidm_b200 import _lib as L
    from fidm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + C)
    x = (torch.randn(B, H, W, C, device="cuda", generator=g) * 2 + 0.3).bfloat16()
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.2 * torch.randn(C, device="cuda", generator=g)
    ss = torch.randn(B, 2 * C, device="cuda", generator=g) * 0.3
    a = L.GnArgs()
    a.dtype, a.y_dtype = L.BF16, L.F16
    a.batch, a.height, a.width, a.channels, a.groups = B, H, W, C, 32
    a.ld_x = a.ld_y = C
    a.x = a.y = L.ptr(x)                                           # only alignment is inspected
    want_launches = 1 if (B * 32 <= 64 and H * W * (C // 32 // 8) <= 16384) or H * W * (C // 32 // 8) <= 4096 else 2
    assert cuda_lib.fidm_groupnorm_num_launches(ctypes.byref(a)) == want_launches
    for use_ss in (False, True):
        y = ops.groupnorm_silu(x, gamma, beta, scale_shift=ss if use_ss else None, silu=True, out_dtype=torch.float16)
        wy, _ = _ref(x.float().permute(0, 3, 1, 2), gamma, beta, ss if use_ss else None, True, "none")
        assert torch.allclose(y.float().permute(0, 3, 1, 2), wy, atol=3e-3, rtol=2e-3), use_ss


def test_groupnorm_skip_norm_resample(cuda_lib):
    from fidm_b200 import ops
    x = torch.randn(2, 8, 8, 64, device="cuda")
    xr = x.permute(0, 3, 1, 2)
    d = ops.groupnorm_silu(x, silu=False, resample="down", skip_norm=True)
    assert torch.allclose(d.permute(0, 3, 1, 2), Fn.avg_pool2d(xr, 2), atol=1e-6)
    u = ops.groupnorm_silu(x, silu=False, resample="up", skip_norm=True)
    assert torch.equal(u.permute(0, 3, 1, 2), Fn.interpolate(xr, scale_factor=2, mode="nearest"))


def test_groupnorm_large_mean_is_stable(cuda_lib):
    """sum / sum-of-squares statistics are accumulated in double across blocks."""
    from fidm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(1, 128, 128, 64, device="cuda", generator=g) * 0.5 + 30.0
    y = ops.groupnorm_silu(x, silu=False)
    want = Fn.group_norm(x.permute(0, 3, 1, 2).double(), 32).float()
    assert torch.allclose(y.permute(0, 3, 1, 2), want, atol=5e-3, rtol=1e-3)
