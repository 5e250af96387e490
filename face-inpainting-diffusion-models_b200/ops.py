"""Tensor-level wrappers of the C-ABI kernels (one call = one kernel invocation).

The engine (`engine.Plan`) pre-builds its argument structs and does not go through these; they are the
operator-level API used by the parity tests and by callers that want a single fused op.
All activations are NHWC ([N,H,W,C], last dim contiguous; a channel slice of a wider buffer is
expressed by passing the slice view -- its pixel stride is taken from `.stride(2)`).
"""
import ctypes as C

import torch

from . import _lib as L


_SPLITK_WS = None


def _nhwc(t):
    assert t.dim() == 4 and t.stride(3) == 1, "NHWC tensor with contiguous channels expected"
    n, h, w, c = t.shape
    ld = t.stride(2)
    assert t.stride(1) == w * ld and t.stride(0) == h * w * ld, "pixels must be densely strided"
    return n, h, w, c, ld


def pack_nchw_to_nhwc(sources, dtype=torch.bfloat16, c_pad=64):
    """sources: list of (fp32 NCHW tensor, repeat)."""
    x0 = sources[0][0]
    n, _, h, w = x0.shape
    out = torch.empty(n, h, w, c_pad, device=x0.device, dtype=dtype)
    a = L.PackArgs()
    a.batch, a.hw, a.n_src = n, h * w, len(sources)
    keep = []
    for i, (t, r) in enumerate(sources):
        t = t.float().contiguous()
        keep.append(t)
        a.src[i], a.src_channels[i], a.src_repeat[i] = t.data_ptr(), t.shape[1], r
    a.dst, a.dst_dtype, a.ld_dst, a.c_pad = out.data_ptr(), L.dtype_code(dtype), c_pad, c_pad
    L.check(L.lib().fidm_pack_nchw_to_nhwc(a, L.stream()), "pack")
    return out


def unpack_nhwc_to_nchw(x, channels=None):
    n, h, w, c, ld = _nhwc(x)
    channels = channels or c
    out = torch.empty(n, channels, h, w, device=x.device, dtype=torch.float32)
    L.check(L.lib().fidm_unpack_nhwc_to_nchw(L.ptr(x), L.dtype_code(x.dtype), ld, L.ptr(out), n, h * w, channels,
                                              L.stream()), "unpack")
    return out


def repack_weight(w, dtype=torch.bfloat16, cout_pad=None, cin_pad=None):
    """OIHW fp32 -> KRSC."""
    if w.dim() == 3:
        w = w.unsqueeze(-1)
    w = w.float().contiguous()
    cout, cin, ks, _ = w.shape
    cout_pad, cin_pad = cout_pad or cout, cin_pad or cin
    out = torch.empty(cout_pad, ks, ks, cin_pad, device=w.device, dtype=dtype)
    L.check(L.lib().fidm_repack_weight_oihw_to_krsc(L.ptr(w), L.ptr(out), L.dtype_code(dtype), cout, cin, ks,
                                                    cout_pad, cin_pad, L.stream()), "repack")
    return out


def quantize_weight_e4m3(w_krsc_f32):
    """KRSC fp32 weights -> (e4m3 bytes as uint8, per-output-channel fp32 scale): w ~= q * scale[c] with
    scale[c] = max|w[c]| / 448, the largest finite e4m3 (precision "fp8", engine.Weights)."""
    amax = w_krsc_f32.abs().amax(dim=(1, 2, 3)).clamp_min(1e-12)
    scale = (amax / 448.0).contiguous()
    q = (w_krsc_f32 / scale[:, None, None, None]).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    return q.view(torch.uint8).contiguous(), scale


def timestep_embedding(t, freqs, dim):
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=torch.float32)
    L.check(L.lib().fidm_timestep_embedding(L.ptr(t.float().contiguous()), L.ptr(freqs), L.ptr(out), t.shape[0], dim,
                                             L.stream()), "timestep_embedding")
    return out


def linear_small(x, w, bias=None, silu_input=False):
    x = x.float().contiguous()
    w = w.contiguous()
    out = torch.empty(x.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
    L.check(L.lib().fidm_linear_small(L.ptr(x), L.ptr(w), L.dtype_code(w.dtype), L.ptr(bias), L.ptr(out), x.shape[0],
                                      x.shape[1], w.shape[0], int(silu_input), L.stream()), "linear_small")
    return out


def groupnorm_silu(x, gamma=None, beta=None, *, scale_shift=None, silu=True, resample="none", want_raw=False,
                   skip_norm=False, out=None, groups=32, eps=1e-5, out_dtype=None, chansum=None):
    n, h, w, c, ld = _nhwc(x)
    mode = {"none": L.RESAMPLE_NONE, "down": L.RESAMPLE_DOWN, "up": L.RESAMPLE_UP}[resample]
    ho, wo = (h // 2, w // 2) if resample == "down" else ((2 * h, 2 * w) if resample == "up" else (h, w))
    y = out if out is not None else torch.empty(n, ho, wo, c, device=x.device, dtype=out_dtype or x.dtype)
    raw = torch.empty(n, ho, wo, c, device=x.device, dtype=x.dtype) if want_raw else None
    stats = torch.zeros(L.lib().fidm_groupnorm_workspace_bytes(n, groups) // 8, device=x.device, dtype=torch.float64)
    a = L.GnArgs()
    a.dtype, a.y_dtype = L.dtype_code(x.dtype), L.dtype_code(y.dtype)
    a.batch, a.height, a.width, a.channels, a.groups, a.eps = n, h, w, c, groups, eps
    a.x, a.ld_x = L.ptr(x), ld
    a.gamma, a.beta = L.ptr(gamma), L.ptr(beta)
    if scale_shift is not None:
        assert scale_shift.dtype == torch.float32 and scale_shift.stride(1) == 1
        a.scale_shift, a.ld_ss = L.ptr(scale_shift), scale_shift.stride(0)
    a.silu, a.resample, a.skip_norm = int(silu), mode, int(skip_norm)
    a.y, a.ld_y = L.ptr(y), _nhwc(y)[4]
    if raw is not None:
        a.y_raw, a.ld_raw = L.ptr(raw), c
    a.stats = L.ptr(stats)
    if chansum is not None:            # [N, ld, 2] fp32 channel sums from conv2d(..., want_chansum=True)
        a.chansum, a.ld_chansum = L.ptr(chansum), chansum.shape[1]
    L.check(L.lib().fidm_groupnorm_silu_nhwc(C.byref(a), L.stream()), "groupnorm")
    return (y, raw) if want_raw else y


def groupnorm_silu_coeff(x, gamma=None, beta=None, *, scale_shift=None, groups=32, eps=1e-5, chansum=None):
    """Statistics only: the halved affine coefficients [N, C, 2] of GroupNorm(+scale/shift)+SiLU, for
    conv2d(..., gn_coef=coef) which applies the activation while staging its operand tiles."""
    n, h, w, c, ld = _nhwc(x)
    coef = torch.empty(n, c, 2, device=x.device, dtype=torch.float32)
    stats = torch.zeros(L.lib().fidm_groupnorm_workspace_bytes(n, groups) // 8, device=x.device, dtype=torch.float64)
    a = L.GnArgs()
    a.dtype, a.y_dtype = L.dtype_code(x.dtype), L.dtype_code(x.dtype)
    a.batch, a.height, a.width, a.channels, a.groups, a.eps = n, h, w, c, groups, eps
    a.x, a.ld_x = L.ptr(x), ld
    a.gamma, a.beta = L.ptr(gamma), L.ptr(beta)
    if scale_shift is not None:
        assert scale_shift.dtype == torch.float32 and scale_shift.stride(1) == 1
        a.scale_shift, a.ld_ss = L.ptr(scale_shift), scale_shift.stride(0)
    a.silu = 1
    a.stats = L.ptr(stats)
    if chansum is not None:
        a.chansum, a.ld_chansum = L.ptr(chansum), chansum.shape[1]
    L.check(L.lib().fidm_groupnorm_silu_coeff(C.byref(a), L.ptr(coef), c, L.stream()), "groupnorm_coeff")
    return coef


def conv2d(x, w_krsc, bias=None, *, stride=1, row_add=None, residual=None, x2=None, w2=None, out=None,
           nchw_out_channels=None, impl="auto", want_chansum=False, gn_coef=None, x_half_res=False,
           residual_half_res=False, w_scale=None, halo_copy=False):
    """x NHWC, w_krsc [Cout,k,k,Cin].  impl: "tc" (tcgen05), "simt", or "auto".
    gn_coef [N, Cin, 2] (from groupnorm_silu_coeff): x is the raw bf16 stream and the conv operand is
    silu(GroupNorm(x)), applied inside the kernel; w_krsc's dtype (fp16 | bf16 | e4m3 bytes) is the staged operand's
    dtype.  w_scale [Cout] fp32: per-output-channel scale of e4m3 weights (see quantize_weight_e4m3)."""
    n, h, w, cin, ld = _nhwc(x)
    if x_half_res:            # x is the half-resolution source of an `up` ResBlock: the conv runs at 2h x 2w
        h, w = 2 * h, 2 * w
    cout, ks = w_krsc.shape[0], w_krsc.shape[1]
    ho, wo = ((h + 2 * (ks // 2) - ks) // stride + 1, (w + 2 * (ks // 2) - ks) // stride + 1)
    a = L.ConvArgs()
    a.dtype, a.batch, a.height, a.width = L.dtype_code(w_krsc.dtype if gn_coef is not None else x.dtype), n, h, w
    a.x_half_res, a.residual_half_res = int(x_half_res), int(residual_half_res)
    if gn_coef is not None:
        assert x.dtype == torch.bfloat16 and gn_coef.dtype == torch.float32 and gn_coef.shape[-1] == 2
        a.gn_coef, a.ld_gn_coef = L.ptr(gn_coef), gn_coef.stride(0) // 2
    a.cin, a.cout, a.ksize, a.stride = cin, cout, ks, stride
    a.x, a.ld_x, a.w = L.ptr(x), ld, L.ptr(w_krsc)
    if w_scale is not None:
        assert w_scale.dtype == torch.float32 and w_scale.numel() == cout
        a.w_scale = L.ptr(w_scale)
    a.halo_copy = int(halo_copy)         # stem: x through the halo kernel without an activation
    if x2 is not None:
        n2, h2, w2_, c2, ld2 = _nhwc(x2)
        assert (n2, h2, w2_) == (n, ho, wo)
        a.x2, a.ld_x2, a.cin2, a.w2 = L.ptr(x2), ld2, c2, L.ptr(w2)
    a.bias = L.ptr(bias)
    if row_add is not None:
        assert row_add.dtype == torch.float32 and row_add.stride(1) == 1
        a.row_add, a.ld_row_add = L.ptr(row_add), row_add.stride(0)
    if residual is not None:
        a.residual, a.ld_res = L.ptr(residual), _nhwc(residual)[4]
    if nchw_out_channels is not None:
        y = torch.empty(n, nchw_out_channels, ho, wo, device=x.device, dtype=torch.float32)
        a.y, a.y_nchw_f32, a.cout_valid = L.ptr(y), 1, nchw_out_channels
    else:
        ydt = torch.bfloat16 if x.dtype == torch.float16 else x.dtype
        y = out if out is not None else torch.empty(n, ho, wo, cout, device=x.device, dtype=ydt)
        a.y, a.ld_y, a.cout_valid = L.ptr(y), _nhwc(y)[4], cout
    if impl == "auto":
        impl = "tc" if (x.dtype in (torch.bfloat16, torch.float16) and cin % 64 == 0 and
                        (stride == 1 or (stride == 2 and ks == 3 and h % 2 == 0 and w % 2 == 0 and x2 is None)) and
                        (cout % 64 == 0 or (nchw_out_channels is not None and cout == 16 and stride == 1))) else "simt"
    fn = L.lib().fidm_conv2d_nhwc_bf16 if impl == "tc" else L.lib().fidm_conv2d_nhwc_simt
    if impl == "tc":
        global _SPLITK_WS
        if _SPLITK_WS is None or _SPLITK_WS.device != x.device:
            _SPLITK_WS = torch.zeros(32 << 20, device=x.device, dtype=torch.uint8)
        a.splitk_ws, a.splitk_ws_bytes = L.ptr(_SPLITK_WS), _SPLITK_WS.numel()
    colsum = None
    if want_chansum:
        slots = L.lib().fidm_conv_colsum_slots(ho, wo)
        assert impl == "tc" and slots > 0 and nchw_out_channels is None
        colsum = torch.empty(n, slots, cout, 2, device=x.device, dtype=torch.float32)
        a.colsum = L.ptr(colsum)
    L.check(fn(C.byref(a), L.stream()), "conv2d/" + impl)
    if want_chansum:
        chansum = torch.zeros(n, cout, 2, device=x.device, dtype=torch.float32)
        L.check(L.lib().fidm_groupnorm_reduce_colsum(L.ptr(colsum), n, slots, cout, L.ptr(chansum), cout, 0, L.stream()),
                "reduce_colsum")
        return y, chansum
    return y


def attention(qkv, heads, *, out=None, impl="auto"):
    """qkv: [N, T, 3C] (or [N,H,W,3C]) channel order [Q heads | K heads | V heads]."""
    if qkv.dim() == 4:
        qkv = qkv.reshape(qkv.shape[0], -1, qkv.shape[3]) if qkv.is_contiguous() else qkv.flatten(1, 2)
    n, t, c3 = qkv.shape
    ld = qkv.stride(1)
    cn = c3 // 3
    d = cn // heads
    y = out if out is not None else torch.empty(n, t, cn, device=qkv.device, dtype=qkv.dtype)
    a = L.AttnArgs()
    a.dtype, a.batch, a.tokens, a.heads, a.head_dim = L.dtype_code(qkv.dtype), n, t, heads, d
    a.qkv, a.ld_qkv, a.out, a.ld_out = L.ptr(qkv), ld, L.ptr(y), y.stride(1)
    if impl == "auto":
        impl = "tc" if (qkv.dtype == torch.bfloat16 and d in (64, 128) and t % 64 == 0) else "simt"
    fn = L.lib().fidm_attention_qkv_nhwc_bf16 if impl == "tc" else L.lib().fidm_attention_qkv_nhwc_simt
    L.check(fn(C.byref(a), L.stream()), "attention/" + impl)
    return y


def prepare_inputs_u8(image_u8, mask_u8):
    """uint8 images [N,H,W,3] + grayscale masks [N,H,W] (black = inpaint) -> the reference's dataset batch
    (data/dataset.py:130-150): dict(image, masked_image, mask, gt_keep_mask) as fp32 NCHW."""
    n, h, w, _ = image_u8.shape
    image_u8, mask_u8 = image_u8.contiguous(), mask_u8.contiguous()
    dev = image_u8.device
    image = torch.empty(n, 3, h, w, device=dev)
    masked = torch.empty(n, 3, h, w, device=dev)
    mask = torch.empty(n, 1, h, w, device=dev)
    keep = torch.empty(n, 1, h, w, device=dev)
    L.check(L.lib().fidm_prepare_inputs_u8(L.ptr(image_u8), L.ptr(mask_u8), L.ptr(image), L.ptr(masked), L.ptr(mask),
                                           L.ptr(keep), n, h * w, L.stream()), "prepare_inputs_u8")
    return {"image": image, "masked_image": masked, "mask": mask, "gt_keep_mask": keep}


def blend_to_u8(sample, gt=None, mask=None):
    """toU8(sample * mask + gt * (1 - mask)) (test_inp_ddim_100.py:33-41, 693-696): [N,H,W,3] uint8."""
    n, _, h, w = sample.shape
    out = torch.empty(n, h, w, 3, device=sample.device, dtype=torch.uint8)
    L.check(L.lib().fidm_blend_to_u8(L.ptr(sample.float().contiguous()), L.ptr(gt), L.ptr(mask), L.ptr(out), n, h * w,
                                     L.stream()), "blend_to_u8")
    return out
