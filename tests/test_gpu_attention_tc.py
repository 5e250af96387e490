"""GPU: K3 tcgen05 flash attention against the reference formulation (nn.py:222-235) in fp32."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _attn_ref(qkv, heads):
    n, t, c3 = qkv.shape
    q, k, v = qkv.float().permute(0, 2, 1).chunk(3, dim=1)
    ch = c3 // 3 // heads
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", (q * s).reshape(n * heads, ch, t), (k * s).reshape(n * heads, ch, t))
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(n * heads, ch, t)).reshape(n, -1, t)
    return a.permute(0, 2, 1)


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("B,T,heads", [(1, 128, 1), (2, 64, 4), (1, 256, 8), (2, 1024, 8), (3, 192, 2)])
def test_attention_tc(cuda_lib, B, T, heads, d):
    """head dims 64 and 128 (num_heads-defined heads of 512-channel blocks, nn.py:245-249) on the tcgen05 kernel"""
    from fidm_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(T + heads + d)
    buf = (torch.randn(B, T, 3 * heads * d + 64, device="cuda", generator=g) * 1.5).bfloat16()
    qkv = buf[..., 64:]                                    # exercised through a strided view
    y = ops.attention(qkv, heads, impl="tc")
    torch.cuda.synchronize()
    want = _attn_ref(qkv, heads)
    rel = ((y.float() - want).norm() / want.norm()).item()
    assert rel < 1e-2 and torch.allclose(y.float(), want, atol=3e-2, rtol=3e-2), rel
    s = ops.attention(qkv.contiguous(), heads, impl="simt")
    assert torch.allclose(y.float(), s.float(), atol=3e-2, rtol=3e-2)
