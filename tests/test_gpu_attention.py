"""Split-K decode attention against the oracle's fp32 softmax attention (tolerance: max-abs 2e-2, cos>0.999)."""
import pytest
import torch

from oracle import shadowkv_oracle as O
from tests.tolerances import assert_close_bf16

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,G,Lp,splits", [(1, 8, 4, 2497, 0), (2, 4, 8, 2592, 0), (3, 2, 16, 700, 0),
                                             (1, 8, 4, 1, 0), (1, 8, 4, 63, 0), (1, 2, 1, 130, 0),
                                             (4, 8, 4, 2592, 1), (1, 8, 4, 2592, 7), (1, 8, 4, 4200, 64)])
def test_flash_attn_with_kvcache(cuda, B, H, G, Lp, splits):
    from shadowkv_b200.attention import flash_attn_with_kvcache
    g = torch.Generator().manual_seed(B * 1000 + Lp)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16)
    Lbuf = Lp + 17
    K = torch.randn(B, H, Lbuf, 128, generator=g).to(torch.bfloat16)
    V = torch.randn(B, H, Lbuf, 128, generator=g).to(torch.bfloat16)
    dK, dV = K.to(cuda), V.to(cuda)
    # exactly the call of base.py:341: strided [B, L', H, D] views of [B, H, Lbuf, D] buffers
    out = flash_attn_with_kvcache(q=q.to(cuda).transpose(1, 2), k_cache=dK[:, :, :Lp].transpose(1, 2),
                                  v_cache=dV[:, :, :Lp].transpose(1, 2), causal=True, num_splits=splits)
    torch.cuda.synchronize()
    want = O.attention(q, K[:, :, :Lp], V[:, :, :Lp])
    assert out.shape == (B, 1, H * G, 128)
    assert_close_bf16(out, want, what="attention")
    # second call reuses the workspace (arrival counters must be back at zero)
    out2 = flash_attn_with_kvcache(q=q.to(cuda).transpose(1, 2), k_cache=dK[:, :, :Lp].transpose(1, 2),
                                   v_cache=dV[:, :, :Lp].transpose(1, 2), causal=True, num_splits=splits)
    assert torch.equal(out, out2)


def test_attention_peaked_softmax(cuda):
    """One key dominates: output must equal that key's value row (exercises the running-max rescale)."""
    from shadowkv_b200.attention import flash_attn_with_kvcache
    B, H, G, Lp = 1, 2, 4, 1500
    q = torch.ones(B, H * G, 1, 128, dtype=torch.bfloat16)
    K = torch.zeros(B, H, Lp, 128, dtype=torch.bfloat16)
    K[:, :, 1234] = 8.0
    V = torch.randn(B, H, Lp, 128).to(torch.bfloat16)
    out = flash_attn_with_kvcache(q.to(cuda).transpose(1, 2), K.to(cuda).transpose(1, 2), V.to(cuda).transpose(1, 2))
    want = V[:, :, 1234].repeat_interleave(G, dim=1).view(B, 1, H * G, 128)
    assert_close_bf16(out, want, max_abs=1e-2, what="peaked")
