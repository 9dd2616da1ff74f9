"""The six RoPE entry points against the oracle's bf16 restatement (bit-exact: same operation order)."""
import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth

pytestmark = pytest.mark.gpu


def _x(B, H, S, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, H, S, 128, generator=g) * 2).to(torch.bfloat16), g


def test_apply_rotary_pos_emb_new_and_wrapper(cuda):
    from shadowkv_b200.tensor_op import apply_rotary_pos_emb_cuda
    x, g = _x(2, 8, 40, 0)
    cs = synth.cos_sin_table(5000)
    pos = torch.randint(0, 5000, (2, 8, 40), generator=g)
    got = apply_rotary_pos_emb_cuda(x.to(cuda), cs.to(cuda), pos.to(cuda))
    assert torch.equal(got.cpu(), O.rope_bf16(x, cs[pos]))
    # strided (expanded) position ids as ShadowKVCache.get_key_cache passes through llama.py:280
    pos1 = torch.randint(0, 5000, (2, 1, 40), generator=g).expand(2, 8, 40)
    got = apply_rotary_pos_emb_cuda(x.to(cuda), cs.to(cuda), pos1.to(cuda))
    assert torch.equal(got.cpu(), O.rope_bf16(x, cs[pos1]))


def test_apply_rotary_pos_emb_separate_tables(cuda):
    from shadowkv_b200 import shadowkv
    x, g = _x(1, 4, 16, 1)
    cs = synth.cos_sin_table(600)
    cos = torch.cat([cs[:, :64], cs[:, :64]], -1).contiguous()
    sin = torch.cat([cs[:, 64:], cs[:, 64:]], -1).contiguous()
    pos = torch.randint(0, 600, (1, 4, 16), generator=g)
    out = torch.empty_like(x, device=cuda)
    xd = x.to(cuda)
    shadowkv.apply_rotary_pos_emb(xd, cos.to(cuda), sin.to(cuda), pos.to(cuda), out, 1, 4, 16, 128, *xd.stride(),
                                  128, 128, *pos.stride(), 64)
    assert torch.equal(out.cpu(), O.rope_bf16(x, cs[pos]))


def test_apply_rotary_pos_emb_new_v2_chunk_ids(cuda):
    from shadowkv_b200 import shadowkv
    x, g = _x(2, 4, 64, 2)
    cs = synth.cos_sin_table(4096)
    ids = torch.randint(0, 500, (2, 4, 8), generator=g).to(torch.int32)
    tok = (ids.long().unsqueeze(-1) * 8 + torch.arange(8)).view(2, 4, 64)
    out = torch.empty_like(x, device=cuda)
    xd = x.to(cuda)
    shadowkv.apply_rotary_pos_emb_new_v2(xd, cs.to(cuda), ids.to(cuda), out, 2, 4, 64, 128, *xd.stride(), 128,
                                         *ids.stride(), 64, 8)
    assert torch.equal(out.cpu(), O.rope_bf16(x, cs[tok]))


@pytest.mark.parametrize("glm", [False, True])
@pytest.mark.parametrize("variant", ["opt", "plain"])
def test_push_cache(cuda, glm, variant):
    from shadowkv_b200 import shadowkv
    from shadowkv_b200.tensor_op import apply_rotary_pos_emb_cuda_push_cache
    B, H, S = 2, 4, 32
    x, g = _x(B, H, S * 8, 3)
    cs = synth.cos_sin_table(8192, glm=glm)
    ids = torch.stack([torch.randperm(1000, generator=g)[:S] for _ in range(B * H)]).view(B, H, S).to(torch.int32)
    cnts = torch.tensor([0, 5, 32, 17, 1, 31, 8, 16], dtype=torch.int32)
    ss, Lbuf = 64, 64 + S * 8 + 40
    cache = torch.randn(B, H, Lbuf, 128, generator=g).to(torch.bfloat16)
    tok = (ids.long().unsqueeze(-1) * 8 + torch.arange(8)).view(B, H, S * 8)
    want = cache.clone()
    rot = O.rope_bf16(x, cs[tok])
    for b in range(B):
        for h in range(H):
            c = int(cnts[b * H + h]) * 8
            want[b, h, ss + c:ss + S * 8] = rot[b, h, c:]
    d_cache = cache.to(cuda)
    if variant == "opt":
        apply_rotary_pos_emb_cuda_push_cache(x.to(cuda), cs.to(cuda), ids.to(cuda), 8, d_cache, ss, ss + S * 8,
                                             cnts.to(cuda))
    else:
        if glm:
            pytest.skip("the reference has no plain GLM variant")
        xd = x.to(cuda)
        shadowkv.apply_rotary_pos_emb_push_cache(xd, cs.to(cuda), ids.to(cuda), d_cache, cnts.to(cuda), B, H, S * 8, 128,
                                                 *xd.stride(), 128, *ids.stride(), *d_cache.stride()[:3], ss,
                                                 ss + S * 8, 64, 8)
    torch.cuda.synchronize()
    assert torch.equal(d_cache.cpu(), want)
