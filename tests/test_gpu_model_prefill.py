"""Prompt path of the model mirror (reference models/base.py:163-187, 296-318, 500-550):
``Llama.prefill`` / ``batch_prefill`` / ``generate_from_prompt`` against the oracle's restatement of the same operator
sequence (oracle/dense_oracle.py ``prefill_logits``) and against the decode state the oracle builds from the GPU's own
keys and values (oracle/shadowkv_oracle.py ``prefill_state``)."""
import pytest
import torch

from oracle import dense_oracle as DO
from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.tolerances import assert_close_bf16, assert_ulp_bf16

pytestmark = pytest.mark.gpu


def _shape(glm):
    return synth.ModelShape("tiny-glm" if glm else "tiny-llama", 2, 1024, 8, 2, 1536, 2048,
                            rope_theta=10000.0 if glm else 500000.0, glm_rope=glm)


def _model(cuda, glm, B, L, budget=1024):
    from shadowkv_b200.llama import Llama
    return Llama(_shape(glm), batch_size=B, max_length=L, device=str(cuda), sparse_budget=budget, seed=7)


def _oracle_layers(m):
    return [dict(ln1=l.input_layernorm_weight.cpu(), ln2=l.post_attention_layernorm_weight.cpu(),
                 eps=l.input_layernorm_variance_epsilon, wqkv=l.wqkv.cpu(), bqkv=None if l.bqkv is None else l.bqkv.cpu(),
                 wo=l.wo.cpu(), w_gate_up=l.gate_up_proj.cpu(), w_down=l.down_proj.cpu()) for l in m.layers]


@pytest.mark.parametrize("glm", [False, True])
def test_prefill_logits_and_state_match_the_oracle(cuda, glm):
    B, L = 1, 4096 + 256
    m = _model(cuda, glm, B, L)
    ids = torch.randint(0, m.vocab_size, (B, L), generator=torch.Generator().manual_seed(1))
    seen = {}
    real = m.kv_cache.prefill_kv_cache

    def spy(v, layer_idx, k_roped, last_q=None, **kw):            # what the prompt path hands to the cache, per layer
        outl = torch.zeros(B, m.num_key_value_heads, m.kv_cache.outlier_chunk, dtype=torch.int64, device=cuda)
        real(v, layer_idx, k_roped, last_q, outliers_out=outl)
        seen[layer_idx] = (k_roped.cpu(), v.cpu(), last_q.cpu(), outl.cpu())

    m.kv_cache.prefill_kv_cache = spy
    pre, real_svd = {}, m.kv_cache.get_svd

    def spy_svd(k, layer_idx):                                    # the pre-RoPE keys handed to get_svd: [bsz, H, L, D]
        pre[layer_idx] = k.cpu()
        return real_svd(k, layer_idx=layer_idx)

    m.kv_cache.get_svd = spy_svd
    logits = m.prefill(ids.to(cuda))
    torch.cuda.synchronize()
    assert logits.shape == (B, 1, m.vocab_size) and logits.dtype == torch.float32
    assert m.kv_cache.get_kv_len() == L
    want, per = DO.prefill_logits(ids, m.embed_tokens.cpu(), _oracle_layers(m), m.norm_weight.cpu(),
                                  m.norm_variance_epsilon, m.lm_head.cpu(), m.num_heads, m.num_key_value_heads,
                                  m.cos_sin_cache.cpu(), 2 if glm else 1)
    # layer 0 sees identical inputs on both sides: the library GEMM and the oracle's fp32 GEMM differ by at most one
    # bf16 ulp, the rotary embedding of a 1-ulp different input by the RoPE-propagated bound; V within one ulp
    k0, v0, q0, _ = seen[0]
    # (abs_floor: the normalised input itself differs by one ulp in a few elements, 2^-8 x |w| ~ 1e-4 per such term)
    assert_ulp_bf16(v0, per[0]["v"], ulps=1, frac_exact=0.9, abs_floor=2e-3, what="layer 0 V")
    assert_close_bf16(k0, per[0]["k_roped"], max_abs=4e-2, what="layer 0 post-RoPE K")
    assert_close_bf16(q0[:, :, 0], per[0]["q_roped"][:, :, -1], max_abs=4e-2, what="layer 0 last query")
    scale = max(1.0, float(want.abs().max()))
    assert_close_bf16(logits.cpu() / scale, want / scale, what="prompt logits")
    # the decode state of every layer is what the oracle builds from the SAME keys / values / last query
    c = m.kv_cache
    for l in range(m.num_layers):
        k_r, v_, q_, outl = seen[l]
        st = O.prefill_state(k_r, v_, q_[:, :, 0], budget=c.sparse_budget, outliers_override=outl,
                             selected_override=c.position_ids[l].cpu())
        assert (st["sparse_start"], st["sparse_end"]) == (c.sparse_start, c.sparse_end)
        n_chunks = L // 8
        assert torch.equal(c.v_cache_cpu[l][:, :, :n_chunks].cpu(), st["v_store"]), f"layer {l}: V store differs"
        assert torch.equal(c.k_landmark_idx[l].cpu(), st["rest_idx"]), f"layer {l}: landmark index differs"
        assert torch.equal(c.k_landmark[l].cpu(), st["landmarks"]), f"layer {l}: landmarks differ"
        assert torch.equal(c.k_cache_buffer[l].cpu(), st["k_buf"]), f"layer {l}: K buffer differs"
        assert torch.equal(c.v_cache_buffer[l].cpu(), st["v_buf"]), f"layer {l}: V buffer differs"
        # low-rank factors: the product U.SV against the oracle's truncated SVD of the SAME pre-RoPE keys (the factors
        # themselves are sign-ambiguous)
        kp = pre[l].transpose(1, 2).reshape(B, L, -1)
        Uo, SVo = O.low_rank_factors(kp, c.rank, m.num_key_value_heads)
        rec = torch.einsum("blr,bhdr->blhd", c.U[l].float().cpu(), c.SV[l].float().cpu())
        rec_o = torch.einsum("blr,bhdr->blhd", Uo.float(), SVo.float())
        rel = float((rec - rec_o).norm() / rec_o.norm())
        assert rel < 2e-2, f"layer {l}: rank-{c.rank} reconstruction differs from the reference SVD's by {rel:.3g}"
    assert_ulp_bf16(pre[0], per[0]["k_pre"], ulps=1, frac_exact=0.9, abs_floor=2e-3, what="layer 0 pre-RoPE K")
    assert not bool((c.position_ids == -1).any())


def test_generate_from_prompt_decodes_and_is_reproducible(cuda):
    """batch_prefill in sub-batches -> first token -> H2D -> the captured decode loop (base.py:573-655): two runs from
    the same prompt give the same ids; the first token is the argmax of the prompt logits; the cache advances."""
    B, L, gen = 2, 4096 + 64, 6
    ids = torch.randint(0, 2048, (B, L), generator=torch.Generator().manual_seed(2))
    outs = []
    for _ in range(2):
        m = _model(cuda, False, B, L)
        out = m.generate_from_prompt(ids.to(cuda), gen_len=gen)
        torch.cuda.synchronize()
        assert out.shape == (B, gen + 1)
        assert m.kv_cache.get_kv_len() == L + gen
        outs.append(out.cpu())
    assert torch.equal(outs[0], outs[1])
    m2 = _model(cuda, False, B, L)
    first = m2.batch_prefill(ids.to(cuda), sub_batch=1)[:, -1, :].argmax(dim=-1)
    assert torch.equal(first.cpu(), outs[0][:, 0])
