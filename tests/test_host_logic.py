"""Host-side geometry against the numbers the survey derived from the reference."""
from shadowkv_b200 import synth


def test_layout_cfg2_llama_122k():
    lay = synth.layout(124928, 2048, 8)
    assert lay.all_chunks == 15616
    assert lay.chunks == 15608
    assert lay.n_landmarks == 15560
    assert lay.prefill_local == 64
    assert (lay.sparse_start, lay.sparse_end, lay.buf_len) == (448, 2496, 2592)
    assert lay.select == 256


def test_layout_other_configs():
    assert synth.layout(524288).n_landmarks == 65480          # cfg3
    assert synth.layout(131072).n_landmarks == 16328          # cfg4
    assert synth.layout(204800).n_landmarks == 25544          # cfg5
    # budgets 1024 / 4096 of test/e2e.py:35-116
    assert synth.layout(61440, 1024).outlier_chunk == 24
    assert synth.layout(249856, 4096).select == 512


def test_cos_sin_tables():
    t = synth.cos_sin_table(64, 128)
    assert t.shape == (64, 128)
    assert float(t[0, :64].float().min()) == 1.0 and float(t[0, 64:].float().abs().max()) == 0.0
    g = synth.cos_sin_table(64, 128, glm=True)
    assert g.shape == (64, 64)


def test_model_shapes():
    m = synth.MODELS["llama-3.1-8b"]
    cfg = m.config()
    assert cfg.num_attention_heads // cfg.num_key_value_heads == 4
    assert cfg.hidden_size // cfg.num_attention_heads == 128
    assert synth.MODELS["glm-4-9b-1m"].glm_rope


def test_sampling_mirrors_follow_the_reference_semantics():
    """tensor_op.top_k_top_p_filter / norm_logits / sample_token (reference models/tensor_op.py:242-297) on CPU."""
    import torch
    from shadowkv_b200 import tensor_op as T
    logits = torch.tensor([[1.0, 3.0, 2.0, 0.5, -1.0], [0.0, 0.0, 5.0, 4.0, 4.5]])
    f = T.top_k_top_p_filter(logits.clone(), top_k=2)
    assert torch.isinf(f[0, [0, 3, 4]]).all() and torch.isfinite(f[0, [1, 2]]).all()
    assert torch.isinf(f[1, [0, 1, 3]]).all() and torch.isfinite(f[1, [2, 4]]).all()
    # nucleus: the token that crosses top_p is kept (shifted mask), everything after it is dropped
    p = T.top_k_top_p_filter(torch.tensor([[2.0, 1.0, 0.0, -1.0]]), top_p=0.7)
    probs = torch.softmax(torch.tensor([2.0, 1.0, 0.0, -1.0]), -1)
    assert probs[0] < 0.7 < probs[0] + probs[1]
    assert torch.isfinite(p[0, :2]).all() and torch.isinf(p[0, 2:]).all()
    pr = T.norm_logits(logits.clone(), temperature=0.5, top_k=3, top_p=-1)
    assert torch.allclose(pr.sum(-1), torch.ones(2)) and (pr[0, [3, 4]] == 0).all()
    assert T.sample_token(logits, temperature=0).tolist() == [[1], [2]]
    g = torch.manual_seed(0)
    tok = T.sample_token(logits.clone(), temperature=0.7, top_k=2, top_p=-1)
    assert tok.shape == (2, 1) and int(tok[0]) in (1, 2) and int(tok[1]) in (2, 4)


def test_hbm_store_sizing():
    """enable_hbm_store("auto"): ways from the free HBM (pure arithmetic, no GPU needed)."""
    from shadowkv_b200.kv_cache import hbm_store_ways_for
    per_way = 2048 + 4
    # Llama-3.1-8B, 122K, batch 1: 32 layers x 8 rows; 100 GB free, half of it for the store
    w = hbm_store_ways_for(100 << 30, 32, 8, 15616)
    assert w == 16384 and w * 32 * 8 * per_way <= 50 << 30          # capped: 16384 ways hold all 15616 chunks of a row
    # batch 32 of the same shape: the budget decides
    w = hbm_store_ways_for(100 << 30, 32, 8 * 32, 15616)
    assert w & (w - 1) == 0 and w * 32 * 256 * per_way <= 50 << 30 < 2 * w * 32 * 256 * per_way
    assert hbm_store_ways_for(1 << 20, 32, 8, 15616) == 0               # nothing fits
    assert hbm_store_ways_for(1 << 40, 1, 1, 100) == 128                # never fewer than the kernels' smallest table
