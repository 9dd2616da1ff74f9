"""Mirror of the three hot-path wrappers of reference models/tensor_op.py:154-238 (same names, same
arguments).  They marshal strides and call the drop-in extension module ``shadowkv_b200.shadowkv``."""
from __future__ import annotations

import torch

from . import shadowkv


def apply_rotary_pos_emb_cuda(x, cos_sin, position_ids):
    """reference models/tensor_op.py:154-169"""
    batch_size, heads, seq_len, embed_dim = x.shape
    output = torch.empty_like(x)
    shadowkv.apply_rotary_pos_emb_new(
        x, cos_sin, position_ids, output, int(batch_size), int(heads), int(seq_len), int(embed_dim),
        int(x.stride(0)), int(x.stride(1)), int(x.stride(2)), int(x.stride(3)), int(cos_sin.stride(0)),
        int(position_ids.stride(0)), int(position_ids.stride(1)), int(position_ids.stride(2)), int(embed_dim // 2))
    return output


def apply_rotary_pos_emb_cuda_push_cache(x, cos_sin, position_ids, chunk_size, cache, sparse_start, sparse_end, cnts):
    """reference models/tensor_op.py:171-199"""
    batch_size, heads, seq_len, embed_dim = x.shape
    if cos_sin.shape[-1] == 128:
        fn = shadowkv.apply_rotary_pos_emb_push_cache_opt
    elif cos_sin.shape[-1] == 64:
        fn = shadowkv.apply_rotary_pos_emb_push_cache_opt_glm
    else:
        raise ValueError(f"Invalid cos_sin shape {cos_sin.shape}")
    fn(x, cos_sin, position_ids, cache, cnts, int(batch_size), int(heads), int(seq_len), int(embed_dim),
       int(x.stride(0)), int(x.stride(1)), int(x.stride(2)), int(x.stride(3)), int(cos_sin.stride(0)),
       int(position_ids.stride(0)), int(position_ids.stride(1)), int(position_ids.stride(2)),
       int(cache.stride(0)), int(cache.stride(1)), int(cache.stride(2)), int(sparse_start), int(sparse_end),
       int(embed_dim // 2), int(chunk_size))
    return cache


def batch_gather_gemm_rotary_pos_emb_cuda(a, b, cos_sin, position_ids, output, chunk_size, cache, sparse_start,
                                          sparse_end, cnts):
    """reference models/tensor_op.py:201-238: K rebuild into `output` (bf16, pre-RoPE), then RoPE + push."""
    batch_size, seq_len, rank = a.shape
    _, heads, head_dim, _ = b.shape
    max_seq_len, _ = cos_sin.shape
    _, _, num_chunks = position_ids.shape
    sparse_budget = num_chunks * chunk_size
    position_ids = position_ids.to(torch.int32).contiguous()
    cs = cos_sin.contiguous()
    shadowkv.batch_gather_gemm(a.contiguous(), b.contiguous(), cs, cs, position_ids, output, batch_size, heads,
                               seq_len, head_dim, rank, sparse_budget, max_seq_len, chunk_size, cnts)
    return apply_rotary_pos_emb_cuda_push_cache(output, cos_sin, position_ids, chunk_size, cache, sparse_start,
                                                sparse_end, cnts)


# ---- token sampling (reference models/tensor_op.py:242-297) ---------------------------------------------------------
# Off the hot path: one [batch, vocab] tensor per generated token.  Greedy decoding (temperature 0, the setting of the
# reference's throughput runs, test/e2e.py) is the CUDA `skv_argmax` kernel; the stochastic branch keeps the reference's
# torch operator sequence, so a given torch generator state yields the same tokens.
def top_k_top_p_filter(logits: torch.Tensor, top_k: int = 0, top_p: float = 0.0):
    """tensor_op.py:243-264: in-place -inf masking of everything below the k-th logit / outside the nucleus."""
    if top_k > 0:
        kth = torch.topk(logits, min(top_k, logits.size(-1)))[0]
        logits[logits < kth[:, [-1]]] = float("-inf")
    if top_p > 0.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative_probs = torch.cumsum(torch.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative_probs > top_p
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = 0
        logits[remove.scatter(1, sorted_indices, remove)] = float("-inf")
    return logits


def norm_logits(logits: torch.Tensor, temperature=0.6, top_k=-1, top_p=0.9) -> torch.Tensor:
    """tensor_op.py:266-284."""
    assert logits.dim() == 2
    if temperature != 1.0:
        logits = logits / temperature
    return torch.softmax(top_k_top_p_filter(logits, top_k=top_k, top_p=top_p), dim=-1)


def sample(probs: torch.Tensor, num_samples=1):
    """tensor_op.py:287-289."""
    return torch.multinomial(probs, num_samples=num_samples, replacement=True)


def sample_token(logits: torch.Tensor, temperature=0, top_k=50, top_p=0.9):
    """tensor_op.py:291-297: [batch, vocab] -> [batch, 1] token ids."""
    if temperature == 0.0:
        if logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous():
            from . import dense
            return dense.argmax(logits).unsqueeze(-1)
        return logits.argmax(dim=-1, keepdim=True)
    return sample(norm_logits(logits, temperature=temperature, top_p=top_p, top_k=top_k))
