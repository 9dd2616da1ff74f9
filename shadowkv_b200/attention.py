"""`flash_attn_with_kvcache` replacement for the decode call at reference models/base.py:341.

Only the form the reference uses on this path is supported: one query token per sequence, no new k/v
to append, no rotary, `causal=True` (a no-op for a single query at the end of the cache)."""
from __future__ import annotations

import math

import torch

from ._lib import check, lib

_workspaces: dict = {}


def _workspace(device, batch, kv_heads, groups):
    key = (device, batch, kv_heads, groups)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(lib.skv_attention_workspace_bytes(batch, kv_heads, groups, 128, 0), dtype=torch.uint8,
                         device=device)
        _workspaces[key] = ws
    return ws


def flash_attn_with_kvcache(q, k_cache, v_cache, causal=True, softmax_scale=None, num_splits=0, **unsupported):
    """q [B,1,Hq,D] bf16; k_cache, v_cache [B,L',H,D] bf16 (any batch/head/row strides, D contiguous)
    -> [B,1,Hq,D] bf16."""
    for name, val in unsupported.items():
        if val is not None:
            raise NotImplementedError(f"flash_attn_with_kvcache({name}=...) is not on the ShadowKV decode path")
    B, qlen, Hq, D = q.shape
    if qlen != 1 or D != 128:
        raise NotImplementedError("decode attention handles one query token with head_dim 128")
    if q.dtype != torch.bfloat16 or k_cache.dtype != torch.bfloat16 or v_cache.dtype != torch.bfloat16:
        raise TypeError("bf16 only")
    _, Lp, H, _ = k_cache.shape
    if k_cache.stride(3) != 1 or v_cache.stride() != k_cache.stride():
        raise ValueError("k_cache / v_cache must share strides with contiguous head_dim")
    G = Hq // H
    qc = q.contiguous()
    out = torch.empty(B, 1, Hq, D, dtype=q.dtype, device=q.device)
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    check(lib.skv_attention_decode(qc.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), out.data_ptr(), B, H, G, D,
                                   Lp, k_cache.stride(0), k_cache.stride(2), k_cache.stride(1), float(scale),
                                   int(num_splits), _workspace(q.device, B, H, G).data_ptr(), 0,
                                   torch.cuda.current_stream().cuda_stream), "attention_decode")
    return out
