"""Restatement of the dense remainder of the reference's decode layer.

TEST INFRASTRUCTURE ONLY.  Nothing under ``shadowkv_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of the benches use it, and only as the checker.

What is restated (paths relative to the reference checkout):

* ``models/llama.py:354-385``  ``pre_attention_compute``: rmsnorm -> ``F.linear(h, wqkv)`` -> split
* ``models/llama.py:283-301``  ``apply_rotary_pos_emb``: ``vllm._custom_ops.rotary_embedding(pos, q, k, 128, cos_sin, True)``
* ``models/glm.py:312-357, 406-428``  the same with ``bias=bqkv`` and ``is_neox=False`` over a 64-wide table
* ``models/llama.py:387-427``  ``post_attention_compute``: o_proj, residual, rmsnorm, gate_up, ``silu_and_mul``, down, residual
* ``models/base.py:150-159``   final rmsnorm, ``F.linear(h, lm_head).float()``
* ``models/tensor_op.py:291-293``  ``sample_token`` with temperature 0

Third-party arithmetic that is NOT under /root/reference and is restated from its published
source, with the rounding points written out:

* flashinfer ``rmsnorm`` (``models/tensor_op.py:29,34-39``; wheel of ``README.md:54``): fp32 sum of squares,
  ``r = rsqrt(sum / d + eps)``, ``out = bf16(float(x) * r * float(w))``.
* torch ``F.linear`` on bf16 (cuBLAS): fp32 accumulation (order unspecified), optional bias added in fp32, one rounding
  to bf16.
* vllm ``rotary_embedding`` (csrc/pos_encoding_kernels.cu; the reference pins vllm 0.5.3.post1,
  ``requirements.txt:50``): ``x' = x*cos - y*sin``, ``y' = y*cos + x*sin``; neox pairs (j, j+rot/2), otherwise
  (2j, 2j+1); ``cos_sin_cache`` row = [cos(rot/2) | sin(rot/2)].  The vllm build in this image (0.22) evaluates the
  expression in fp32 and rounds ONCE to bf16 -- established bit for bit by the golden vectors -- and that is the
  default here and what the CUDA epilogue does.  The 0.5.3 source writes the same expression on ``scalar_t`` =
  c10::BFloat16 values, whose operators round every product and the sum to bf16; that variant cannot be executed
  here and is kept as ``per_op_rounding=True`` (it differs from the default by at most one rounding of each product,
  i.e. 2^-8 of the pair norm, inside the tolerance the parity tests state).
* vllm ``silu_and_mul`` (csrc/activation_kernels.cu): ``bf16(x / (1 + expf(-x)))`` computed in fp32, then a bf16
  multiply with the up half.

Pinning status: the reference holds no test or golden vector for any of these stages.  They are pinned against outputs
of the libraries themselves (the image's flashinfer / vllm builds and torch's cuBLAS ``F.linear``) run on a B200:
``tests/golden/dense_golden.npz`` made by ``tests/golden/make_dense_golden.py`` and checked in
``tests/test_dense_oracle.py``.
"""
from __future__ import annotations

import math

import torch

BF16 = torch.bfloat16


def _bf(x: torch.Tensor) -> torch.Tensor:
    """Round an fp32 tensor to bf16 and return it as fp32 again."""
    return x.to(BF16).float()


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """flashinfer rmsnorm as called by models/tensor_op.py:34-39.  x [..., d] bf16, w [d] bf16 -> bf16."""
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).sum(dim=-1, keepdim=True) / x.shape[-1] + eps)
    return (xf * r * w.float()).to(BF16)


def linear(x: torch.Tensor, W: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """F.linear on bf16 operands (llama.py:378, 408, 415, 424; glm.py:344): fp32 accumulate, round once."""
    y = x.float() @ W.float().t()
    if bias is not None:
        y = y + bias.float()
    return y.to(BF16)


def rotary_vllm(x: torch.Tensor, cos_sin_row: torch.Tensor, neox: bool, per_op_rounding: bool = False) -> torch.Tensor:
    """vllm rotary_embedding for one position.  x [..., heads, 128] bf16; cos_sin_row [rot] bf16 = [cos | sin].
    llama.py:296 (neox, rot 128); glm.py:421 (interleaved, rot 64: dims 64..127 pass through).
    per_op_rounding: round each product and the sum (c10::BFloat16 operator semantics, see the module docstring)."""
    rot = cos_sin_row.shape[-1]
    half = rot // 2
    c, s = cos_sin_row[:half].float(), cos_sin_row[half:].float()
    xf = x.float()
    out = xf.clone()
    r = _bf if per_op_rounding else (lambda t: t)
    if neox:
        a, b = xf[..., :half], xf[..., half:rot]
        out[..., :half] = r(a * c) - r(b * s)
        out[..., half:rot] = r(b * c) + r(a * s)
    else:
        a, b = xf[..., 0:rot:2], xf[..., 1:rot:2]
        out[..., 0:rot:2] = r(a * c) - r(b * s)
        out[..., 1:rot:2] = r(b * c) + r(a * s)
    return out.to(BF16)


def silu_and_mul(gu: torch.Tensor) -> torch.Tensor:
    """vllm silu_and_mul (llama.py:416-421): gu [..., 2F] bf16 = [gate | up] -> [..., F] bf16."""
    F_ = gu.shape[-1] // 2
    g, u = gu[..., :F_].float(), gu[..., F_:].float()
    return (_bf(g / (1.0 + torch.exp(-g))) * u).to(BF16)


def add_bf16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch `a + b` on bf16 tensors (llama.py:411, 426): fp32 add, one rounding."""
    return (a.float() + b.float()).to(BF16)


def pre_attention(x, ln_w, eps, wqkv, bqkv, q_heads, kv_heads, cos_sin, pos, rope_mode):
    """llama.py:354-385 + 283-301 (glm.py:312-357 + 406-428) for one decode token per sequence.
    x [B, hidden] bf16 -> q [B, q_heads, 128], k, v [B, kv_heads, 128] bf16, q and k post-RoPE.
    rope_mode: 0 none, 1 neox (Llama / Yi), 2 GLM."""
    B = x.shape[0]
    qkv = linear(rmsnorm(x, ln_w, eps), wqkv, bqkv)
    q, k, v = qkv.split([q_heads * 128, kv_heads * 128, kv_heads * 128], dim=-1)
    q, k, v = q.reshape(B, q_heads, 128), k.reshape(B, kv_heads, 128), v.reshape(B, kv_heads, 128)
    if rope_mode:
        q = rotary_vllm(q, cos_sin[pos], neox=rope_mode == 1)
        k = rotary_vllm(k, cos_sin[pos], neox=rope_mode == 1)
    return q, k, v


def post_attention(attn, resid, wo, ln_w, eps, w_gate_up, w_down):
    """llama.py:387-427.  attn [B, q_heads*128], resid [B, hidden] bf16 -> hidden states [B, hidden] bf16."""
    h = add_bf16(resid, linear(attn, wo))
    act = silu_and_mul(linear(rmsnorm(h, ln_w, eps), w_gate_up))
    return add_bf16(h, linear(act, w_down))


def lm_head(x, norm_w, eps, w_head):
    """base.py:150-159: final rmsnorm, lm_head, `.float()`."""
    return linear(rmsnorm(x, norm_w, eps), w_head).float()


def greedy(logits: torch.Tensor) -> torch.Tensor:
    """tensor_op.py:291-293 with temperature 0."""
    return logits.argmax(dim=-1)


# ---------------------------------------------------------------------------------------------------------------------
# Prompt (prefill) path of the dense stack: models/base.py:131-159 (inference), 296-318 (layer_compute, prefill branch),
# models/llama.py:354-427.  Same operators as above applied to [B, L, hidden]; the attention of base.py:313
# (flash_attn_with_kvcache, causal) is restated as a plain fp32 softmax(q k^T / sqrt(128)) v rounded once to bf16.
# ---------------------------------------------------------------------------------------------------------------------
def rotary_positions(x: torch.Tensor, cos_sin_rows: torch.Tensor, neox: bool) -> torch.Tensor:
    """vllm rotary_embedding for x [B, heads, L, 128] bf16 at the L positions whose table rows are cos_sin_rows [L, rot]
    (fp32 expression, one rounding; see rotary_vllm)."""
    rot = cos_sin_rows.shape[-1]
    half = rot // 2
    c, s = cos_sin_rows[:, :half].float(), cos_sin_rows[:, half:].float()
    xf = x.float()
    out = xf.clone()
    if neox:
        a, b = xf[..., :half], xf[..., half:rot]
        out[..., :half] = a * c - b * s
        out[..., half:rot] = b * c + a * s
    else:
        a, b = xf[..., 0:rot:2], xf[..., 1:rot:2]
        out[..., 0:rot:2] = a * c - b * s
        out[..., 1:rot:2] = b * c + a * s
    return out.to(BF16)


def causal_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """q [B, Hq, L, 128], k / v [B, H, L, 128] bf16 -> [B, L, Hq * 128] bf16 (query head j uses kv head j // (Hq / H))."""
    B, Hq, L, D = q.shape
    g = Hq // k.shape[1]
    kf, vf = k.float().repeat_interleave(g, dim=1), v.float().repeat_interleave(g, dim=1)
    mask = torch.ones(L, L, dtype=torch.bool).tril()
    out = torch.empty(B, Hq, L, D)
    for b in range(B):
        for h in range(Hq):
            sc = (q[b, h].float() @ kf[b, h].T) / math.sqrt(D)
            sc = sc.masked_fill(~mask, float("-inf"))
            out[b, h] = torch.softmax(sc, dim=-1) @ vf[b, h]
    return out.to(BF16).transpose(1, 2).reshape(B, L, Hq * D)


def prefill_layer(x, lw, q_heads, kv_heads, cos_sin_rows, rope_mode):
    """One decoder layer over a prompt.  x [B, L, hidden] bf16; lw: dict with ln1, ln2, eps, wqkv, bqkv, wo, w_gate_up,
    w_down.  Returns (hidden states [B, L, hidden], dict(k_pre [B,H,L,128], k_roped, v, q_roped))."""
    B, L, hid = x.shape
    qkv = linear(rmsnorm(x, lw["ln1"], lw["eps"]), lw["wqkv"], lw.get("bqkv")).view(B, L, q_heads + 2 * kv_heads, 128)
    q, k, v = (t.transpose(1, 2) for t in qkv.split([q_heads, kv_heads, kv_heads], dim=2))
    qr, kr = q, k
    if rope_mode:
        qr = rotary_positions(q, cos_sin_rows, neox=rope_mode == 1)
        kr = rotary_positions(k, cos_sin_rows, neox=rope_mode == 1)
    attn = causal_attention(qr, kr, v)
    h = add_bf16(x, linear(attn, lw["wo"]))
    act = silu_and_mul(linear(rmsnorm(h, lw["ln2"], lw["eps"]), lw["w_gate_up"]))
    return add_bf16(h, linear(act, lw["w_down"])), dict(k_pre=k, k_roped=kr, v=v, q_roped=qr)


def prefill_logits(ids, embed, layers, norm_w, eps, w_head, q_heads, kv_heads, cos_sin, rope_mode):
    """base.py:131-159 for a prompt: ids [B, L] -> (logits of the last position [B, 1, vocab] float32, per-layer dicts of
    prefill_layer)."""
    x = embed[ids]
    rows = cos_sin[: ids.shape[1]]
    per_layer = []
    for lw in layers:
        x, d = prefill_layer(x, lw, q_heads, kv_heads, rows, rope_mode)
        per_layer.append(d)
    return lm_head(x[:, -1:, :], norm_w, eps, w_head), per_layer
