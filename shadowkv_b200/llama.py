"""Host-side mirror of the reference's model classes for the DECODE step (reference models/llama.py ``LlamaLayer`` /
``Llama``, models/glm.py ``GLMLayer`` / ``GLM``, token loop models/base.py:131-159, 296-349, 573-655): the same method
names and argument meaning, every operator landing in the kernels of this repo (dense.py -> Part 4 of the C ABI,
kv_cache.py -> Parts 1-2).  One class serves Llama-3 / 3.1, Yi (same architecture) and GLM-4 (QKV bias, interleaved
rotary embedding over 64 dims); the differences are data (``ModelShape.glm_rope``), not code paths.

There is no hub access here, so ``init_parameters`` draws random weights N(0, 0.02) of the named architecture (the
reference loads them from ``LlamaForCausalLM.from_pretrained``, llama.py:310-352).

Prefill (base.py:163-187, 296-318, 500-550): ``prefill`` / ``batch_prefill`` run the prompt
through the dense stack with LIBRARY kernels, as the reference does -- cuBLAS ``F.linear``, torch scaled-dot-product
attention (flash backend) in place of ``flash_attn_with_kvcache`` -- and build the decode state of every layer through
``ShadowKVCache_CPU.get_svd`` / ``prefill_kv_cache`` (this repo's CUDA kernels, csrc/prefill.cu).  It is one-off work
per prompt, outside the decode step this repo is about; nothing of it runs when a token is decoded.
"""
from __future__ import annotations

import torch

from . import dense, synth
from .kv_cache import ShadowKVCache_CPU

BF16 = torch.bfloat16


def rotary_embedding_rows(x: torch.Tensor, cos_sin_rows: torch.Tensor, glm: bool) -> torch.Tensor:
    """vllm ``rotary_embedding`` (reference llama.py:296, glm.py:421) as torch ops, for the prompt path: x [..., L, 128]
    bf16 at the L positions whose table rows are cos_sin_rows [L, rot] = [cos | sin]; neox pairs (j, j + 64) over 128 dims,
    GLM pairs (2j, 2j + 1) over the first 64 (dims 64..127 pass through).  The expression is evaluated in fp32 and rounded
    once, like the vllm build of this image (oracle/dense_oracle.py ``rotary_vllm``, pinned to its golden vectors)."""
    half = cos_sin_rows.shape[-1] // 2
    c, s_ = cos_sin_rows[:, :half].float(), cos_sin_rows[:, half:].float()
    xf = x.float()
    if not glm:
        a, b = xf[..., :half], xf[..., half:]
        return torch.cat([a * c - b * s_, b * c + a * s_], dim=-1).to(BF16)
    a, b = xf[..., 0:2 * half:2], xf[..., 1:2 * half:2]
    o = torch.stack([a * c - b * s_, b * c + a * s_], dim=-1).flatten(-2)
    return torch.cat([o, xf[..., 2 * half:]], dim=-1).to(BF16)


class LlamaLayer:
    """Weights of one decoder layer, named as in reference models/llama.py:62-123 (``GLMLayer``: glm.py:79-131, whose
    fused gate/up matrix is called ``up_proj`` there and whose QKV projection has a bias ``bqkv``)."""

    def __init__(self, layer_idx: int) -> None:
        self.layer_idx = layer_idx
        self.wqkv = self.wo = self.gate_up_proj = self.down_proj = None
        self.bqkv = None
        self.input_layernorm_weight = self.post_attention_layernorm_weight = None
        self.input_layernorm_variance_epsilon = self.post_attention_layernorm_variance_epsilon = 1e-5
        self.q_size = self.kv_size = 0

    def init_parameters(self, shape: synth.ModelShape, generator: torch.Generator, device, std: float = 0.02):
        hid, ffn = shape.hidden_size, shape.intermediate_size
        self.q_size = shape.num_attention_heads * 128
        self.kv_size = shape.num_key_value_heads * 128

        def w(*s):
            return (torch.randn(*s, generator=generator, device=device) * std).to(BF16)

        self.wqkv = w(self.q_size + 2 * self.kv_size, hid)              # cat(q_proj, k_proj, v_proj), llama.py:95-99
        self.wo = w(hid, self.q_size)
        self.gate_up_proj = w(2 * ffn, hid)                             # cat(gate_proj, up_proj), llama.py:104
        self.down_proj = w(hid, ffn)
        self.bqkv = w(self.q_size + 2 * self.kv_size) if shape.glm_rope else None     # glm.py:118
        self.input_layernorm_weight = (1.0 + 0.1 * torch.randn(hid, generator=generator, device=device)).to(BF16)
        self.post_attention_layernorm_weight = (1.0 + 0.1 * torch.randn(hid, generator=generator, device=device)).to(BF16)
        return self


class Llama:
    """Decode-side mirror of reference ``Llama`` (models/llama.py:126-427) + ``LLM`` (models/base.py)."""

    def __init__(self, model: str | synth.ModelShape = "llama-3.1-8b", batch_size: int = 1, max_length: int = 124928,
                 device: str = "cuda:0", sparse_budget: int = 2048, rank: int = 160, chunk_size: int = 8,
                 num_layers: int | None = None, seed: int = 0, kv_cache: ShadowKVCache_CPU | None = None,
                 hbm_store_ways: int = 0, l2_prefetch_mb: float = 96.0, next_prefetch_mb: float = 32.0,
                 fused_finish: bool = True, split_qkv: bool = True) -> None:
        """l2_prefetch_mb: bytes of the layer's o_proj / gate_up weights pulled into L2 (126 MB on B200) on a side
        stream while the sparse-KV step of the same layer waits on its PCIe fetch (0 disables).
        next_prefetch_mb: head of the next linear's weights each dense kernel pulls into L2 during its own tail.
        fused_finish: o_proj / down_proj leave partial planes and the following norm kernel finishes them
        (dense.linear_partial + dense.finish_rmsnorm) instead of linear(residual=) + rmsnorm; same bits.
        split_qkv: only the q projection runs on the critical path; the k / v projection of the layer runs on a side
        stream in the PCIe window of the sparse-KV step and appends its rows in place."""
        self.shape = synth.MODELS[model] if isinstance(model, str) else model
        self.device = torch.device(device)
        self.batch_size, self.max_length = batch_size, max_length
        self.num_layers = num_layers or self.shape.num_hidden_layers
        self.hidden_size = self.shape.hidden_size
        self.num_heads = self.shape.num_attention_heads
        self.num_key_value_heads = self.shape.num_key_value_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.vocab_size = self.shape.vocab_size
        self.norm_variance_epsilon = 1e-5
        if self.head_dim != 128:
            raise ValueError("head_dim must be 128")
        if batch_size > dense.MAX_BATCH:
            raise ValueError(f"the dense decode kernels take at most {dense.MAX_BATCH} sequences per GPU")
        self.kv_cache = kv_cache if kv_cache is not None else ShadowKVCache_CPU(
            self.shape.config(self.num_layers), batch_size=batch_size, max_length=max_length, device=device,
            sparse_budget=sparse_budget, chunk_size=chunk_size, rank=rank, hbm_store_ways=hbm_store_ways)
        self.init_parameters(seed)
        B, hid, dev = batch_size, self.hidden_size, self.device
        # static activations of one decode token (graph-capturable)
        self.x = torch.zeros(B, hid, dtype=BF16, device=dev)            # residual stream
        self.h = torch.zeros(B, hid, dtype=BF16, device=dev)            # normalised input of the next linear
        self.q = torch.zeros(B, self.num_heads, 1, 128, dtype=BF16, device=dev)
        self.k = torch.zeros(B, self.num_key_value_heads, 1, 128, dtype=BF16, device=dev)
        self.v = torch.zeros(B, self.num_key_value_heads, 1, 128, dtype=BF16, device=dev)
        self.attn = torch.zeros(B, 1, self.num_heads, 128, dtype=BF16, device=dev)
        self.act = torch.zeros(B, self.shape.intermediate_size, dtype=BF16, device=dev)
        self.logits = torch.zeros(B, self.vocab_size, dtype=torch.float32, device=dev)
        self.tok = torch.zeros(B, dtype=torch.int64, device=dev)
        self.graph = None
        self.l2_prefetch_bytes = int(l2_prefetch_mb * 1e6)
        self.next_prefetch_bytes = int(next_prefetch_mb * 1e6) & ~15
        self.fused_finish = fused_finish
        self.split_qkv = split_qkv
        self.kv_stream = torch.cuda.Stream(device=dev)
        self.greedy = True                      # the token body ends in the argmax kernel (temperature 0)
        self.prefetch_stream = torch.cuda.Stream(device=dev)
        self._prepared = False
        self.stream = torch.cuda.Stream(device=dev)
        # kernels of libshadowkv_b200.so per token: embed+norm, per layer qkv + 5 (sparse-KV step) + o_proj + norm +
        # gate/up + down + norm, lm_head, argmax, counter advance
        self.launches_per_token = 1 + self.num_layers * (5 + 5 + (1 if self.kv_cache.hbm_store_ways else 0)) + 2

    # ---- parameters ----------------------------------------------------------------------------------------
    def init_parameters(self, seed: int = 0):
        g = torch.Generator(device=self.device).manual_seed(seed)
        sh, dev = self.shape, self.device
        self.embed_tokens = (torch.randn(sh.vocab_size, sh.hidden_size, generator=g, device=dev) * 0.02).to(BF16)
        self.lm_head = (torch.randn(sh.vocab_size, sh.hidden_size, generator=g, device=dev) * 0.02).to(BF16)
        self.norm_weight = (1.0 + 0.1 * torch.randn(sh.hidden_size, generator=g, device=dev)).to(BF16)
        self.cos_sin_cache = synth.cos_sin_table(self.max_length + 1024, theta=sh.rope_theta, glm=sh.glm_rope, device=dev)
        self.layers = [LlamaLayer(i).init_parameters(sh, g, dev) for i in range(self.num_layers)]

    def weight_bytes(self) -> int:
        """Bytes one decode token streams from the dense weights (every layer + lm_head; embedding rows excluded)."""
        per = sum(t.numel() * 2 for t in (self.layers[0].wqkv, self.layers[0].wo, self.layers[0].gate_up_proj,
                                          self.layers[0].down_proj))
        return per * self.num_layers + self.lm_head.numel() * 2

    # ---- reference-named pieces of layer_compute --------------------------------------------------------------
    def pre_attention_compute(self, hidden_states, buffer: LlamaLayer, num_heads: int, num_key_value_heads: int,
                              head_dim: int, position: int | None = None, pos_dev=None, normed=None, out=None):
        """llama.py:354-385 (+ ``apply_rotary_pos_emb``, :283-301, when ``position`` is given -- the reference applies
        it right after, base.py:320).  hidden_states [B, hidden] (or [B,1,hidden]); returns q [B,Hq,1,D], k, v
        [B,H,1,D].  ``normed``: the already normalised input (the whole-token runner fuses the norm into the
        previous step)."""
        x = hidden_states.reshape(-1, self.hidden_size) if hidden_states is not None else None
        h = normed if normed is not None else dense.rmsnorm(x, buffer.input_layernorm_weight,
                                                            buffer.input_layernorm_variance_epsilon)
        q_out, k_out, v_out = out if out is not None else (None, None, None)
        return dense.linear_qkv_rope(h, buffer.wqkv, num_heads, num_key_value_heads,
                                     self.cos_sin_cache if position is not None else None, position or 0,
                                     bias=buffer.bqkv, glm=self.shape.glm_rope, pos_dev=pos_dev,
                                     q_out=q_out, k_out=k_out, v_out=v_out, static_weights=True)

    def post_attention_compute(self, attn_output, residual, buffer: LlamaLayer, out=None):
        """llama.py:387-427: o_proj + residual, rmsnorm, gate/up, SiLU*mul, down + residual.  attn_output
        [B, 1, Hq, D] / [B, hidden]; residual [B, hidden]; returns the new hidden states [B, hidden] (``out`` may be
        ``residual``)."""
        B = residual.shape[0]
        x1 = dense.linear(attn_output.reshape(B, -1), buffer.wo, residual=residual, out=out, static_weights=True)
        h = dense.rmsnorm(x1, buffer.post_attention_layernorm_weight, buffer.post_attention_layernorm_variance_epsilon,
                          out=self.h if out is not None else None)
        act = dense.linear_silu_mul(h, buffer.gate_up_proj, out=self.act if out is not None else None, static_weights=True)
        return dense.linear(act, buffer.down_proj, residual=x1, out=x1 if out is not None else None, static_weights=True)

    def layer_compute(self, buffer: LlamaLayer, layer_idx: int, hidden_states, position_ids):
        """base.py:296-349, decode branch: hidden_states [B,1,hidden], position_ids [B,1] (all rows equal, as the
        reference's generate loop makes them, base.py:628-636)."""
        if hidden_states.shape[1] > 4 * 1024:                            # base.py:296: the prefill branch
            return self._prefill_layer(buffer, layer_idx, hidden_states, position_ids)
        if hidden_states.shape[1] != 1:
            raise NotImplementedError("multi-token continuation (1 < q_len <= 4096, base.py:189 prefill_cont) is not built")
        pos = int(position_ids.reshape(-1)[0])
        residual = hidden_states.reshape(-1, self.hidden_size).contiguous()
        q, k, v = self.pre_attention_compute(residual, buffer, self.num_heads, self.num_key_value_heads, self.head_dim,
                                             position=pos)
        attn = self.kv_cache.decode_layer(layer_idx, q, k, v, self.cos_sin_cache)
        return self.post_attention_compute(attn, residual, buffer).view(-1, 1, self.hidden_size)

    def inference(self, input_ids, position_ids):
        """base.py:131-159: [B,1] ids (decode token) or [B,L] ids with L > 4096 (prompt) -> logits [B,1,vocab] float32
        (for a prompt only the last position goes through the lm_head, base.py:153-155)."""
        hidden_states = self.embed_tokens[input_ids]                      # F.embedding, base.py:141
        for idx in range(self.num_layers):
            hidden_states = self.layer_compute(self.layers[idx], idx, hidden_states, position_ids)
        if hidden_states.shape[1] > 16:
            hidden_states = hidden_states[:, -1:, :]
        h = dense.rmsnorm(hidden_states.reshape(-1, self.hidden_size).contiguous(), self.norm_weight,
                          self.norm_variance_epsilon)
        return dense.linear_f32(h, self.lm_head, static_weights=True).view(-1, 1, self.vocab_size)

    # ---- prefill (section 8 f4): library kernels for the dense stack, this repo's kernels for the decode state ------
    def get_ctx(self, input_ids):
        """base.py:109-126: position ids of `input_ids` behind what the cache already holds."""
        past, n = self.kv_cache.get_kv_len(), input_ids.size(1)
        return torch.arange(past, past + n, device=self.device, dtype=torch.long).unsqueeze(0).repeat(input_ids.size(0), 1)

    def apply_rotary_pos_emb(self, query_states, key_states, position_ids):
        """llama.py:283-301 / glm.py:406-428 for [bsz, heads, L, 128] tensors, every row of position_ids equal."""
        rows = self.cos_sin_cache[position_ids[0]]                          # [L, rot] = [cos | sin]
        return (rotary_embedding_rows(query_states, rows, self.shape.glm_rope),
                rotary_embedding_rows(key_states, rows, self.shape.glm_rope))

    @staticmethod
    def _causal_attention(q, k, v):
        """Stands in for ``flash_attn_with_kvcache(q, k, v, causal=True)`` of base.py:313 on the prompt: torch SDPA with the
        memory-efficient backends only (the math backend would materialise an L x L score matrix per head).
        q [bsz,Hq,L,D], k / v [bsz,H,L,D] -> [bsz,L,Hq*D]."""
        import torch.nn.functional as F
        from torch.nn.attention import SDPBackend, sdpa_kernel
        with sdpa_kernel([SDPBackend.FLASH_ATTENTION, SDPBackend.CUDNN_ATTENTION, SDPBackend.EFFICIENT_ATTENTION]):
            try:
                o = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
            except RuntimeError:                                          # a backend without grouped-query support
                g = q.shape[1] // k.shape[1]
                o = F.scaled_dot_product_attention(q, k.repeat_interleave(g, dim=1), v.repeat_interleave(g, dim=1),
                                                   is_causal=True)
        return o.transpose(1, 2).reshape(q.shape[0], q.shape[2], -1)

    def _prefill_layer(self, buffer: LlamaLayer, layer_idx: int, hidden_states, position_ids, mlp_rows: int = 32768):
        """base.py:296-318 + llama.py:354-427 for a prompt [bsz, L, hidden] (L > 4096): QKV, low-rank factors of the
        pre-RoPE keys (get_svd), RoPE, the layer's decode state (prefill_kv_cache), causal attention, o_proj and the MLP
        (in blocks of `mlp_rows` tokens, which bounds the 2 x FFN-wide intermediate)."""
        import torch.nn.functional as F
        bsz, L, hid = hidden_states.shape
        Hq, H, D = self.num_heads, self.num_key_value_heads, self.head_dim
        x = hidden_states.reshape(bsz * L, hid)
        x = x if x.is_contiguous() else x.contiguous()
        h = dense.rmsnorm(x, buffer.input_layernorm_weight, buffer.input_layernorm_variance_epsilon)
        qkv = F.linear(h, buffer.wqkv, buffer.bqkv).view(bsz, L, Hq + 2 * H, D)
        del h
        q, k, v = (t.transpose(1, 2) for t in qkv.split([Hq, H, H], dim=2))      # [bsz, heads, L, D] views
        self.kv_cache.get_svd(k, layer_idx=layer_idx)
        q, k = self.apply_rotary_pos_emb(q, k, position_ids)
        v = v.contiguous()
        del qkv
        self.kv_cache.prefill_kv_cache(v, layer_idx, k, q[:, :, -1:])
        attn = self._causal_attention(q, k, v)
        del q, k, v
        x1 = x + F.linear(attn.reshape(bsz * L, Hq * D), buffer.wo)               # llama.py:408-411
        del attn
        ffn = self.shape.intermediate_size
        out = torch.empty_like(x1)
        for r0 in range(0, bsz * L, mlp_rows):
            xs = x1[r0:r0 + mlp_rows]
            gu = F.linear(dense.rmsnorm(xs, buffer.post_attention_layernorm_weight,
                                        buffer.post_attention_layernorm_variance_epsilon), buffer.gate_up_proj)
            act = F.silu(gu[:, :ffn]) * gu[:, ffn:]                               # vllm silu_and_mul, llama.py:416-421
            out[r0:r0 + mlp_rows] = xs + F.linear(act, buffer.down_proj)          # llama.py:424-426
        return out.view(bsz, L, hid)

    @torch.inference_mode()
    def prefill(self, input_ids):
        """base.py:163-187: clear the cache, run the prompt, return the last position's logits [bsz, 1, vocab]."""
        self.kv_cache.clear()
        logits = self.inference(input_ids=input_ids, position_ids=self.get_ctx(input_ids))
        assert self.kv_cache.get_kv_len() == input_ids.shape[-1], \
            f"KV length mismatch, got {self.kv_cache.get_kv_len()}, expected {input_ids.shape[-1]}"
        return logits

    @torch.inference_mode()
    def batch_prefill(self, input_ids, sub_batch: int | None = None):
        """base.py:500-550: the prompts of the whole batch in sub-batches (the reference takes 8 sequences at a time
        between 120K and 200K tokens and 4 otherwise; here the default is 1 above 64K tokens, which keeps the fp32
        key matrix of get_svd and the attention inputs of a sub-batch within a few GB)."""
        self.kv_cache.clear()
        bsz, L = input_ids.shape
        assert bsz == self.batch_size, f"batch_size mismatch, got {bsz}, expected {self.batch_size}"
        if L > self.max_length:
            raise ValueError(f"Input length must be less than {self.max_length}, but got {L}")
        T = sub_batch or (1 if L > 64 * 1024 else 4)
        logits = torch.zeros(bsz, 1, self.vocab_size, device=self.device, dtype=torch.float32)
        for b0 in range(0, bsz, T):
            ids = input_ids[b0:b0 + T]
            logits[b0:b0 + T].copy_(self.inference(input_ids=ids, position_ids=self.get_ctx(ids)))
        assert self.kv_cache.get_kv_len() == L, f"KV length mismatch, got {self.kv_cache.get_kv_len()}, expected {L}"
        return logits

    @torch.inference_mode()
    def generate_from_prompt(self, input_ids, gen_len: int = 256, temperature: float = 0.0, top_p: float = -1,
                             top_k: int = 50, benchmark: bool = False):
        """``LLM.batch_generate`` of base.py:573-655 from token ids: batch_prefill -> sample the first token -> H2D ->
        the decode loop (``batch_generate`` below).  Returns [B, 1 + gen_len] ids (first sampled token included)."""
        from .tensor_op import sample_token
        logits = self.batch_prefill(input_ids)
        next_token = sample_token(logits[:, -1, :], temperature=temperature, top_p=top_p, top_k=top_k)
        self.kv_cache.H2D()
        self.graph = None                                                 # the captured token belongs to the old state
        self._prepared = False
        res = self.batch_generate(next_token, gen_len=gen_len, temperature=temperature, top_p=top_p, top_k=top_k,
                                  benchmark=benchmark)
        ids = res[0] if benchmark else res
        ids = torch.cat([next_token.reshape(-1, 1).to(ids.device), ids], dim=1)
        return (ids, res[1]) if benchmark else ids

    # ---- one decode token with device-side counters: the body of the CUDA graph ---------------------------------
    def _pre_attention_step(self, l: int):
        """Fused QKV + RoPE of layer l from the already normalised self.h (device-side position counter).  With
        split_qkv only the q heads are projected here, on the critical path; k and v follow in `_after_select_hook`."""
        c, buf = self.kv_cache, self.layers[l]
        if not self.split_qkv:
            self.pre_attention_compute(None, buf, self.num_heads, self.num_key_value_heads, 128, position=c.prefill,
                                       pos_dev=c.gen_dev, normed=self.h, out=(self.q, self.k, self.v))
            return
        Hq = self.num_heads
        dense.linear_qkv_rope(self.h, buf.wqkv[:Hq * 128], Hq, self.num_key_value_heads, self.cos_sin_cache, c.prefill,
                              bias=buf.bqkv, glm=self.shape.glm_rope, pos_dev=c.gen_dev, q_out=self.q,
                              static_weights=True, heads=(0, Hq))

    def _project_kv_in_place(self, l: int):
        """k / v projection + RoPE(k) of layer l on `kv_stream`, appended in place to the layer's cache buffers
        (update_kv_cache, kv_cache.py:1227-1271).  Runs while the sparse-KV step waits on its PCIe fetch; the attention
        joins it through `_before_attention_hook`."""
        c, buf = self.kv_cache, self.layers[l]
        Hq, H = self.num_heads, self.num_key_value_heads
        cur, side = torch.cuda.current_stream(), self.kv_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            dense.linear_qkv_rope(self.h, buf.wqkv[Hq * 128:], Hq, H, self.cos_sin_cache, c.prefill, bias=buf.bqkv,
                                  glm=self.shape.glm_rope, pos_dev=c.gen_dev, static_weights=True, heads=(Hq, 2 * H),
                                  kv_cache=(c.k_cache_buffer[l], c.v_cache_buffer[l], c.sparse_end))

    def _after_select_hook(self, l: int):
        if not (self.l2_prefetch_bytes or self.split_qkv):
            return None

        def hook(l=l, b=self.layers[l]):
            if self.split_qkv:
                self._project_kv_in_place(l)
            if self.l2_prefetch_bytes:
                self._prefetch_post_attention_weights(b)
        return hook

    def _before_attention_hook(self):
        return (lambda: torch.cuda.current_stream().wait_stream(self.kv_stream)) if self.split_qkv else None

    def _post_attention_step(self, l: int):
        """o_proj + residual, norm, gate/up + SiLU*mul, down + residual, next norm -- in place on self.x / self.h.
        Each linear names the head of the NEXT linear's weights: they are pulled into L2 during its own tail."""
        L, buf, n = self.layers, self.layers[l], self.next_prefetch_bytes
        gu_done = max(0, self.l2_prefetch_bytes - buf.wo.numel() * 2)        # what the hot-path window already pulled
        nxt_w = L[l + 1].wqkv if l + 1 < len(L) else self.lm_head
        nxt_n = L[l + 1].input_layernorm_weight if l + 1 < len(L) else self.norm_weight
        if self.fused_finish:
            # the linear leaves fp32 partial planes; the norm kernel that follows anyway adds them, the residual and the
            # norm in one pass (no cross-CTA fix-up at the end of the linear); bit-identical to the branch below
            pl = dense.linear_partial(self.attn.view(self.batch_size, -1), buf.wo, static_weights=True,
                                      prefetch=(buf.gate_up_proj, gu_done, n) if n else None)
            dense.finish_rmsnorm(pl, buf.wo.shape[1], self.x, buf.post_attention_layernorm_weight,
                                 buf.post_attention_layernorm_variance_epsilon, h_out=self.h)
        else:
            dense.linear(self.attn.view(self.batch_size, -1), buf.wo, residual=self.x, out=self.x, static_weights=True,
                         prefetch=(buf.gate_up_proj, gu_done, n) if n else None)
            dense.rmsnorm(self.x, buf.post_attention_layernorm_weight, buf.post_attention_layernorm_variance_epsilon,
                          out=self.h)
        dense.linear_silu_mul(self.h, buf.gate_up_proj, out=self.act, static_weights=True,
                              prefetch=(buf.down_proj, 0, n) if n else None)
        if self.fused_finish:
            pl = dense.linear_partial(self.act, buf.down_proj, static_weights=True, prefetch=(nxt_w, 0, n) if n else None)
            dense.finish_rmsnorm(pl, buf.down_proj.shape[1], self.x, nxt_n, self.norm_variance_epsilon, h_out=self.h)
        else:
            dense.linear(self.act, buf.down_proj, residual=self.x, out=self.x, static_weights=True,
                         prefetch=(nxt_w, 0, n) if n else None)
            dense.rmsnorm(self.x, nxt_n, self.norm_variance_epsilon, out=self.h)

    def _head_step(self):
        n = self.next_prefetch_bytes
        dense.linear_f32(self.h, self.lm_head, out=self.logits, static_weights=True,
                         prefetch=(self.layers[0].wqkv, 0, n) if n else None)      # the next token starts there
        if self.greedy:
            dense.argmax(self.logits, out=self.tok)

    def _token_body(self):
        c, L = self.kv_cache, self.layers
        dense.embed_rmsnorm(self.tok, self.embed_tokens, L[0].input_layernorm_weight,
                            L[0].input_layernorm_variance_epsilon, self.x, self.h)
        for l in range(len(L)):
            self._pre_attention_step(l)
            kv = (None, None) if self.split_qkv else (self.k, self.v)
            c.decode_layer(l, self.q, kv[0], kv[1], self.cos_sin_cache, out=self.attn, device_counters=True,
                           after_select=self._after_select_hook(l), before_attention=self._before_attention_hook())
            self._post_attention_step(l)
        self._head_step()
        if self.l2_prefetch_bytes:
            torch.cuda.current_stream().wait_stream(self.prefetch_stream)
        c.join_side_streams()
        c.advance_device_counters()

    def _prefetch_post_attention_weights(self, buf: LlamaLayer):
        """Runs on `prefetch_stream`, ordered behind the selection kernel of the layer: o_proj first, then as much of
        the gate/up matrix as the budget allows.  A hint only: nothing waits for it except the end of the token."""
        cur, pf = torch.cuda.current_stream(), self.prefetch_stream
        pf.wait_stream(cur)
        with torch.cuda.stream(pf):
            left = self.l2_prefetch_bytes
            for w in (buf.wo, buf.gate_up_proj):
                n = min(left, w.numel() * 2)
                if n > 0:
                    dense.l2_prefetch(w, nbytes=n)
                left -= n

    def prepare(self):
        """Before the first token of the runner: order it behind whatever built the weights / cache state on other
        streams, and load the device-side counters from the host-side generation offset."""
        torch.cuda.synchronize(self.device)
        self.kv_cache.reset_device_counters()
        self._prepared = True

    def capture(self, warmup: int = 1, rewind: bool = False):
        """Warm up eagerly (real decode steps: they allocate scratch, encode tensor maps, set kernel attributes), then
        capture one token as a CUDA graph.  rewind=True puts the generation offset back afterwards, so the warm-up
        tokens leave only cache-state effects behind (which chunks sit in the sparse slots), never results."""
        c = self.kv_cache
        self.prepare()
        g0 = c.gen_offset
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self._token_body()
                c.note_replayed_steps(1)
        self.stream.synchronize()
        if rewind and c.gen_offset != g0:
            c.kv_offset -= c.gen_offset - g0
            c.gen_offset = g0
            c.reset_device_counters()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._token_body()
        return self

    def decode_token(self):
        """Enqueue one greedy decode token (input: ``self.tok``; output: ``self.tok`` / ``self.logits``)."""
        c = self.kv_cache
        if c.sparse_end + c.gen_offset + 1 > c.buf_len:
            raise RuntimeError(f"generation buffer full ({c.buf_len - c.sparse_end} tokens)")
        if not self._prepared:
            self.prepare()
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._token_body()
        c.note_replayed_steps(1)

    def batch_generate(self, next_token: torch.Tensor, gen_len: int = 256, temperature: float = 0.0, top_p: float = -1,
                       top_k: int = 50, benchmark: bool = False):
        """Decode loop of reference ``LLM.batch_generate`` (base.py:626-655) from the first sampled token on (the cache
        already holds the prompt: prefill is outside this path).  Returns [B, gen_len] ids, and tokens/s when
        ``benchmark``.  Temperature 0 keeps the argmax inside the captured graph; otherwise the graph ends at the
        logits and ``tensor_op.sample_token`` (the reference's operator sequence) draws the token."""
        import time
        from .tensor_op import sample_token
        greedy = temperature == 0.0
        if self.graph is None or greedy != self.greedy:
            self.greedy = greedy
            self.capture(warmup=1, rewind=True)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.tok.copy_(next_token.reshape(-1))
        out = []
        self.stream.synchronize()
        t0 = time.perf_counter()
        for _ in range(gen_len):
            self.decode_token()
            with torch.cuda.stream(self.stream):
                if not greedy:
                    self.tok.copy_(sample_token(self.logits.clone(), temperature=temperature, top_p=top_p, top_k=top_k)[:, 0])
                out.append(self.tok.clone())
        self.stream.synchronize()
        ids = torch.stack(out, dim=1)
        if benchmark:
            return ids, self.batch_size * gen_len / (time.perf_counter() - t0)
        return ids

    def generate(self, first_token: torch.Tensor, gen_len: int):
        """Greedy loop of base.py:626-640 (temperature 0): returns [B, gen_len] generated ids."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.tok.copy_(first_token.reshape(-1))
        out = []
        for _ in range(gen_len):
            self.decode_token()
            with torch.cuda.stream(self.stream):
                out.append(self.tok.clone())
        self.stream.synchronize()
        return torch.stack(out, dim=1)
