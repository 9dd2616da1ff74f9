"""The reference's decode step for ONE layer, call for call, on the reference's own CUDA path: the unmodified
extension built by oracle/build_ref.sh (oracle/_ref/kernels) + ``flash_attn_with_kvcache`` from the flash-attn wheel,
with the argument marshalling of reference models/kv_cache.py:983-1176, models/tensor_op.py:201-238 and the stream
choreography of models/base.py:320-341.

Checker-side helper (it executes oracle/_ref): used by tests/golden/make_step_golden.py to capture step-level golden
vectors on a B200.  None of this repo's kernels is launched here."""
from __future__ import annotations

import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference_extension():
    d = os.path.join(ROOT, "oracle", "_ref", "kernels")
    if d not in sys.path:
        sys.path.insert(0, d)
    import shadowkv as ref                              # TORCH_EXTENSION_NAME=shadowkv
    from flash_attn import flash_attn_with_kvcache      # models/base.py:38
    return ref, flash_attn_with_kvcache


class ReferenceLayer:
    """State of one layer laid out as ShadowKVCache_CPU keeps it (kv_cache.py:554-649, 694-696, 767-780), built from
    a ``synth.layer_state`` dict."""

    def __init__(self, st: dict, ctx: int, budget: int, dev):
        lay = st["layout"]
        self.lay, self.ctx, self.budget, self.dev = lay, ctx, budget, dev
        B, H, N, D = st["k_landmark"].shape
        self.B, self.H, self.N, self.D, self.S = B, H, N, D, lay.select
        self.v_cache_cpu = torch.zeros(B, H, ctx // 8, 8 * D, dtype=torch.bfloat16, pin_memory=True)
        self.v_cache_cpu[:, :, :st["v_chunks"].shape[2]].copy_(st["v_chunks"])
        self.k_buf, self.v_buf = st["k_buf"].to(dev).clone(), st["v_buf"].to(dev).clone()
        self.U, self.SV = st["U"].to(dev), st["SV"].to(dev)
        self.LM, self.LMI = st["k_landmark"].to(dev), st["k_landmark_idx"].to(dev)
        self.position_ids = st["resident"].to(dev).clone()
        self.R = self.U.shape[-1]
        self.offsets = torch.zeros(B * H * self.S, dtype=torch.int32, device=dev)
        self.cnts = torch.zeros(B * H, dtype=torch.int32, device=dev)
        self.signals = torch.zeros(B * H, dtype=torch.int32, device=dev)
        self.temp = torch.zeros(B, H, self.S, 8 * D, dtype=torch.bfloat16, device=dev)
        self.output = torch.zeros(B, H, budget, D, dtype=torch.bfloat16, device=dev)
        self.copy_stream = torch.cuda.Stream()
        self._scratch_G = None

    def step(self, ref, flash, q, kn, vn, gen: int, cs):
        """q [B,Hq,1,D], kn / vn [B,H,1,D] post-RoPE on the device; cs = cos_sin table on the device."""
        lay, B, H, D, S, N = self.lay, self.B, self.H, self.D, self.S, self.N
        G = q.shape[1] // H
        if self._scratch_G != G:
            self.gemm_o = torch.zeros(B, H, G, N, dtype=torch.bfloat16, device=self.dev)
            self.softmax_o = torch.zeros_like(self.gemm_o)
            T = (N + 255) // 256
            self.norm = torch.zeros(B * H, G, T, device=self.dev)
            self.ssum = torch.zeros(B * H, G, T, device=self.dev)
            self._scratch_G = G
        pos = lay.sparse_end + gen
        self.v_buf[:, :, pos:pos + 1].copy_(vn, non_blocking=True)                                  # kv_cache.py:1255-1265
        self.k_buf[:, :, pos:pos + 1].copy_(kn, non_blocking=True)
        ref.batch_gemm_softmax(q.contiguous(), self.LM.contiguous(), self.gemm_o, self.norm, self.ssum, self.softmax_o,
                               B * H, G, N, D, 1 / math.sqrt(128), 0)                               # :1006-1019
        chunk_attn, _ = torch.max(self.softmax_o.view(B, H, G, -1), dim=-2)                         # :1024-1027
        top = torch.topk(chunk_attn.view(B, H, -1), k=S, dim=-1).indices                            # :1031-1035
        selected = self.LMI.gather(dim=-1, index=top)                                               # :1039-1042
        ref.reorder_keys_and_compute_offsets(self.position_ids, selected, self.offsets, self.cnts, B, H, S)   # :1047-1055
        cur = torch.cuda.current_stream()
        kernel_offset, kernel_stride = lay.sparse_start * D, lay.buf_len * D
        max_ctx_chunks_len = (self.ctx // 8) * 8
        with torch.cuda.stream(self.copy_stream):                                                   # base.py:329-332
            self.copy_stream.wait_stream(cur)
            ref.gather_copy_with_offsets(self.v_cache_cpu, self.v_buf, self.temp, self.offsets, self.cnts, self.signals,
                                         B, H, int(max_ctx_chunks_len * D), int(self.budget * D), kernel_offset,
                                         kernel_stride, S)                                          # kv_cache.py:1081-1095
        ref.gather_copy_d2d_with_offsets(self.k_buf, self.offsets, self.cnts, B, H, int(self.budget * D), kernel_offset,
                                         kernel_stride, S)                                          # :1140-1150
        pid32 = self.position_ids.to(torch.int32).contiguous()                                      # tensor_op.py:218
        csc = cs.contiguous()
        ref.batch_gather_gemm(self.U.contiguous(), self.SV.contiguous(), csc, csc, pid32, self.output, B, H, self.ctx, D,
                              self.R, self.budget, cs.shape[0], 8, self.cnts)                       # tensor_op.py:220-236
        x, kb = self.output, self.k_buf
        push = ref.apply_rotary_pos_emb_push_cache_opt if cs.shape[-1] == 128 else ref.apply_rotary_pos_emb_push_cache_opt_glm
        push(x, cs, pid32, kb, self.cnts, B, H, self.budget, D, x.stride(0), x.stride(1), x.stride(2), x.stride(3),
             cs.stride(0), pid32.stride(0), pid32.stride(1), pid32.stride(2), kb.stride(0), kb.stride(1), kb.stride(2),
             lay.sparse_start, lay.sparse_end, D // 2, 8)                                            # tensor_op.py:175-195
        cur.wait_stream(self.copy_stream)                                                           # base.py:338
        kv_len = pos + 1
        return flash(q=q.transpose(1, 2), k_cache=kb[:, :, :kv_len].transpose(1, 2),
                     v_cache=self.v_buf[:, :, :kv_len].transpose(1, 2), causal=True)                # base.py:341
