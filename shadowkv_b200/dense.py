"""Dense remainder of the decode layer: torch-tensor wrappers over Part 4 of the C ABI
(include/shadowkv_b200.h).  Each function replaces the library call named in its docstring for the decode case
(one token per sequence, batch <= 64); every output can be preallocated (``out=``) so a whole token can be captured
in a CUDA graph.  No fallback: the kernels are the only implementation.

``static_weights=True`` (model weights: written long before the call) lets the kernel stream W into shared memory
while the previous kernel of the stream is still running (programmatic dependent launch); with the default False
nothing is read before the previous kernel has completed.  ``prefetch=(tensor, byte_offset, nbytes)`` names the head
of the weights the NEXT linear will read: the kernel pulls it into L2 during its own tail."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import lib, check

F32, RESID, SILU_MUL, QKV_ROPE, PARTIAL = 0, 1, 2, 3, 4
MAX_BATCH = 64

_workspaces: dict = {}


def _cs() -> int:
    return torch.cuda.current_stream().cuda_stream


def workspace(device) -> torch.Tensor:
    """Scratch of skv_dense_linear (zero-filled once; the kernel keeps its counters at zero): one per (device, stream),
    because launches of one stream run in order and may share it, launches of different streams may not."""
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(lib.skv_dense_workspace_bytes(), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def _check_x(x: torch.Tensor, W: torch.Tensor):
    if x.dtype != torch.bfloat16 or W.dtype != torch.bfloat16:
        raise TypeError("dense ops take bfloat16 tensors")
    if x.dim() != 2 or W.dim() != 2 or x.shape[1] != W.shape[1]:
        raise ValueError(f"x {tuple(x.shape)} does not match W {tuple(W.shape)} ([out, in])")
    if not x.is_contiguous() or not W.is_contiguous():
        raise ValueError("x and W must be contiguous")
    if not x.is_cuda or not W.is_cuda:
        raise ValueError("x and W must be CUDA tensors")


def _launch(x, W, epi: "_lib.DenseEpilogue", what: str, prefetch=None):
    if prefetch is not None:            # (tensor, byte offset, bytes): head of the next linear's weights -> L2
        t, off, n = prefetch
        n = min(n, t.numel() * t.element_size() - off)
        if n > 0:
            epi.prefetch_next, epi.prefetch_bytes = t.data_ptr() + off, n
    check(lib.skv_dense_linear(x.data_ptr(), W.data_ptr(), x.shape[0], x.shape[1], W.shape[0], C.byref(epi),
                               workspace(x.device).data_ptr(), _cs()), what)


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, out: torch.Tensor | None = None) -> torch.Tensor:
    """``layer_norm`` of reference models/tensor_op.py:34-39 (flashinfer rmsnorm) on [B, hidden] bf16."""
    if x.dim() != 2 or not x.is_contiguous():
        raise ValueError("rmsnorm takes a contiguous [batch, hidden] tensor")
    out = torch.empty_like(x) if out is None else out
    check(lib.skv_rmsnorm(x.data_ptr(), None, None, w.data_ptr(), float(eps), None, out.data_ptr(), x.shape[0],
                          x.shape[1], _cs()), "rmsnorm")
    return out


def embed_rmsnorm(tok: torch.Tensor, emb: torch.Tensor, w: torch.Tensor, eps: float, x_out: torch.Tensor,
                  h_out: torch.Tensor) -> torch.Tensor:
    """``F.embedding`` (reference models/base.py:141) + the first layer's input rmsnorm in one launch:
    x_out[b] = emb[tok[b]], h_out[b] = rmsnorm(x_out[b])."""
    if tok.dtype != torch.int64:
        raise TypeError("token ids must be int64")
    check(lib.skv_rmsnorm(None, tok.data_ptr(), emb.data_ptr(), w.data_ptr(), float(eps), x_out.data_ptr(),
                          h_out.data_ptr(), tok.numel(), emb.shape[1], _cs()), "embed_rmsnorm")
    return h_out


def linear(x: torch.Tensor, W: torch.Tensor, bias: torch.Tensor | None = None, residual: torch.Tensor | None = None,
           out: torch.Tensor | None = None, static_weights: bool = False, prefetch=None) -> torch.Tensor:
    """``F.linear(x, W, bias)`` (+ ``residual + ...``): o_proj and down_proj with their residual adds, reference
    models/llama.py:408-412, 424-426.  ``out`` may be ``residual`` (in-place residual stream)."""
    _check_x(x, W)
    out = torch.empty(x.shape[0], W.shape[0], dtype=torch.bfloat16, device=x.device) if out is None else out
    epi = _lib.DenseEpilogue(mode=RESID, bias=bias.data_ptr() if bias is not None else None,
                             resid=residual.data_ptr() if residual is not None else None, out=out.data_ptr(),
                             static_weights=int(static_weights))
    _launch(x, W, epi, "dense_linear", prefetch)
    return out


_planes: dict = {}


def linear_partial(x: torch.Tensor, W: torch.Tensor, static_weights: bool = False, prefetch=None) -> torch.Tensor:
    """First half of ``residual + F.linear(x, W)`` followed by a ``layer_norm`` (o_proj / down_proj, reference
    models/llama.py:408-413, 424-426): the partial sums of every CTA as fp32 planes, consumed by ``finish_rmsnorm``.
    Same result as ``linear(..., residual=)`` + ``rmsnorm``, bit for bit, without the cross-CTA fix-up inside the
    linear.  The plane buffer is cached per shape: consume it before the next ``linear_partial`` of that shape."""
    _check_x(x, W)
    key = (x.device, torch.cuda.current_stream(x.device).cuda_stream, x.shape[0], x.shape[1], W.shape[0])
    planes = _planes.get(key)
    if planes is None:
        n = lib.skv_dense_partial_bytes(x.shape[0], x.shape[1], W.shape[0])
        planes = torch.zeros(n // 4, dtype=torch.float32, device=x.device)
        _planes[key] = planes
    _launch(x, W, _lib.DenseEpilogue(mode=PARTIAL, out=planes.data_ptr(), static_weights=int(static_weights)),
            "dense_linear_partial", prefetch)
    return planes


def finish_rmsnorm(planes: torch.Tensor, in_features: int, x_inout: torch.Tensor, w: torch.Tensor, eps: float,
                   h_out: torch.Tensor | None = None, bias: torch.Tensor | None = None) -> torch.Tensor:
    """Second half: x_inout <- bf16(x_inout + bf16(sum of planes (+ bias))), returns rmsnorm(x_inout, w, eps)."""
    if x_inout.dim() != 2 or not x_inout.is_contiguous() or x_inout.dtype != torch.bfloat16:
        raise ValueError("finish_rmsnorm takes a contiguous [batch, hidden] bfloat16 residual stream")
    h_out = torch.empty_like(x_inout) if h_out is None else h_out
    check(lib.skv_dense_finish_rmsnorm(planes.data_ptr(), x_inout.shape[0], in_features, x_inout.shape[1],
                                       bias.data_ptr() if bias is not None else None, x_inout.data_ptr(), w.data_ptr(),
                                       float(eps), h_out.data_ptr(), _cs()), "dense_finish_rmsnorm")
    return h_out


def linear_f32(x: torch.Tensor, W: torch.Tensor, out: torch.Tensor | None = None,
               static_weights: bool = False, prefetch=None) -> torch.Tensor:
    """``F.linear(x, lm_head).float()`` (reference models/base.py:157)."""
    _check_x(x, W)
    out = torch.empty(x.shape[0], W.shape[0], dtype=torch.float32, device=x.device) if out is None else out
    _launch(x, W, _lib.DenseEpilogue(mode=F32, out=out.data_ptr(), static_weights=int(static_weights)), "dense_linear_f32", prefetch)
    return out


def linear_silu_mul(x: torch.Tensor, W_gate_up: torch.Tensor, out: torch.Tensor | None = None,
                    static_weights: bool = False, prefetch=None) -> torch.Tensor:
    """``F.linear(x, gate_up_proj)`` + ``vllm._custom_ops.silu_and_mul`` (reference models/llama.py:415-421):
    W_gate_up [2F, in] = [gate; up] -> [B, F]."""
    _check_x(x, W_gate_up)
    F_ = W_gate_up.shape[0] // 2
    out = torch.empty(x.shape[0], F_, dtype=torch.bfloat16, device=x.device) if out is None else out
    _launch(x, W_gate_up, _lib.DenseEpilogue(mode=SILU_MUL, out=out.data_ptr(), static_weights=int(static_weights)), "dense_linear_silu_mul", prefetch)
    return out


def linear_qkv_rope(x: torch.Tensor, wqkv: torch.Tensor, q_heads: int, kv_heads: int, cos_sin: torch.Tensor | None,
                    pos: int, *, bias: torch.Tensor | None = None, glm: bool = False, pos_dev: torch.Tensor | None = None,
                    q_out=None, k_out=None, v_out=None, static_weights: bool = False, prefetch=None,
                    heads: tuple[int, int] | None = None, kv_cache=None):
    """``F.linear(h, wqkv[, bqkv])`` + split + ``vllm._custom_ops.rotary_embedding`` on q and k (reference
    models/llama.py:377-385 + 296; models/glm.py:344-357 + 421).  Returns q [B,q_heads,1,128], k, v [B,kv_heads,1,128]
    -- the layout ``ShadowKVCache_CPU.decode_layer`` takes.  cos_sin None: no rotary embedding.  The token position is
    ``pos`` (+ ``pos_dev[0]``, an int32 device counter, when given).

    heads=(first, count): ``wqkv`` holds only those heads of the fused matrix (q heads, then k, then v; pass the row
    slice ``wqkv_full[first*128:(first+count)*128]``; ``bias`` stays the full vector), so the q projection can run on the
    critical path and the k / v projection beside it.  kv_cache=(k_buf, v_buf, row): k and v are appended IN PLACE at
    row ``row`` (+ ``pos_dev[0]``) of the layer's cache buffers [B,H,rows,128] (``update_kv_cache``,
    kv_cache.py:1227-1271) instead of being returned."""
    _check_x(x, wqkv)
    B, dev = x.shape[0], x.device
    h0, cnt = heads if heads is not None else (0, q_heads + 2 * kv_heads)
    has_q, has_kv = h0 < q_heads, h0 + cnt > q_heads
    q = k = v = None
    if has_q:
        q = torch.empty(B, q_heads, 1, 128, dtype=torch.bfloat16, device=dev) if q_out is None else q_out
    kv_row = kv_sb = kv_sh = kv_rows = 0
    if has_kv:
        if kv_cache is not None:
            k, v, kv_row = kv_cache
            if k.dim() != 4 or k.stride(3) != 1 or k.stride(2) != 128 or k.stride() != v.stride():
                raise ValueError("kv_cache buffers must be [B, H, rows, 128] with contiguous rows")
            kv_sb, kv_sh, kv_rows = k.stride(0), k.stride(1), k.shape[2]
        else:
            k = torch.empty(B, kv_heads, 1, 128, dtype=torch.bfloat16, device=dev) if k_out is None else k_out
            v = torch.empty(B, kv_heads, 1, 128, dtype=torch.bfloat16, device=dev) if v_out is None else v_out
    epi = _lib.DenseEpilogue(mode=QKV_ROPE, q_heads=q_heads, kv_heads=kv_heads,
                             rope_mode=0 if cos_sin is None else (2 if glm else 1),
                             cos_sin_rows=0 if cos_sin is None else cos_sin.shape[0], pos=int(pos),
                             pos_dev=pos_dev.data_ptr() if pos_dev is not None else None,
                             bias=bias.data_ptr() if bias is not None else None,
                             cos_sin=cos_sin.data_ptr() if cos_sin is not None else None,
                             out=q.data_ptr() if q is not None else k.data_ptr(),
                             k_out=k.data_ptr() if k is not None else None, v_out=v.data_ptr() if v is not None else None,
                             static_weights=int(static_weights), head0=h0, head_count=cnt, kv_row=int(kv_row),
                             kv_rows=int(kv_rows), kv_stride_b=kv_sb, kv_stride_h=kv_sh)
    _launch(x, wqkv, epi, "dense_linear_qkv_rope", prefetch)
    return q, k, v


def argmax(logits: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """``sample_token`` with temperature 0 (reference models/tensor_op.py:291-293): [B, vocab] f32 -> [B] int64."""
    if logits.dtype != torch.float32 or logits.dim() != 2 or not logits.is_contiguous():
        raise ValueError("argmax takes a contiguous [batch, vocab] float32 tensor")
    out = torch.empty(logits.shape[0], dtype=torch.int64, device=logits.device) if out is None else out
    check(lib.skv_argmax(logits.data_ptr(), out.data_ptr(), logits.shape[0], logits.shape[1], _cs()), "argmax")
    return out


def l2_prefetch(t: torch.Tensor, nbytes: int | None = None, offset: int = 0) -> None:
    """Hint: pull (a byte range of) a contiguous tensor into L2 on the current stream (weights of an upcoming linear)."""
    total = t.numel() * t.element_size()
    nbytes = total - offset if nbytes is None else min(nbytes, total - offset)
    if nbytes > 0:
        check(lib.skv_l2_prefetch(t.data_ptr() + offset, nbytes, _cs()), "l2_prefetch")
