"""Drop-in for the reference's native module ``kernels.shadowkv`` (reference kernels/main.cu:45-80).

Same 12 names, same positional arguments (torch tensors + ints/floats), all returning ``None``; every
call lands in the C-ABI library on the *current* torch CUDA stream (the reference used the legacy
default stream for all but one of them -- SURVEY.md section 2.1 -- taking the current stream is strictly
safer and keeps the order reorder -> {V fetch || K path} -> attention).

Differences from the reference, on purpose: dtype / device mismatches and unsupported ``map_size`` raise
instead of being silently ignored (reference kernels/gather_copy.cu:97-145 no-ops on map_size outside
{128,256,512,1024}).
"""
from __future__ import annotations

import torch

from ._lib import lib, check

__all__ = [
    "gather_copy", "gather_copy_d2d_with_offsets", "reorder_keys_and_compute_offsets",
    "gather_copy_with_offsets", "apply_rotary_pos_emb", "apply_rotary_pos_emb_new",
    "apply_rotary_pos_emb_new_v2", "apply_rotary_pos_emb_push_cache", "apply_rotary_pos_emb_push_cache_opt",
    "apply_rotary_pos_emb_push_cache_opt_glm", "batch_gather_gemm", "batch_gemm_softmax",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor, dtype, name: str, cuda: bool | None = True) -> int:
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if cuda is True and not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor")
    if cuda is False and not (t.is_cuda or t.is_pinned()):
        raise ValueError(f"{name}: host tensor must be pinned (the kernel reads it over PCIe)")
    return t.data_ptr()


BF, I64, I32, F32 = torch.bfloat16, torch.int64, torch.int32, torch.float32


def gather_copy(values, v_cache_buffer, position_ids, batch_size, heads, cpu_v_length, gpu_v_length, map_size=256):
    """reference kernels/functions.h:71-79"""
    check(lib.skv_gather_copy(_ptr(values, BF, "values", cuda=False), _ptr(v_cache_buffer, BF, "v_cache_buffer"),
                              _ptr(position_ids, I64, "position_ids"), batch_size, heads, cpu_v_length,
                              gpu_v_length, map_size, _stream()), "gather_copy")


def gather_copy_d2d_with_offsets(keys, offsets, cnts, batch_size, heads, gpu_k_length, gpu_k_offset, gpu_k_stride,
                                 map_size=256):
    """reference kernels/functions.h:97-106"""
    check(lib.skv_gather_copy_d2d_with_offsets(_ptr(keys, BF, "keys"), _ptr(offsets, I32, "offsets"),
                                               _ptr(cnts, I32, "cnts"), batch_size, heads, gpu_k_length,
                                               gpu_k_offset, gpu_k_stride, map_size, _stream()),
          "gather_copy_d2d_with_offsets")


def reorder_keys_and_compute_offsets(cached_pos_ids, cur_pos_ids, offsets, cnts, batch_size, heads, map_size=256):
    """reference kernels/functions.h:123-128"""
    check(lib.skv_reorder_keys_and_compute_offsets(_ptr(cached_pos_ids, I64, "cached_pos_ids"),
                                                   _ptr(cur_pos_ids.contiguous(), I64, "cur_pos_ids"),
                                                   _ptr(offsets, I32, "offsets"), _ptr(cnts, I32, "cnts"),
                                                   batch_size, heads, map_size, _stream()),
          "reorder_keys_and_compute_offsets")


def gather_copy_with_offsets(values, v_cache_buffer, temp, offsets, cnts, signals, batch_size, heads, cpu_v_length,
                             gpu_v_length, gpu_v_offset, gpu_v_stride, map_size=256):
    """reference kernels/functions.h:151-158"""
    check(lib.skv_gather_copy_with_offsets(_ptr(values, BF, "values", cuda=False),
                                           _ptr(v_cache_buffer, BF, "v_cache_buffer"),
                                           temp.data_ptr() if temp is not None else 0,
                                           _ptr(offsets, I32, "offsets"), _ptr(cnts, I32, "cnts"),
                                           signals.data_ptr() if signals is not None else 0, batch_size, heads,
                                           cpu_v_length, gpu_v_length, gpu_v_offset, gpu_v_stride, map_size,
                                           _stream()), "gather_copy_with_offsets")


def apply_rotary_pos_emb(x, cos, sin, position_ids, output, batch_size, heads, seq_len, embed_dim, stride_xb,
                         stride_xh, stride_xs, stride_xe, stride_cos, stride_sin, stride_pid_b, stride_pid_h,
                         stride_pid_s, half_dim):
    """reference kernels/functions.h:187-193"""
    check(lib.skv_apply_rotary_pos_emb(_ptr(x, BF, "x"), _ptr(cos, BF, "cos"), _ptr(sin, BF, "sin"),
                                       _ptr(position_ids, I64, "position_ids"), _ptr(output, BF, "output"),
                                       batch_size, heads, seq_len, embed_dim, stride_xb, stride_xh, stride_xs,
                                       stride_xe, stride_cos, stride_sin, stride_pid_b, stride_pid_h, stride_pid_s,
                                       half_dim, _stream()), "apply_rotary_pos_emb")


def apply_rotary_pos_emb_new(x, cos_sin, position_ids, output, batch_size, heads, seq_len, embed_dim, stride_xb,
                             stride_xh, stride_xs, stride_xe, stride_cos_sin, stride_pid_b, stride_pid_h,
                             stride_pid_s, half_dim):
    """reference kernels/functions.h:240-246"""
    check(lib.skv_apply_rotary_pos_emb_new(_ptr(x, BF, "x"), _ptr(cos_sin, BF, "cos_sin"),
                                           _ptr(position_ids, I64, "position_ids"), _ptr(output, BF, "output"),
                                           batch_size, heads, seq_len, embed_dim, stride_xb, stride_xh, stride_xs,
                                           stride_xe, stride_cos_sin, stride_pid_b, stride_pid_h, stride_pid_s,
                                           half_dim, _stream()), "apply_rotary_pos_emb_new")


def apply_rotary_pos_emb_new_v2(x, cos_sin, position_ids, output, batch_size, heads, seq_len, embed_dim, stride_xb,
                                stride_xh, stride_xs, stride_xe, stride_cos_sin, stride_pid_b, stride_pid_h,
                                stride_pid_s, half_dim, chunk_size):
    """reference kernels/functions.h:281-287"""
    check(lib.skv_apply_rotary_pos_emb_new_v2(_ptr(x, BF, "x"), _ptr(cos_sin, BF, "cos_sin"),
                                              _ptr(position_ids, I32, "position_ids"), _ptr(output, BF, "output"),
                                              batch_size, heads, seq_len, embed_dim, stride_xb, stride_xh,
                                              stride_xs, stride_xe, stride_cos_sin, stride_pid_b, stride_pid_h,
                                              stride_pid_s, half_dim, chunk_size, _stream()),
          "apply_rotary_pos_emb_new_v2")


def _push(fn, name, x, cos_sin, position_ids, output_cache, cnts, *ints):
    check(fn(_ptr(x, BF, "x"), _ptr(cos_sin, BF, "cos_sin"), _ptr(position_ids, I32, "position_ids"),
             _ptr(output_cache, BF, "output_cache"), _ptr(cnts, I32, "cnts"), *ints, _stream()), name)


def apply_rotary_pos_emb_push_cache(x, cos_sin, position_ids, output_cache, cnts, *ints):
    """reference kernels/functions.h:328-337 (19 trailing ints)"""
    _push(lib.skv_apply_rotary_pos_emb_push_cache, "apply_rotary_pos_emb_push_cache", x, cos_sin, position_ids,
          output_cache, cnts, *ints)


def apply_rotary_pos_emb_push_cache_opt(x, cos_sin, position_ids, output_cache, cnts, *ints):
    """reference kernels/functions.h:362-371"""
    _push(lib.skv_apply_rotary_pos_emb_push_cache_opt, "apply_rotary_pos_emb_push_cache_opt", x, cos_sin,
          position_ids, output_cache, cnts, *ints)


def apply_rotary_pos_emb_push_cache_opt_glm(x, cos_sin, position_ids, output_cache, cnts, *ints):
    """reference kernels/functions.h:396-405"""
    _push(lib.skv_apply_rotary_pos_emb_push_cache_opt_glm, "apply_rotary_pos_emb_push_cache_opt_glm", x, cos_sin,
          position_ids, output_cache, cnts, *ints)


def batch_gather_gemm(a, b, cos, sin, position_ids, output, batch_size, heads, seq_len, embed_dim, rank,
                      sparse_budget, max_seq_len, chunk_size, offset_array):
    """reference kernels/functions.h:431-437"""
    check(lib.skv_batch_gather_gemm(_ptr(a, BF, "a"), _ptr(b, BF, "b"), cos.data_ptr(), sin.data_ptr(),
                                    _ptr(position_ids, I32, "position_ids"), _ptr(output, BF, "output"),
                                    batch_size, heads, seq_len, embed_dim, rank, sparse_budget, max_seq_len,
                                    chunk_size, _ptr(offset_array, I32, "offset_array"), _stream()),
          "batch_gather_gemm")


def batch_gemm_softmax(A, B, D, Norm, Sum, Softmax, batch_count, m, n, k, alpha=1.0, beta=0.0):
    """reference kernels/functions.h:460-463"""
    check(lib.skv_batch_gemm_softmax(_ptr(A, BF, "A"), _ptr(B, BF, "B"), _ptr(D, BF, "D"), _ptr(Norm, F32, "Norm"),
                                     _ptr(Sum, F32, "Sum"), _ptr(Softmax, BF, "Softmax"), batch_count, m, n, k,
                                     float(alpha), float(beta), _stream()), "batch_gemm_softmax")
