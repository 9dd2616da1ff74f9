"""CPU restatement of the reference's per-layer sparse-KV decode step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``shadowkv_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or as
the timed CPU baseline.

Every function cites the reference file:line it restates (paths relative to the
reference checkout).  Arithmetic is torch-on-CPU / numpy; bf16 rounding points
are reproduced where the reference has them.

Pinning status
--------------
* integer path (reorder / cached gather): pinned against the reference's own
  known-answer generator ``kernels/test_cached_gather_copy.cu:70-216`` (same
  glibc ``srand(12345)`` stream) in ``tests/test_oracle_golden.py`` and against
  outputs of the compiled reference extension captured on a B200
  (``tests/golden/ref_*.npz``, made by ``tests/golden/make_ref_golden.py``).
* floating-point path (scoring, K rebuild, RoPE, attention): the reference has
  no tests for it ("parity unpinned" by the reference itself); pinned here only
  through the same captured outputs of the reference extension, within the bf16
  tolerance stated in ``tests/tolerances.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

BF16 = torch.bfloat16


# --------------------------------------------------------------------------
# (a3) landmark scoring          models/kv_cache.py:1006-1019
#                                kernels/batch_gemm_softmax.cu:229-293
# --------------------------------------------------------------------------
def landmark_scores(q: torch.Tensor, landmarks: torch.Tensor, scale: float = 1.0 / math.sqrt(128)):
    """q [B,Hq,1,D] bf16 (post-RoPE), landmarks [B,H,N,D] bf16.

    Returns (gemm_o, softmax_o), both [B,H,G,N] bf16:
      gemm_o    = bf16(scale * (q . landmark))           f32 accumulate
                  (CUTLASS LinearCombination<bf16> epilogue, batch_gemm_softmax.cu:260-281)
      softmax_o = bf16(exp(f32(gemm_o) - rowmax) * (1/rowsum))
                  (epilogue_visitor_with_softmax max/sum over the bf16-rounded
                  values; ApplySoftmax kernels/batch_gemm_softmax.h:289-292)
    """
    B, Hq, qlen, D = q.shape
    H = landmarks.shape[1]
    G = Hq // H
    A = q.reshape(B, H, G * qlen, D).float()
    Bm = landmarks.float()
    acc = torch.matmul(A, Bm.transpose(-1, -2))                  # [B,H,G,N] f32
    gemm_o = (acc * scale).to(BF16)
    s = gemm_o.float()
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    inv = 1.0 / e.sum(dim=-1, keepdim=True)
    softmax_o = (e * inv).to(BF16)
    return gemm_o, softmax_o


# --------------------------------------------------------------------------
# (a4) GQA max + top-k + id gather     models/kv_cache.py:1023-1042
# --------------------------------------------------------------------------
def select_chunks(softmax_o: torch.Tensor, landmark_idx: torch.Tensor, S: int):
    """softmax_o [B,H,G,N] bf16, landmark_idx [B,H,N] int64 -> chunk ids [B,H,S].

    ``torch.topk`` leaves the order among equal values unspecified; this
    restatement fixes it as lowest landmark index first (stable descending
    sort), which is the rule the CUDA path implements.  Returned in top-k order
    (value descending, index ascending among ties).
    """
    score = softmax_o.float().max(dim=-2).values                 # [B,H,N]; max of bf16 is exact
    order = torch.sort(score, dim=-1, descending=True, stable=True).indices
    top = order[..., :S]
    return landmark_idx.gather(-1, top), top, score.to(BF16)


def selection_equal_modulo_ties(score: torch.Tensor, got_top: torch.Tensor, S: int) -> bool:
    """Set parity for one row: every landmark scoring strictly above the S-th
    value must be present; the rest must score exactly the S-th value."""
    score = score.float()
    kth = torch.sort(score, descending=True).values[S - 1]
    got = torch.zeros_like(score, dtype=torch.bool)
    got[got_top] = True
    if int(got.sum()) != S:
        return False
    if not bool(got[score > kth].all()):
        return False
    return bool((score[got] >= kth).all())


# --------------------------------------------------------------------------
# (a5) chunk-cache bookkeeping     kernels/gather_copy.cu:259-307
#                                  kernels/map.cuh:754-796 (+ :508-539 block_sort2)
# --------------------------------------------------------------------------
def reorder_keys_and_compute_offsets(cached: np.ndarray, cur: np.ndarray):
    """cached, cur: int64 [rows, S] (rows = B*H).  Returns (new_cached int64
    [rows,S], offsets int32 [rows,S], cnts int32 [rows]).

    Per row: ids are truncated to int32 (map.cuh:771,776).  A selected id is a
    *hit* if it is resident; its offset is the old slot.  Hits come first,
    ordered by old slot ascending; misses follow ordered by chunk id ascending
    with offset = chunk id.  -1 is the reserved empty key.  If the resident
    table holds the same id twice the reference's winner is a data race
    (map.cuh:82-95); this restatement (and the CUDA path) take the lowest slot.
    """
    cached = np.asarray(cached, dtype=np.int64)
    cur = np.asarray(cur, dtype=np.int64)
    rows, S = cached.shape
    new_cached = np.empty_like(cached)
    offsets = np.empty((rows, S), dtype=np.int32)
    cnts = np.empty(rows, dtype=np.int32)
    for r in range(rows):
        old = cached[r].astype(np.int32)
        new = cur[r].astype(np.int32)
        slot_of = {}
        for j in range(S - 1, -1, -1):          # lowest slot wins
            slot_of[int(old[j])] = j
        slot_of.pop(-1, None)
        hits = sorted((slot_of[int(k)], int(k)) for k in new if int(k) in slot_of)
        misses = sorted(int(k) for k in new if int(k) not in slot_of)
        c = len(hits)
        cnts[r] = c
        for i, (slot, k) in enumerate(hits):
            new_cached[r, i] = k
            offsets[r, i] = slot
        for i, k in enumerate(misses):
            new_cached[r, c + i] = k
            offsets[r, c + i] = k
    return new_cached, offsets, cnts


# --------------------------------------------------------------------------
# (a6)/(a7) cached gathers         kernels/copy.cuh:649-687 (K, D2D)
#                                  kernels/copy.cuh:785-846 (V, D2D + H2D)
# --------------------------------------------------------------------------
def gather_copy_d2d(buf_sparse: torch.Tensor, offsets: np.ndarray, cnts: np.ndarray, chunk_tokens: int = 8):
    """In-place forward compaction of hit chunks.  buf_sparse [rows, S*C, D]
    (the sparse region of one layer's K or V buffer), modified in place:
    slot i < cnt  <-  slot offsets[i].  Slots >= cnt are left untouched."""
    rows = buf_sparse.shape[0]
    S = offsets.shape[1]
    view = buf_sparse.view(rows, S, -1)
    for r in range(rows):
        c = int(cnts[r])
        if c:
            src = torch.from_numpy(offsets[r, :c].astype(np.int64))
            view[r, :c] = view[r, src].clone()
    return buf_sparse


def gather_copy_with_offsets(v_host: torch.Tensor, buf_sparse: torch.Tensor, offsets: np.ndarray,
                             cnts: np.ndarray):
    """v_host [rows, chunks, C*D] (pinned store, chunk-major, kv_cache.py:554-563),
    buf_sparse [rows, S*C, D].  Hits compacted in place, misses fetched from the
    host store: slot i >= cnt  <-  v_host[row, offsets[i]]."""
    rows = buf_sparse.shape[0]
    S = offsets.shape[1]
    gather_copy_d2d(buf_sparse, offsets, cnts)
    view = buf_sparse.view(rows, S, -1)
    for r in range(rows):
        c = int(cnts[r])
        if c < S:
            src = torch.from_numpy(offsets[r, c:].astype(np.int64))
            view[r, c:] = v_host[r, src]
    return buf_sparse


def gather_copy(v_host: torch.Tensor, position_ids: np.ndarray):
    """Uncached gather of all S chunks (kernels/copy.cuh:481-508)."""
    rows = v_host.shape[0]
    idx = torch.from_numpy(np.asarray(position_ids, dtype=np.int64)).view(rows, -1)
    return torch.stack([v_host[r, idx[r]] for r in range(rows)])


# --------------------------------------------------------------------------
# RoPE in bf16                      kernels/rope_new.cu:321-370 (neox / Llama, Yi)
#                                   kernels/rope_new.cu:429-493 (GLM, interleaved)
# --------------------------------------------------------------------------
def _bf(x):
    return x.to(BF16)


def rope_bf16(x: torch.Tensor, cos_sin_rows: torch.Tensor, fused: bool = True) -> torch.Tensor:
    """x [..., D] bf16, cos_sin_rows [..., D] (neox: [cos(D/2)|sin(D/2)]) or
    [..., D/2] (GLM: [cos(D/4)|sin(D/4)], rotates the first D/2 dims in
    interleaved pairs, copies the rest).

    ``o1 = __hadd(__hmul(x1,c), __hmul(__hneg(x2),s))`` -- every product and the
    sum individually rounded to bf16.  ``fused=True`` models ptxas contracting
    the add with its first product into one bf16 FMA (the second product is
    still rounded), which is what SASS of the reference shows for sm_100a; see
    DESIGN.md "RoPE rounding".
    """
    D = x.shape[-1]
    x = x.to(BF16)
    cs = cos_sin_rows.to(BF16)

    def comb(a, ca, b, sb, neg):
        # a*ca + (+-b)*sb
        pb = _bf((-b.float() if neg else b.float()) * sb.float())
        if fused:
            return _bf(a.float().double() * ca.float().double() + pb.float().double())
        pa = _bf(a.float() * ca.float())
        return _bf(pa.float() + pb.float())

    if cs.shape[-1] == D:                       # neox half-split
        h = D // 2
        x1, x2 = x[..., :h], x[..., h:]
        c, s = cs[..., :h], cs[..., h:]
        o1 = comb(x1, c, x2, s, True)
        o2 = comb(x2, c, x1, s, False)
        return torch.cat([o1, o2], dim=-1)
    if cs.shape[-1] == D // 2:                  # GLM
        q = D // 4
        xe, xo = x[..., 0:2 * q:2], x[..., 1:2 * q:2]
        c, s = cs[..., :q], cs[..., q:]
        oe = comb(xe, c, xo, s, True)
        oo = comb(xo, c, xe, s, False)
        rot = torch.stack([oe, oo], dim=-1).flatten(-2)
        return torch.cat([rot, x[..., 2 * q:]], dim=-1)
    raise ValueError(f"bad cos_sin width {cs.shape[-1]} for head_dim {D}")


# --------------------------------------------------------------------------
# (a7) K rebuild                    kernels/batch_gather_gemm.cu:193-287
#                                   kernels/special_batch_gather_predicated_tile_iterator.h:294
#                                   models/tensor_op.py:201-238
# --------------------------------------------------------------------------
def rebuild_keys_pre_rope(U: torch.Tensor, SV: torch.Tensor, chunk_ids: torch.Tensor, C: int = 8):
    """U [B,L,R] bf16, SV [B,H,D,R] bf16 (R contiguous), chunk_ids [B,H,S] ->
    [B,H,S*C,D] bf16 = bf16(sum_k U[b, id*C+r, k] * SV[b,h,:,k]) (f32 accumulate)."""
    B, H, S = chunk_ids.shape
    tok = (chunk_ids.unsqueeze(-1) * C + torch.arange(C)).view(B, H, S * C)      # token positions
    out = torch.empty(B, H, S * C, SV.shape[2], dtype=BF16)
    for b in range(B):
        Ub = U[b].float()
        for h in range(H):
            out[b, h] = (Ub[tok[b, h]] @ SV[b, h].float().t()).to(BF16)
    return out, tok


def rebuild_keys(U, SV, chunk_ids, cos_sin, C: int = 8, fused_rope: bool = True):
    """K rebuild + RoPE at the retrieval position ids (rope_new.cu:348-367)."""
    kpre, tok = rebuild_keys_pre_rope(U, SV, chunk_ids, C)
    return rope_bf16(kpre, cos_sin[tok], fused=fused_rope)


# --------------------------------------------------------------------------
# (a8) attention                    models/base.py:341 (flash_attn_with_kvcache stand-in)
# --------------------------------------------------------------------------
def attention(q: torch.Tensor, K: torch.Tensor, V: torch.Tensor):
    """q [B,Hq,1,D], K,V [B,H,Lp,D] -> [B,1,Hq,D] bf16; f32 softmax(q K^T/sqrt(D)) V."""
    B, Hq, _, D = q.shape
    H = K.shape[1]
    G = Hq // H
    qf = q.reshape(B, H, G, D).float()
    s = torch.matmul(qf, K.float().transpose(-1, -2)) / math.sqrt(D)
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, V.float())                               # [B,H,G,D]
    return o.reshape(B, 1, Hq, D).to(BF16)


# --------------------------------------------------------------------------
# (a1) state + the whole step       models/kv_cache.py:509-1319 (ShadowKVCache_CPU)
# --------------------------------------------------------------------------
@dataclass
class LayerState:
    """One layer of ShadowKVCache_CPU state."""
    landmarks: torch.Tensor        # [B,H,N,D] bf16
    landmark_idx: torch.Tensor     # [B,H,N]   int64
    U: torch.Tensor                # [B,L,R]   bf16
    SV: torch.Tensor               # [B,H,D,R] bf16
    v_host: torch.Tensor           # [B,H,chunks,C*D] bf16
    k_buf: torch.Tensor            # [B,H,Lbuf,D] bf16
    v_buf: torch.Tensor            # [B,H,Lbuf,D] bf16
    resident: torch.Tensor         # [B,H,S] int64   (position_ids[layer])


@dataclass
class Geometry:
    sparse_start: int
    sparse_end: int
    chunk: int = 8
    gen_offset: int = 0
    extra: dict = field(default_factory=dict)


def decode_step(st: LayerState, geo: Geometry, q, k_new, v_new, cos_sin, fused_rope=True, selected_override=None):
    """One layer of the cached decode step, Appendix B of SURVEY.md; mutates st.
    Returns (attn_out [B,1,Hq,D] bf16, info dict).  `selected_override` [B,H,S] replaces the oracle's own
    top-k result (used by the parity tests after the selection itself has been checked modulo ties, so
    that one boundary chunk does not decide every downstream comparison)."""
    B, H, S = st.resident.shape
    C = geo.chunk
    ss, se = geo.sparse_start, geo.sparse_end
    pos = se + geo.gen_offset
    # (a2) append     kv_cache.py:1255-1265
    st.k_buf[:, :, pos:pos + 1] = k_new
    st.v_buf[:, :, pos:pos + 1] = v_new
    # (a3)(a4)
    gemm_o, softmax_o = landmark_scores(q, st.landmarks)
    selected, top, score = select_chunks(softmax_o, st.landmark_idx, S)
    own_selected = selected
    if selected_override is not None:
        selected = selected_override.to(torch.int64).cpu()
    # (a5)
    new_res, offsets, cnts = reorder_keys_and_compute_offsets(
        st.resident.view(B * H, S).numpy(), selected.reshape(B * H, S).numpy())
    st.resident.copy_(torch.from_numpy(new_res).view(B, H, S))
    # (a6) V
    vs = st.v_buf[:, :, ss:se].reshape(B * H, S * C, -1).clone()
    gather_copy_with_offsets(st.v_host.view(B * H, st.v_host.shape[2], -1), vs, offsets, cnts)
    st.v_buf[:, :, ss:se] = vs.view(B, H, S * C, -1)
    # (a7) K
    ks = st.k_buf[:, :, ss:se].reshape(B * H, S * C, -1).clone()
    gather_copy_d2d(ks, offsets, cnts)
    krope = rebuild_keys(st.U, st.SV, st.resident, cos_sin, C, fused_rope).view(B * H, S, C, -1)
    ksv = ks.view(B * H, S, C, -1)
    for r in range(B * H):
        c = int(cnts[r])
        ksv[r, c:] = krope[r, c:]
    st.k_buf[:, :, ss:se] = ks.view(B, H, S * C, -1)
    # (a8)
    out = attention(q, st.k_buf[:, :, :pos + 1], st.v_buf[:, :, :pos + 1])
    return out, dict(gemm_o=gemm_o, softmax_o=softmax_o, score=score, top=top, selected=selected,
                     own_selected=own_selected, offsets=offsets, cnts=cnts)


# --------------------------------------------------------------------------
# (a10) uncached pure-torch variant  models/kv_cache.py:421-470 (ShadowKVCache)
# --------------------------------------------------------------------------
def decode_step_uncached(landmarks, landmark_idx, U, SV_rd, v_full, k_buf, v_buf, geo: Geometry,
                         q, k_new, v_new, rope_fn, fp32_accumulate: bool = False, selected_override=None):
    """The batch-1 reference PyTorch path: no chunk cache, everything rebuilt
    every step, selection kept in top-k order.  SV_rd is [B,H,R,D]
    (kv_cache.py:304), v_full [B,H,L,D].  rope_fn(x, token_positions) is the
    model's apply_rotary_pos_emb_single.  This is what ``--impl reference`` times.

    fp32_accumulate=True (parity tests): the two bf16 contractions are evaluated in fp32 and rounded to bf16
    once, which is what the reference gets from cuBLAS on the GPU; torch's CPU bf16 einsum may round partial
    sums, which is fine for timing but not for comparing numbers."""
    B, H, N, D = landmarks.shape
    Hq = q.shape[1]
    G = Hq // H
    C = geo.chunk
    ss, se = geo.sparse_start, geo.sparse_end
    S = (se - ss) // C
    pos = se + geo.gen_offset
    k_buf[:, :, pos:pos + 1] = k_new
    v_buf[:, :, pos:pos + 1] = v_new
    # kv_cache.py:425-434
    if fp32_accumulate:
        attn = (torch.einsum('bhgqd,bhdc->bhgqc', q.view(B, H, G, 1, D).float(), landmarks.transpose(2, 3).float())
                .to(BF16).squeeze(2) / math.sqrt(128))
    else:
        attn = torch.einsum('bhgqd,bhdc->bhgqc', q.view(B, H, G, 1, D), landmarks.transpose(2, 3)).squeeze(2) / math.sqrt(128)
    attn = torch.softmax(attn, dim=-1, dtype=torch.float32).to(BF16)
    attn = attn.sum(dim=-2) if attn.dim() == 5 else attn
    attn = attn.view(B, H, G, N)
    if G > 1:
        attn = attn.max(dim=-2).values
    else:
        attn = attn.squeeze(-2)
    top = torch.topk(attn, k=S, dim=-1).indices
    sel = own_sel = landmark_idx.gather(-1, top)
    if selected_override is not None:       # parity tests: same chunk set as the path under test (see decode_step)
        sel = selected_override.to(torch.int64).cpu()
    tok = (sel.unsqueeze(-1) * C + torch.arange(C)).view(B, H, -1)
    # kv_cache.py:449-450
    v_buf[:, :, ss:se] = v_full.gather(-2, tok.unsqueeze(-1).expand(-1, -1, -1, D))
    # kv_cache.py:457-467
    u_exp = U.unsqueeze(1).expand(-1, H, -1, -1)
    U_head = torch.gather(u_exp, 2, tok.unsqueeze(-1).expand(-1, -1, -1, U.shape[-1]))
    if fp32_accumulate:
        k = torch.einsum('bhrk,bhkd->bhrd', U_head.float(), SV_rd.float()).to(BF16)
    else:
        k = torch.einsum('bhrk,bhkd->bhrd', U_head, SV_rd)
    k_buf[:, :, ss:se] = rope_fn(k, tok)
    out = attention(q, k_buf[:, :, :pos + 1], v_buf[:, :, :pos + 1])
    return out, own_sel


def rope_torch(x, cos_sin, tok):
    """models/tensor_op.py:127-151 (rotate_half form) with the combined
    [cos(64)|sin(64)] table of models/llama.py:332; GLM form models/glm.py:454-469."""
    D = x.shape[-1]
    cs = cos_sin[tok]
    if cs.shape[-1] == D:
        h = D // 2
        cos = torch.cat([cs[..., :h], cs[..., :h]], -1)
        sin = torch.cat([cs[..., h:], cs[..., h:]], -1)
        rot = torch.cat((-x[..., h:], x[..., :h]), dim=-1)
        return (x * cos) + (rot * sin)
    q = D // 4
    xr, xp = x[..., :2 * q], x[..., 2 * q:]
    o = torch.stack([xr[..., 0::2] * cs[..., :q] - xr[..., 1::2] * cs[..., q:],
                     xr[..., 1::2] * cs[..., :q] + xr[..., 0::2] * cs[..., q:]], -1).flatten(-2)
    return torch.cat((o, xp), dim=-1)


# --------------------------------------------------------------------------
# (f1) prefill-side state build      models/kv_cache.py:788-980 (prefill_kv_cache)
# --------------------------------------------------------------------------
def chunk_stats(k_ctx: torch.Tensor):
    """k_ctx [B,H,chunks,C,D] bf16 (post-RoPE prompt keys, chunked).  Returns (means [B,H,chunks,D] bf16,
    mincos [B,H,chunks] bf16): the landmark candidates ``k_ctx.mean(dim=-2)`` (kv_cache.py:856) and
    ``cosine_similarity(mean, key, dim=-1).min(dim=-1)`` (:859-868).

    The reference calls torch ops on bf16 tensors, so every intermediate is rounded to bf16.  The chain below is
    torch 2.3's ``cosine_similarity`` (ATen/native/Distance.cpp; the version the reference pins,
    requirements.txt): ``n = linalg_vector_norm(x)`` (fp32 accumulate, bf16 result) -> ``clamp_min(eps=1e-8)`` ->
    ``(x1 / n1) * (x2 / n2)`` elementwise (each bf16) -> ``sum`` (fp32 accumulate, bf16 result).  It is spelled
    out in fp32 with explicit roundings so it does not depend on the torch version installed here;
    tests/test_oracle_props.py checks it against ``torch.nn.functional.cosine_similarity`` of this image."""
    x = k_ctx.float()
    means = (x.sum(dim=-2) * (1.0 / k_ctx.shape[-2])).to(BF16)
    m = means.float().unsqueeze(-2)

    def norm(t):
        return t.pow(2).sum(dim=-1, keepdim=True).sqrt().to(BF16).float().clamp_min(1e-8)

    a = (m / norm(m)).to(BF16).float()
    b = (x / norm(x)).to(BF16).float()
    cos = (a * b).to(BF16).float().sum(dim=-1).to(BF16)
    return means, cos.float().min(dim=-1).values.to(BF16)


def select_outliers(mincos: torch.Tensor, n_out: int):
    """mincos [B,H,chunks] bf16 -> (outliers [B,H,n_out] ascending, rest_idx [B,H,chunks-n_out] ascending).
    ``topk(largest=False)`` (kv_cache.py:868) with ties resolved to the lowest chunk index; the reference keeps
    the outliers in top-k order, which only permutes rows of the attention buffer."""
    order = torch.sort(mincos.float(), dim=-1, descending=False, stable=True).indices
    out = order[..., :n_out].sort(dim=-1).values
    keep = torch.ones_like(mincos, dtype=torch.bool)
    keep.scatter_(-1, out, False)
    B, H, n = mincos.shape
    rest = torch.arange(n).expand(B, H, n)[keep].view(B, H, n - n_out)
    return out, rest


def prefill_state(k_roped: torch.Tensor, v: torch.Tensor, last_q: torch.Tensor, *, budget: int = 2048, chunk: int = 8,
                  local_chunk: int = 4, buf_len: int | None = None, outliers_override=None, selected_override=None):
    """One layer of ShadowKVCache_CPU.prefill_kv_cache (kv_cache.py:788-980) for k_roped, v [B,H,L,D] bf16 and the
    prompt's last query last_q [B,Hq,D].  Returns a dict with the host V store, the K/V buffers, landmarks, landmark
    index, outlier and selected chunk ids."""
    B, H, L, D = v.shape
    n_out = (budget // 1024) * 24                                            # :548
    chunks = L // chunk - local_chunk
    chunks -= chunks % 8                                                     # :836
    prefill_local = L - chunks * chunk
    ss = prefill_local + n_out * chunk
    se = ss + budget
    buf_len = buf_len or budget + 128 + (n_out + local_chunk) * chunk        # :574
    S = budget // chunk
    v_store = v[:, :, :(L // chunk) * chunk].reshape(B, H, L // chunk, chunk * D).clone()      # :819-823
    k_buf = torch.zeros(B, H, buf_len, D, dtype=BF16)
    v_buf = torch.zeros(B, H, buf_len, D, dtype=BF16)
    k_buf[:, :, :prefill_local] = k_roped[:, :, -prefill_local:]             # :842-844
    v_buf[:, :, :prefill_local] = v[:, :, -prefill_local:]
    k_ctx = k_roped[:, :, :chunks * chunk].view(B, H, chunks, chunk, D)
    v_ctx = v[:, :, :chunks * chunk].view(B, H, chunks, chunk, D)
    means, mincos = chunk_stats(k_ctx)
    outliers, rest = select_outliers(mincos, n_out)
    own_outliers = outliers
    if outliers_override is not None:
        outliers = outliers_override.cpu().sort(dim=-1).values
        keep = torch.ones(B, H, chunks, dtype=torch.bool)
        keep.scatter_(-1, outliers, False)
        rest = torch.arange(chunks).expand(B, H, chunks)[keep].view(B, H, chunks - n_out)
    ex = outliers[..., None, None].expand(-1, -1, -1, chunk, D)
    k_buf[:, :, prefill_local:ss] = k_ctx.gather(2, ex).reshape(B, H, -1, D)  # :870-907
    v_buf[:, :, prefill_local:ss] = v_ctx.gather(2, ex).reshape(B, H, -1, D)
    landmarks = means.gather(2, rest.unsqueeze(-1).expand(-1, -1, -1, D))     # :917-921
    # first fill of the sparse region with the prompt's last query (:926-970)
    G = last_q.shape[1] // H
    _, soft = landmark_scores(last_q.view(B, H * G, 1, D), landmarks)
    selected, _, _ = select_chunks(soft, rest, S)
    own_selected = selected
    if selected_override is not None:
        selected = selected_override.cpu()
    tok = (selected.unsqueeze(-1) * chunk + torch.arange(chunk)).view(B, H, -1)
    exs = tok.unsqueeze(-1).expand(-1, -1, -1, D)
    v_buf[:, :, ss:se] = v.gather(-2, exs)
    k_buf[:, :, ss:se] = k_roped.gather(-2, exs)
    return dict(v_store=v_store, k_buf=k_buf, v_buf=v_buf, means=means, mincos=mincos, outliers=outliers,
                own_outliers=own_outliers, rest_idx=rest, landmarks=landmarks, selected=selected,
                own_selected=own_selected, prefill_local=prefill_local, sparse_start=ss, sparse_end=se, chunks=chunks)


def low_rank_factors(k_pre: torch.Tensor, rank: int, H: int):
    """get_svd (kv_cache.py:666-737): k_pre [B,L,H*D] pre-RoPE keys -> U [B,L,rank] bf16, SV [B,H,D,rank] bf16 from
    the rank-truncated SVD in fp32.  Singular vectors are defined up to sign (and rotation inside repeated
    singular values), so parity is on the product U.SV, never on the factors."""
    u, s, vh = torch.linalg.svd(k_pre.float(), full_matrices=False)
    B = k_pre.shape[0]
    U = u[:, :, :rank].to(BF16)
    sv = (s[:, :rank, None] * vh[:, :rank]).to(BF16)                          # [B, rank, H*D]
    SV = sv.view(B, rank, H, -1).permute(0, 2, 3, 1).contiguous()
    return U, SV
