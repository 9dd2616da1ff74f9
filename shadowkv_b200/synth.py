"""Synthetic decode state for tests, smoke and bench (there is no hub access, so no real weights or
prompts): shapes follow SURVEY.md section 8 (a1)/(d); values are seeded normal draws sized so that the
rebuilt keys have O(1) entries.  Model shapes come from the public model cards (they are not in the
reference tree)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace

import torch

BF16 = torch.bfloat16


@dataclass(frozen=True)
class ModelShape:
    name: str
    num_hidden_layers: int
    hidden_size: int
    num_attention_heads: int
    num_key_value_heads: int
    intermediate_size: int
    vocab_size: int
    rope_theta: float = 500000.0
    glm_rope: bool = False

    def config(self, layers: int | None = None):
        return SimpleNamespace(num_hidden_layers=layers or self.num_hidden_layers, hidden_size=self.hidden_size,
                               num_attention_heads=self.num_attention_heads,
                               num_key_value_heads=self.num_key_value_heads)


MODELS = {
    "llama-3.1-8b": ModelShape("Llama-3.1-8B", 32, 4096, 32, 8, 14336, 128256),
    "llama-3-8b-1m": ModelShape("Llama-3-8B-1M", 32, 4096, 32, 8, 14336, 128256, rope_theta=3580165449.0),
    "yi-9b-200k": ModelShape("Yi-9B-200K", 48, 4096, 32, 4, 11008, 64000, rope_theta=10000000.0),
    "glm-4-9b-1m": ModelShape("GLM-4-9B-1M", 40, 4096, 32, 4, 13696, 151552, rope_theta=10000.0, glm_rope=True),
}


def cos_sin_table(rows: int, head_dim: int = 128, theta: float = 500000.0, glm: bool = False, device="cpu"):
    """Llama/Yi: [rows, 128] = [cos(64) | sin(64)] (reference models/llama.py:332).
    GLM: [rows, 64] = [cos(32) | sin(32)], rotary dim 64 (reference models/glm.py:273)."""
    rot = head_dim // 2 if glm else head_dim
    inv = 1.0 / (theta ** (torch.arange(0, rot, 2, dtype=torch.float64, device=device) / rot))
    ang = torch.arange(rows, dtype=torch.float64, device=device)[:, None] * inv[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=-1).to(BF16).contiguous()


def layout(prompt_len: int, sparse_budget: int = 2048, chunk: int = 8, uncached: bool = False):
    """Region boundaries as the reference computes them (models/kv_cache.py:547-548, 819-894).
    uncached=True: the batch-1 GPU-resident class (48 outlier chunks, buffer budget+4096; kv_cache.py:206,233)."""
    local_chunk = 4
    outlier_chunk = 48 if uncached else (sparse_budget // 1024) * 24
    chunks = prompt_len // chunk - local_chunk
    chunks -= chunks % 8
    prefill_local = prompt_len - chunks * chunk
    sparse_start = prefill_local + outlier_chunk * chunk
    return SimpleNamespace(chunks=chunks, n_landmarks=chunks - outlier_chunk, prefill_local=prefill_local,
                           outlier_chunk=outlier_chunk, sparse_start=sparse_start,
                           sparse_end=sparse_start + sparse_budget,
                           buf_len=sparse_budget + 4096 if uncached else
                           sparse_budget + 128 + (outlier_chunk + local_chunk) * chunk,
                           select=sparse_budget // chunk, all_chunks=prompt_len // chunk)


def _rope_float(x, cs):
    """fp32 rotary embedding of x [..., D] with table rows cs ([cos|sin], neox or GLM layout)."""
    D = x.shape[-1]
    if cs.shape[-1] == D:
        h = D // 2
        c, s = cs[..., :h], cs[..., h:]
        return torch.cat([x[..., :h] * c - x[..., h:] * s, x[..., h:] * c + x[..., :h] * s], dim=-1)
    q = D // 4
    c, s = cs[..., :q], cs[..., q:]
    xe, xo = x[..., 0:2 * q:2], x[..., 1:2 * q:2]
    rot = torch.stack([xe * c - xo * s, xo * c + xe * s], dim=-1).flatten(-2)
    return torch.cat([rot, x[..., 2 * q:]], dim=-1)


def layer_state(shape: ModelShape, batch: int, prompt_len: int, *, seed: int, rank: int = 160,
                sparse_budget: int = 2048, chunk: int = 8, device="cpu", cos_sin=None, uncached: bool = False):
    """One layer of decode state.  Returns a dict of tensors:
    U [B,L,R], SV [B,H,D,R], k_landmark [B,H,N,D], k_landmark_idx [B,H,N] (ascending, `outlier_chunk` random
    chunks removed), v_chunks [B,H,L/C,C*D], k_buf / v_buf [B,H,Lbuf,D] (local + outlier rows random, sparse
    region filled consistently with `resident`: V rows are the store's chunks; K rows are RoPE(U.SV) of
    the resident chunks when `cos_sin` is given, else random), resident [B,H,S] distinct chunk ids."""
    lay = layout(prompt_len, sparse_budget, chunk, uncached)
    H, D = shape.num_key_value_heads, shape.hidden_size // shape.num_attention_heads
    g = torch.Generator(device=device).manual_seed(seed)

    def rn(*s, scale=1.0):
        return (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * scale).to(BF16)

    U = rn(batch, prompt_len, rank, scale=1.0 / math.sqrt(rank))
    SV = rn(batch, H, D, rank)
    lm = rn(batch, H, lay.n_landmarks, D)
    v_chunks = rn(batch, H, lay.all_chunks, chunk * D)
    # landmark index: ascending complement of random outlier chunks
    perm = torch.stack([torch.stack([torch.randperm(lay.chunks, generator=g, device=device) for _ in range(H)])
                        for _ in range(batch)])
    lm_idx = perm[..., lay.outlier_chunk:].sort(dim=-1).values.contiguous()
    # initial resident set: S distinct landmark chunks
    pick = torch.stack([torch.stack([torch.randperm(lay.n_landmarks, generator=g, device=device)[:lay.select]
                                     for _ in range(H)]) for _ in range(batch)])
    resident = lm_idx.gather(-1, pick).contiguous()
    k_buf = torch.zeros(batch, H, lay.buf_len, D, dtype=BF16, device=device)
    v_buf = torch.zeros(batch, H, lay.buf_len, D, dtype=BF16, device=device)
    k_buf[:, :, :lay.sparse_start] = rn(batch, H, lay.sparse_start, D)
    v_buf[:, :, :lay.sparse_start] = rn(batch, H, lay.sparse_start, D)
    # sparse region: V = the resident chunks of the store; K = arbitrary per-chunk content (it is what a
    # previous step left there; the tests only need it to be a known function of the chunk)
    idx = resident.unsqueeze(-1).expand(-1, -1, -1, chunk * D)
    v_buf[:, :, lay.sparse_start:lay.sparse_end] = v_chunks.gather(2, idx).view(batch, H, -1, D)
    if cos_sin is None:
        k_buf[:, :, lay.sparse_start:lay.sparse_end] = rn(batch, H, lay.select * chunk, D)
    else:   # what earlier steps would have left there: RoPE(U.SV) of the resident chunks
        tok = (resident.unsqueeze(-1) * chunk + torch.arange(chunk, device=device)).view(batch, H, -1)
        for b in range(batch):
            for h in range(H):
                kpre = (U[b, tok[b, h]].float() @ SV[b, h].float().t()).to(BF16).float()
                k_buf[b, h, lay.sparse_start:lay.sparse_end] = _rope_float(kpre, cos_sin.to(device)[tok[b, h]].float()).to(BF16)
    out = dict(U=U, SV=SV, k_landmark=lm, k_landmark_idx=lm_idx, v_chunks=v_chunks, k_buf=k_buf, v_buf=v_buf,
               resident=resident, layout=lay)
    return out


def queries(shape: ModelShape, batch: int, *, seed: int, device="cpu", base=None, noise: float = 1.0):
    """Post-RoPE q [B,Hq,1,D], k_new / v_new [B,H,1,D].  With `base` given, q = base*sqrt(1-noise^2) +
    noise*N(0,1): the knob that sets temporal locality (chunk-cache hit rate) in the bench."""
    D = shape.hidden_size // shape.num_attention_heads
    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.randn(batch, shape.num_attention_heads, 1, D, generator=g, device=device)
    if base is not None:
        q = base.float() * math.sqrt(max(0.0, 1.0 - noise * noise)) + noise * q
    k = torch.randn(batch, shape.num_key_value_heads, 1, D, generator=g, device=device)
    v = torch.randn(batch, shape.num_key_value_heads, 1, D, generator=g, device=device)
    return q.to(BF16), k.to(BF16), v.to(BF16)


def install(cache, layer_states, cos_sin=None, skip_v_store: bool = False):
    """Load per-layer synthetic states into a ShadowKVCache_CPU (stacking along the layer dim).
    skip_v_store: the caller already copied every layer's `v_chunks` into cache.v_cache_cpu."""
    dev = cache.device
    L = len(layer_states)
    lay = layer_states[0]["layout"]
    prompt_len = layer_states[0]["U"].shape[1]

    def stack(name):
        return torch.stack([s[name].to(dev) for s in layer_states])

    for l, s in enumerate(layer_states):
        if not skip_v_store:
            nchunks = s["v_chunks"].shape[2]
            cache.v_cache_cpu[l][:, :, :nchunks].copy_(s["v_chunks"])
        cache.k_cache_buffer[l].copy_(s["k_buf"])
        cache.v_cache_buffer[l].copy_(s["v_buf"])
    cache.load_state(prompt_len=prompt_len, U=stack("U"), SV=stack("SV"), k_landmark=stack("k_landmark"),
                     k_landmark_idx=stack("k_landmark_idx"), position_ids=stack("resident"))
    assert cache.sparse_start == lay.sparse_start and cache.buf_len == lay.buf_len
    return cache
