"""Generates tests/golden/ref_golden.npz by running the UNMODIFIED reference extension (built by
oracle/build_ref.sh into oracle/_ref) on a B200.  Run on the GPU box:

    python tests/golden/make_ref_golden.py gpurun_out/ref_golden.npz

The fixture pins the oracle (CPU test) and the CUDA path (GPU test) to what the reference's own kernels
produce on sm_100a for small deterministic inputs (tests/golden/inputs.py)."""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "kernels"))
from tests.golden import inputs as I  # noqa: E402


def cases():
    """Input builders shared with the tests (pure functions of constants)."""
    c = {}
    # reorder: 8 rows, S=256, mixture of hit rates
    S, rows, universe = 256, 8, 16384
    cached = I.distinct_ids(rows, S, universe, seed=1)
    cur = cached.copy()
    for r in range(rows):
        nmiss = [0, 256, 1, 255, 100, 102, 37, 200][r]
        fresh = I.distinct_ids(1, 2 * S, universe, seed=50 + r)[0]
        fresh = fresh[~np.isin(fresh, cached[r])][:nmiss]
        pos = (np.arange(nmiss) * 5 + r) % S if nmiss < S else np.arange(S)
        pos = np.unique(pos)[:len(fresh)]
        cur[r, pos] = fresh[:len(pos)]
        cur[r] = cur[r][I.distinct_ids(1, S, S, seed=90 + r)[0]]
    c["reorder"] = dict(cached=cached, cur=cur)
    return c


def main(out_path):
    import shadowkv as ref   # the reference extension
    dev = torch.device("cuda:0")
    out = {}
    cs_ = cases()

    # ---- reorder -----------------------------------------------------------------------------------
    cached = torch.from_numpy(cs_["reorder"]["cached"].copy()).to(dev)
    cur = torch.from_numpy(cs_["reorder"]["cur"].copy()).to(dev)
    rows, S = cached.shape
    off = torch.zeros(rows * S, dtype=torch.int32, device=dev)
    cnt = torch.zeros(rows, dtype=torch.int32, device=dev)
    ref.reorder_keys_and_compute_offsets(cached, cur, off, cnt, rows, 1, S)
    torch.cuda.synchronize()
    out["reorder_new"] = cached.cpu().numpy()
    out["reorder_off"] = off.cpu().numpy().reshape(rows, S)
    out["reorder_cnt"] = cnt.cpu().numpy()

    # ---- cached gathers (K d2d, V d2d+h2d) -----------------------------------------------------------
    B, H, nchunks, ss, Lbuf = 2, 4, 2048, 64, 64 + 256 * 8 + 32
    host = I.bf16_tensor((B * H, nchunks, 1024), seed=3).pin_memory()
    kbuf = I.bf16_tensor((B * H, Lbuf, 128), seed=4).to(dev)
    vbuf = I.bf16_tensor((B * H, Lbuf, 128), seed=5).to(dev)
    # offsets for the gather must index the host store: use ids < nchunks
    cached2 = torch.from_numpy(cs_["reorder"]["cached"] % nchunks)
    cur2 = torch.from_numpy(cs_["reorder"]["cur"] % nchunks)
    # make rows distinct-valued again after the modulo
    for r in range(rows):
        cached2[r] = torch.from_numpy(I.distinct_ids(1, S, nchunks, seed=200 + r)[0])
        keep = [256, 0, 255, 1, 154, 100, 219, 56][r]
        fresh = torch.from_numpy(I.distinct_ids(1, 3 * S, nchunks, seed=300 + r)[0])
        fresh = fresh[~torch.isin(fresh, cached2[r])][:S - keep]
        cur2[r] = torch.cat([cached2[r][:keep], fresh])[torch.from_numpy(I.distinct_ids(1, S, S, seed=400 + r)[0])]
    out["gather_cached"] = cached2.numpy().copy()
    out["gather_cur"] = cur2.numpy().copy()
    d_cached, d_cur = cached2.to(dev), cur2.to(dev)
    ref.reorder_keys_and_compute_offsets(d_cached, d_cur, off, cnt, B, H, S)
    temp = torch.zeros(B * H, S, 1024, dtype=torch.bfloat16, device=dev)
    signals = torch.zeros(B * H, dtype=torch.int32, device=dev)
    ref.gather_copy_d2d_with_offsets(kbuf, off, cnt, B, H, S * 1024, ss * 128, Lbuf * 128, S)
    ref.gather_copy_with_offsets(host, vbuf, temp, off, cnt, signals, B, H, nchunks * 1024, S * 1024, ss * 128,
                                 Lbuf * 128, S)
    torch.cuda.synchronize()
    out["gather_new"] = d_cached.cpu().numpy()
    out["gather_off"] = off.cpu().numpy().reshape(B * H, S)
    out["gather_cnt"] = cnt.cpu().numpy()
    out["gather_k_rows"] = I.bits(kbuf[:, ::8])           # every 8th row (one row per chunk) keeps it small
    out["gather_v_rows"] = I.bits(vbuf[:, ::8])
    out["gather_k_sum"] = kbuf.view(torch.int16).long().sum(dim=(1, 2)).cpu().numpy()
    out["gather_v_sum"] = vbuf.view(torch.int16).long().sum(dim=(1, 2)).cpu().numpy()
    out["gather_signals"] = signals.cpu().numpy()

    # ---- RoPE ---------------------------------------------------------------------------------------
    from shadowkv_b200 import synth
    Br, Hr, Sr = 2, 2, 4                      # chunks per head
    x = I.bf16_tensor((Br, Hr, Sr * 8, 128), seed=6, exp_base=0x7D).to(dev)
    ids = torch.from_numpy(I.distinct_ids(Br * Hr, Sr, 500, seed=7)).view(Br, Hr, Sr).to(torch.int32).to(dev)
    rcnt = torch.tensor([0, 1, 3, 2], dtype=torch.int32, device=dev)
    for name, glm in (("neox", False), ("glm", True)):
        cs = synth.cos_sin_table(4096 + 1024, glm=glm).to(dev)
        cache = torch.zeros(Br, Hr, 16 + Sr * 8, 128, dtype=torch.bfloat16, device=dev)
        fn = ref.apply_rotary_pos_emb_push_cache_opt_glm if glm else ref.apply_rotary_pos_emb_push_cache_opt
        fn(x, cs, ids, cache, rcnt, Br, Hr, Sr * 8, 128, *x.stride(), cs.stride(0), *ids.stride(), *cache.stride()[:3],
           16, 16 + Sr * 8, 64, 8)
        torch.cuda.synchronize()
        out[f"rope_push_{name}"] = I.bits(cache)
    cs = synth.cos_sin_table(4096 + 1024).to(dev)
    tok = (ids.long().unsqueeze(-1) * 8 + torch.arange(8, device=dev)).view(Br, Hr, Sr * 8)
    o = torch.empty_like(x)
    ref.apply_rotary_pos_emb_new(x, cs, tok, o, Br, Hr, Sr * 8, 128, *x.stride(), cs.stride(0), *tok.stride(), 64)
    torch.cuda.synchronize()
    out["rope_new"] = I.bits(o)

    # ---- landmark scoring ---------------------------------------------------------------------------
    bc, m, n = 4, 4, 2048
    A = I.bf16_tensor((bc, m, 128), seed=8, exp_base=0x7D).to(dev)
    Bm = I.bf16_tensor((bc, n, 128), seed=9, exp_base=0x7D).to(dev)
    D = torch.zeros(bc, m, n, dtype=torch.bfloat16, device=dev)
    Sm = torch.zeros_like(D)
    T = (n + 255) // 256
    norm = torch.zeros(bc, m, T, device=dev)
    ssum = torch.zeros(bc, m, T, device=dev)
    ref.batch_gemm_softmax(A, Bm, D, norm, ssum, Sm, bc, m, n, 128, 1 / math.sqrt(128), 0)
    torch.cuda.synchronize()
    out["score_D"] = I.bits(D)
    out["score_softmax"] = I.bits(Sm)

    # ---- K rebuild: batch_gather_gemm + RoPE push ------------------------------------------------------
    Bg, Hg, Lg, R, Sg = 1, 2, 4096, 160, 128
    U = I.bf16_tensor((Bg, Lg, R), seed=10, exp_base=0x79).to(dev)
    SV = I.bf16_tensor((Bg, Hg, 128, R), seed=11, exp_base=0x7D).to(dev)
    gid = torch.from_numpy(I.distinct_ids(Bg * Hg, Sg, Lg // 8, seed=12)).view(Bg, Hg, Sg).to(torch.int32).to(dev)
    gcnt = torch.tensor([0, 45], dtype=torch.int32, device=dev)
    scratch = torch.zeros(Bg, Hg, Sg * 8, 128, dtype=torch.bfloat16, device=dev)
    cs = synth.cos_sin_table(Lg + 1024).to(dev)
    ref.batch_gather_gemm(U, SV, cs, cs, gid, scratch, Bg, Hg, Lg, 128, R, Sg * 8, Lg + 1024, 8, gcnt)
    torch.cuda.synchronize()
    kc = torch.zeros(Bg, Hg, 32 + Sg * 8, 128, dtype=torch.bfloat16, device=dev)
    ref.apply_rotary_pos_emb_push_cache_opt(scratch, cs, gid, kc, gcnt, Bg, Hg, Sg * 8, 128, *scratch.stride(),
                                            cs.stride(0), *gid.stride(), *kc.stride()[:3], 32, 32 + Sg * 8, 64, 8)
    torch.cuda.synchronize()
    out["rebuild_pre_rows"] = I.bits(scratch[:, :, ::4])
    out["rebuild_rope_rows"] = I.bits(kc[:, :, ::4])

    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_golden.npz"))
