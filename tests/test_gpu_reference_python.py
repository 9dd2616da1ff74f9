"""INTEGRATION.md level 1 as an executed claim: the reference's OWN, UNMODIFIED Python for the hot path
(models/kv_cache.py ShadowKVCache_CPU :509-1319 + models/tensor_op.py :154-238, copied verbatim into the git-ignored
oracle/_ref/models by oracle/build_ref.sh) runs on top of ``shadowkv_b200.shadowkv`` -- the drop-in for
``kernels.shadowkv`` -- and ``shadowkv_b200.attention.flash_attn_with_kvcache``, driven with the call sequence of
models/base.py:296-341 (prefill: get_svd + prefill_kv_cache + H2D; decode: update_kv_cache ->
get_retrieval_position_ids -> {get_value_cache on copy_stream || get_key_cache} -> flash_attn_with_kvcache).

Checked, per decode step and layer:
  (1) against the oracle (CPU restatement) fed with the state the reference object built;
  (2) against the SAME reference Python running on the reference's own CUDA extension + the flash-attn wheel
      (oracle/_ref/kernels), i.e. the reference's CUDA path, on identical inputs: chunk tables, hit counts, V bytes
      bit for bit (modulo top-k ties), K rows and attention output within the bf16 tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests import refpy
from tests.tolerances import MAX_ABS, assert_close_bf16, assert_rope_rows_close

pytestmark = pytest.mark.gpu


def _prompt(shape, B, L, dev, seed):
    """A synthetic prompt: pre-RoPE keys that are close to rank 96 (so the rank-160 factorisation reproduces them),
    values, the post-RoPE keys and the prompt's last query."""
    H, D = shape.num_key_value_heads, 128
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(B, L, 96, generator=g) / 96 ** 0.5
    m = torch.randn(B, 96, H * D, generator=g)
    k_pre = (a @ m + 0.02 * torch.randn(B, L, H * D, generator=g)).to(torch.bfloat16)
    k_pre = k_pre.view(B, L, H, D).transpose(1, 2).contiguous()                  # [B,H,L,D]
    v = torch.randn(B, H, L, D, generator=g).to(torch.bfloat16)
    cs = synth.cos_sin_table(L + 1024, theta=shape.rope_theta, glm=shape.glm_rope)
    k_roped = synth._rope_float(k_pre.float(), cs[:L].float()[None, None]).to(torch.bfloat16)
    last_q = torch.randn(B, shape.num_attention_heads, 1, D, generator=g).to(torch.bfloat16)
    return k_pre.to(dev), v.to(dev), k_roped.to(dev), last_q.to(dev), cs


def _build(mod, shape, B, L, budget, layers, prompt, dev):
    k_pre, v, k_roped, last_q, _ = prompt
    cache = mod.ShadowKVCache_CPU(shape.config(layers), batch_size=B, max_length=L, device=str(dev),
                                  sparse_budget=budget, chunk_size=8, rank=160)
    for l in range(layers):                                                       # base.py:296-303
        cache.get_svd(k_pre, layer_idx=l)
        cache.prefill_kv_cache(v, l, k_roped, last_q)
    cache.H2D()                                                                   # base.py:434
    return cache


def _oracle_state(cache, l):
    return O.LayerState(landmarks=cache.k_landmark[l].cpu(), landmark_idx=cache.k_landmark_idx[l].cpu(),
                        U=cache.U[l].cpu(), SV=cache.SV[l].cpu(), v_host=cache.v_cache_cpu[l].clone(),
                        k_buf=cache.k_cache_buffer[l].cpu().clone(), v_buf=cache.v_cache_buffer[l].cpu().clone(),
                        resident=cache.position_ids[l].cpu().clone())


@pytest.mark.skipif(not refpy.available("b200"), reason="oracle/_ref/models not built (oracle/build_ref.sh)")
@pytest.mark.parametrize("model,B,L,budget", [("llama-3.1-8b", 2, 8232, 1024), ("glm-4-9b-1m", 1, 16384, 2048)])
def test_unmodified_reference_python_runs_on_the_shim(cuda, model, B, L, budget):
    shape = synth.MODELS[model]
    layers, steps = 2, 4
    H, S = shape.num_key_value_heads, budget // 8
    prompt = _prompt(shape, B, L, cuda, seed=11)
    cs = prompt[4]
    d_cs = cs.to(cuda)

    mod_a, _, attn_a = refpy.load("b200")
    assert mod_a.shadowkv.__name__ == "shadowkv_b200.shadowkv"
    ca = _build(mod_a, shape, B, L, budget, layers, prompt, cuda)
    have_ref = refpy.available("reference")
    if have_ref:
        mod_b, _, attn_b = refpy.load("reference")
        assert mod_b.shadowkv is not mod_a.shadowkv
        cb = _build(mod_b, shape, B, L, budget, layers, prompt, cuda)
        # prefill is torch-only and identical code; pin the factors anyway (cuSOLVER is free to differ run to run)
        cb.U.copy_(ca.U); cb.SV.copy_(ca.SV)
        for name in ("k_landmark", "k_landmark_idx", "position_ids", "k_cache_buffer", "v_cache_buffer"):
            assert torch.equal(getattr(ca, name), getattr(cb, name)), f"prefill state {name} differs between the two objects"
    ora = [_oracle_state(ca, l) for l in range(layers)]
    geo = O.Geometry(sparse_start=ca.sparse_start, sparse_end=ca.sparse_end)
    assert ca.sparse_end - ca.sparse_start == budget

    base, worst_sel, worst_ref_sel, in_sync = None, 0, 0, [True] * layers
    for step in range(steps):
        geo.gen_offset = step
        for l in range(layers):
            q, kn, vn = synth.queries(shape, B, seed=31 * step + l, base=base, noise=0.35)
            base = q if base is None else base
            dq, dk, dv = q.to(cuda), kn.to(cuda), vn.to(cuda)
            out_a = refpy.decode_layer_like_base_py(ca, attn_a, l, dq, dk, dv, d_cs)
            torch.cuda.synchronize()
            # ---- (1) the oracle, forced onto the selection the reference Python made on our kernels' softmax -------
            res_a = ca.position_ids[l].cpu()
            want, info = O.decode_step(ora[l], geo, q, kn, vn, cs, selected_override=res_a)
            for b in range(B):
                for h in range(H):
                    got, own = set(res_a[b, h].tolist()), set(info["own_selected"][b, h].tolist())
                    assert len(got) == S
                    worst_sel = max(worst_sel, len(got ^ own) // 2)
            np.testing.assert_array_equal(ca.cnts.cpu().numpy(), info["cnts"])
            np.testing.assert_array_equal(ca.offsets.cpu().numpy().reshape(B * H, S), info["offsets"])
            assert torch.equal(res_a, ora[l].resident)
            L_att = ca.sparse_end + step + 1
            assert torch.equal(ca.v_cache_buffer[l].cpu()[:, :, :L_att], ora[l].v_buf[:, :, :L_att]), "V bytes vs oracle"
            assert_rope_rows_close(ca.k_cache_buffer[l].cpu()[:, :, :L_att], ora[l].k_buf[:, :, :L_att],
                                   glm=shape.glm_rope, frac_exact=0.95, what="K rows vs oracle")
            assert_close_bf16(out_a, want, what=f"attention vs oracle, step {step} layer {l}")
            ora[l].k_buf.copy_(ca.k_cache_buffer[l].cpu())
            # ---- (2) the same Python on the reference's own CUDA extension + flash-attn -------------------------
            if have_ref:
                out_b = refpy.decode_layer_like_base_py(cb, attn_b, l, dq, dk, dv, d_cs)
                torch.cuda.synchronize()
                assert int(cb.signals.abs().sum()) == 0
                res_b = cb.position_ids[l].cpu()
                diff = max(len(set(res_a[b, h].tolist()) ^ set(res_b[b, h].tolist())) // 2
                           for b in range(B) for h in range(H))
                worst_ref_sel = max(worst_ref_sel, diff)
                if diff == 0 and in_sync[l]:
                    # identical selections on identical tables: everything downstream is comparable slot for slot
                    assert torch.equal(res_a, res_b), "chunk table differs from the reference CUDA path"
                    assert torch.equal(ca.cnts, cb.cnts) and torch.equal(ca.offsets, cb.offsets)
                    assert torch.equal(ca.v_cache_buffer[l][:, :, :L_att], cb.v_cache_buffer[l][:, :, :L_att]), \
                        "V buffer bytes differ from the reference CUDA path"
                    assert_rope_rows_close(ca.k_cache_buffer[l].cpu()[:, :, :L_att], cb.k_cache_buffer[l].cpu()[:, :, :L_att],
                                           glm=shape.glm_rope, frac_exact=0.95, what="K rows vs the reference CUDA path")
                    cb.k_cache_buffer[l].copy_(ca.k_cache_buffer[l])      # keep 1-ulp rebuild differences from piling up
                else:
                    in_sync[l] = False                                   # a top-k tie went the other way: sets drift
                assert_close_bf16(out_a, out_b, max_abs=MAX_ABS, what=f"attention vs reference CUDA, step {step} layer {l}")
    print(f"\n[{model}] selection vs oracle: worst {worst_sel} chunks of {S}; vs reference CUDA path: worst "
          f"{worst_ref_sel} (top-k ties); layers in lock-step to the end: {in_sync}")
    assert worst_sel <= 2, "selection differs from the oracle in more than 2 chunks"
    if have_ref:
        assert worst_ref_sel <= 2
