"""Host-side logic of the N > 1 paths on CPU: world_size-2 gloo processes (no GPU needed).

* batch sharding: the shards partition the batch, the reported time is the slowest rank's.
* kv-head split: every rank runs the (oracle) decode step on its kv heads only; the all-gathered attention
  outputs, put back in model head order, equal the unsharded result bit for bit -- the per-(b,h) independence
  the GPU path relies on."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import shadowkv_oracle as O
from shadowkv_b200 import parallel, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _state(B, prompt_len, budget, seed):
    shape = synth.MODELS["llama-3.1-8b"]
    cs = synth.cos_sin_table(prompt_len + 1024)
    st = synth.layer_state(shape, B, prompt_len, seed=seed, sparse_budget=budget, cos_sin=cs)
    return shape, cs, st


def _oracle_step(st, lay, cs, q, k, v, heads=slice(None), qheads=slice(None)):
    ls = O.LayerState(landmarks=st["k_landmark"][:, heads], landmark_idx=st["k_landmark_idx"][:, heads], U=st["U"],
                      SV=st["SV"][:, heads], v_host=st["v_chunks"][:, heads], k_buf=st["k_buf"][:, heads].clone(),
                      v_buf=st["v_buf"][:, heads].clone(), resident=st["resident"][:, heads].clone())
    geo = O.Geometry(sparse_start=lay.sparse_start, sparse_end=lay.sparse_end)
    out, info = O.decode_step(ls, geo, q[:, qheads], k[:, heads], v[:, heads], cs)
    return out, ls


def _worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- batch sharding -------------------------------------------------------------------------
        total = 5
        start, count = parallel.shard_range(total, world, rank)
        got = [None] * world
        dist.all_gather_object(got, (start, count))
        assert sorted(got) == [(0, 3), (3, 2)]
        assert parallel.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        assert parallel.sum_over_ranks(count) == total
        # ---- kv-head split ----------------------------------------------------------------------------
        B, prompt_len, budget = 1, 4096, 1024
        shape, cs, st = _state(B, prompt_len, budget, seed=77)          # same seed on every rank = replicated inputs
        lay = st["layout"]
        cfg = parallel.head_shard_config(shape.config(1), world, rank)
        H, G = shape.num_key_value_heads, shape.num_attention_heads // shape.num_key_value_heads
        h0, hl = cfg.kv_head_start, cfg.num_key_value_heads
        assert hl == H // world and cfg.num_attention_heads == hl * G
        assert cfg.hidden_size // cfg.num_attention_heads == 128
        q, k, v = synth.queries(shape, B, seed=5)
        out_local, ls = _oracle_step(st, lay, cs, q, k, v, heads=slice(h0, h0 + hl), qheads=slice(h0 * G, (h0 + hl) * G))
        full = parallel.heads_to_model_order(parallel.gather_heads(out_local))
        if rank == 0:
            want, ls_full = _oracle_step(st, lay, cs, q, k, v)
            assert torch.equal(full, want), "head-split result differs from the unsharded step"
            assert torch.equal(ls.resident, ls_full.resident[:, h0:h0 + hl])
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_range_properties():
    for total in (0, 1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)
    with pytest.raises(ValueError):
        parallel.head_shard_config(synth.MODELS["llama-3.1-8b"].config(1), 3, 0)
