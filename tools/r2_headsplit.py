#!/usr/bin/env python
"""2+ GPU check of the kv-head split with the peer-store exchange (run under torchrun on one node):
every rank owns H/N kv heads of the same sequence; after every token the gathered buffer of EVERY rank must hold the
attention outputs of ALL ranks, equal to what a one-GPU run of all heads produces (rank 0 checks against its own
full-head cache), for both exchange kinds.  Prints one line per rank."""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from shadowkv_b200 import parallel, synth  # noqa: E402
from shadowkv_b200.engine import DecodeStepGraph  # noqa: E402
from shadowkv_b200.kv_cache import ShadowKVCache_CPU  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    full = synth.MODELS["llama-3.1-8b"]
    B, ctx, budget, L = 1, 16384, 2048, 4
    cs = synth.cos_sin_table(ctx + 1024, device=dev)
    states = [synth.layer_state(full, B, ctx, seed=500 + l, sparse_budget=budget, cos_sin=cs.cpu()) for l in range(L)]
    cfg = parallel.head_shard_config(full.config(L), world, rank)
    h0, hl = cfg.kv_head_start, cfg.num_key_value_heads
    G = full.num_attention_heads // full.num_key_value_heads

    def shard(st):
        s = dict(st)
        for k in ("SV", "k_landmark", "k_landmark_idx", "v_chunks", "k_buf", "v_buf", "resident"):
            s[k] = st[k][:, h0:h0 + hl].contiguous()
        return s

    def build(shape, sts):
        c = ShadowKVCache_CPU(shape.config(L), batch_size=B, max_length=ctx, device=str(dev), sparse_budget=budget)
        return synth.install(c, sts)

    local_shape = dataclasses.replace(full, num_key_value_heads=hl, num_attention_heads=hl * G, hidden_size=hl * G * 128)
    worst = 0.0
    for kind in ("p2p", "nccl"):
        os.environ["SKV_HEAD_EXCHANGE"] = kind
        cache = build(local_shape, [shard(s) for s in states])
        ref = build(full, states) if rank == 0 else None
        run = DecodeStepGraph(cache, cs, host_io=False, gather_heads=True).capture(warmup=0)
        g = torch.Generator().manual_seed(9)
        for t in range(4):
            q = torch.randn(L, B, full.num_attention_heads, 1, 128, generator=g).to(torch.bfloat16).to(dev)
            k = torch.randn(L, B, full.num_key_value_heads, 1, 128, generator=g).to(torch.bfloat16).to(dev)
            run.step_device(q[:, :, h0 * G:(h0 + hl) * G].contiguous(), k[:, :, h0:h0 + hl].contiguous(),
                            k[:, :, h0:h0 + hl].contiguous())
            run.stream.synchronize()
            dist.barrier()
            got = torch.stack([parallel.heads_to_model_order(run.out_all[l]) for l in range(L)])      # [L,B,1,Hq,D]
            # every rank holds the same gathered tensor
            ref_t = got.clone()
            dist.broadcast(ref_t, src=0)
            assert torch.equal(ref_t, got), f"rank {rank}: gathered outputs differ from rank 0's ({kind}, token {t})"
            if rank == 0:
                for l in range(L):
                    want = ref.decode_layer(l, q[l], k[l], k[l], cs)
                    torch.cuda.synchronize()
                    err = (want.float() - got[l].float()).abs().max().item()
                    worst = max(worst, err)
                    assert err <= 4e-3, f"{kind} token {t} layer {l}: max-abs {err}"
        run.graph.reset()
        dist.barrier()
        print(f"rank {rank}/{world}: {kind} exchange ok ({run.exchange[:40]}...), worst max-abs vs one-GPU run {worst:.2e}", flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
