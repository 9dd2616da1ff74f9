"""Debug aid for the kv-head split under torchrun (2 ranks): eager all-gather, then graph capture, with progress."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dataclasses
import torch, torch.distributed as dist
from shadowkv_b200 import synth, parallel
from shadowkv_b200.engine import DecodeStepGraph
from shadowkv_b200.kv_cache import ShadowKVCache_CPU

def log(*a):
    print(f"[rank {os.environ.get('RANK')}] ", *a, file=sys.stderr, flush=True)

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
full = synth.MODELS["llama-3.1-8b"]
cfg = parallel.head_shard_config(full.config(1), world, rank)
shape = dataclasses.replace(full, num_key_value_heads=cfg.num_key_value_heads, num_attention_heads=cfg.num_attention_heads, hidden_size=cfg.hidden_size)
L, B, ctx = 2, 1, 8192
cs = synth.cos_sin_table(ctx + 1024, device=dev)
cache = ShadowKVCache_CPU(shape.config(L), batch_size=B, max_length=ctx, device=str(dev), sparse_budget=1024)
synth.install(cache, [synth.layer_state(shape, B, ctx, seed=5 + l, sparse_budget=1024, device=dev, cos_sin=cs) for l in range(L)])
log("state ready")
x = torch.ones(4, device=dev) * rank
g = torch.empty(world * 4, device=dev)
dist.all_gather_into_tensor(g, x); torch.cuda.synchronize(); log("eager all_gather ok", g.tolist())
run = DecodeStepGraph(cache, cs, host_io=False, gather_heads=True)
log("runner built")
run.capture(warmup=1)
torch.cuda.synchronize(); log("captured")
q = torch.randn(L, B, shape.num_attention_heads, 1, 128, device=dev).to(torch.bfloat16)
k = torch.randn(L, B, shape.num_key_value_heads, 1, 128, device=dev).to(torch.bfloat16)
for t in range(3):
    run.step_device(q, k, k); run.stream.synchronize(); log("replay", t, "ok")
dist.barrier(); log("barrier ok")
dist.destroy_process_group(); log("done")
