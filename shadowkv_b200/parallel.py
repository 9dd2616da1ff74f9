"""Multi-GPU partitioning of the sparse-KV decode step.  One process per GPU.

* batch sharding  -- sequences are independent (all state is indexed [layer, b, h, ...]): every rank owns
  a contiguous slice of the batch with its own decode state and pinned V store; NO collective on the data path.
* kv-head split   -- for batch-1 long contexts: every rank owns a contiguous slice of the kv heads (and their
  G query heads); scoring, selection, V fetch, K rebuild and attention are per-(b,h) independent, U is
  replicated; the only exchange is one all-gather of the attention output [B, 1, Hq/world, D] per layer.

The helpers work on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, BEFORE the pinned V store is
    allocated: pinned pages are placed by first touch, and the V fetch reads them over that GPU's PCIe link -- with
    several ranks on a multi-socket host a store on the wrong socket crosses the inter-socket fabric and the ranks
    contend for it.  Best effort: returns what it found ({"node": n, "cpus": k} or {"node": None, "why": ...})."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"node": None, "why": "device reports no NUMA node (single-node host)"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"node": node, "why": "no allowed CPU on that node (cgroup cpuset)"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except (OSError, ValueError, AttributeError) as e:
        return {"node": None, "why": f"{type(e).__name__}: {e}"}


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced partition of `total` items: returns (start, count) of `rank`."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def head_shard_config(config, world: int, rank: int):
    """Model config of one rank of the kv-head split: H/world kv heads with their query heads.  hidden_size is
    scaled so that head_dim = hidden_size / num_attention_heads is unchanged."""
    H, Hq = config.num_key_value_heads, config.num_attention_heads
    if H % world:
        raise ValueError(f"{H} kv heads do not split over {world} ranks")
    G, D = Hq // H, config.hidden_size // Hq
    h_local = H // world
    return SimpleNamespace(num_hidden_layers=config.num_hidden_layers, hidden_size=h_local * G * D,
                           num_attention_heads=h_local * G, num_key_value_heads=h_local,
                           kv_head_start=rank * h_local)


def max_over_ranks(value: float, device="cpu") -> float:
    """Max of a host scalar over all ranks (elapsed times are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_heads(out_local: torch.Tensor, gathered: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather of the per-rank attention outputs [B, 1, Hq/world, D] (the one exchange of the kv-head split).
    Returns [world, B, 1, Hq/world, D]; `heads_to_model_order` turns it into [B, 1, Hq, D]."""
    world = dist.get_world_size()
    if gathered is None:
        gathered = torch.empty((world,) + tuple(out_local.shape), dtype=out_local.dtype, device=out_local.device)
    # the collective wants the ranks concatenated along dim 0; [world, B, ...] viewed as [world*B, ...] is that
    dist.all_gather_into_tensor(gathered.view((world * out_local.shape[0],) + tuple(out_local.shape[1:])),
                                out_local.contiguous())
    return gathered


def heads_to_model_order(gathered: torch.Tensor) -> torch.Tensor:
    """[world, B, 1, Hq/world, D] -> [B, 1, Hq, D] (rank r holds query heads [r*Hq/world, (r+1)*Hq/world))."""
    world, B, one, hq, D = gathered.shape
    return gathered.permute(1, 2, 0, 3, 4).reshape(B, one, world * hq, D)
