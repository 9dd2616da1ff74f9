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


class HeadExchange:
    """Peer-store exchange of the kv-head split: every rank's gathered buffer
    ``out_all [layers, world, B, 1, Hq/world, D]`` and flag words ``flags [layers, world, rows]`` live in symmetric memory
    (``torch.distributed._symmetric_memory``: the same allocation mapped into every rank of the node); the CTA that
    merges a row inside ``skv_fetch_rebuild_attend`` stores the row's attention output into EVERY rank's buffer over
    NVLink and then raises that row's flag word everywhere; ``wait(layer)`` / ``wait_spec(layer)`` wait until all
    row words of the layer carry the current epoch.  No collective launch on the per-layer chain.
    world == 1 (or no process group): plain device tensors and self pointers -- the same kernel path, used by the tests."""

    def __init__(self, layers: int, batch: int, hq_local: int, device, head_dim: int = 128, group=None, rows: int = 1):
        """rows: local (b, kv head) rows of a layer = flag words per (layer, rank): B * H/world."""
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.world > 8:
            raise ValueError("the peer-store exchange handles up to 8 ranks of one node")
        self.layers, self.row_elems = layers, batch * hq_local * head_dim
        self.rows = int(rows)
        shape = (layers, self.world, batch, 1, hq_local, head_dim)
        if self.world == 1:
            self.out_all = torch.zeros(shape, dtype=torch.bfloat16, device=device)
            self.flags = torch.zeros(layers, 1, self.rows, dtype=torch.int32, device=device)
            self._out_ptrs, self._flag_ptrs = [self.out_all.data_ptr()], [self.flags.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm
            grp = group if group is not None else dist.group.WORLD
            try:
                symm.enable_symm_mem_for_group(grp.group_name)
            except Exception:  # noqa: BLE001  -- newer builds enable it implicitly
                pass
            self.out_all = symm.empty(shape, dtype=torch.bfloat16, device=device)
            self.flags = symm.empty((layers, self.world, self.rows), dtype=torch.int32, device=device)
            self.out_all.zero_()
            self.flags.zero_()
            torch.cuda.synchronize(device)
            h_out, h_flag = symm.rendezvous(self.out_all, grp), symm.rendezvous(self.flags, grp)
            self._out_ptrs, self._flag_ptrs = list(h_out.buffer_ptrs), list(h_flag.buffer_ptrs)
            self._handles = (h_out, h_flag)
            dist.barrier(grp)
            torch.cuda.synchronize(device)
        self.epoch = torch.ones(1, dtype=torch.int32, device=device)       # flags start at 0 < epoch 1

    def struct(self, layer: int):
        """skv_exchange for `layer`: where THIS rank's output of the layer goes on every peer, and its flag there."""
        from . import _lib
        ex = _lib.Exchange(world=self.world, rank=self.rank)
        slot = (layer * self.world + self.rank)
        for r in range(self.world):
            ex.peer_out[r] = self._out_ptrs[r] + slot * self.row_elems * 2
            ex.peer_flags[r] = self._flag_ptrs[r] + slot * self.rows * 4
        ex.epoch_dev = self.epoch.data_ptr()
        return ex

    def wait(self, layer: int):
        """Block the current stream until every rank's output of `layer` (current epoch) is in this rank's buffer."""
        from ._lib import check, lib
        check(lib.skv_wait_flags(self.flags.data_ptr() + layer * self.world * self.rows * 4, self.world * self.rows,
                                 self.epoch.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream), "wait_flags")

    def wait_spec(self, layer: int):
        """(flags pointer, count, epoch pointer) for ``decode_layer(wait_for=...)`` of the layer BEHIND `layer`: its scoring
        launch then waits for `layer`'s gathered output itself, with no launch of its own in between."""
        return (self.flags.data_ptr() + layer * self.world * self.rows * 4, self.world * self.rows, self.epoch.data_ptr())

    def advance(self):
        """Next token: epoch += 1 on the device (part of the captured step; never rewound)."""
        self.epoch.add_(1)


class RowHandover:
    """One-GPU form of the hand-over ``HeadExchange`` does between ranks: the CTA that merges a (b, kv head) row inside
    ``skv_fetch_rebuild_attend`` raises that row's word (value = the token's epoch) behind the row's attention output,
    and the NEXT layer's scoring launch starts on the words of all rows (``SKV_GEOM_WAIT_FLAGS_ONLY``) instead of on the
    completion and flush of the whole grid; it executes its wait for that completion at its end, so every later kernel
    keeps the transitive guarantee.  ``flags [layers, rows]`` int32, ``epoch`` a device counter advanced once per token."""

    def __init__(self, layers: int, rows: int, device):
        self.layers, self.rows = int(layers), int(rows)
        self.flags = torch.zeros(self.layers, self.rows, dtype=torch.int32, device=device)
        self.epoch = torch.ones(1, dtype=torch.int32, device=device)       # flags start at 0 < epoch 1

    def struct(self, layer: int):
        from . import _lib
        ex = _lib.Exchange(world=1, rank=0)
        ex.peer_flags[0] = self.flags.data_ptr() + layer * self.rows * 4     # peer_out[0] stays NULL: no copy of the row
        ex.epoch_dev = self.epoch.data_ptr()
        return ex

    def wait_spec(self, layer: int):
        return (self.flags.data_ptr() + layer * self.rows * 4, self.rows, self.epoch.data_ptr())

    def advance(self):
        self.epoch.add_(1)

