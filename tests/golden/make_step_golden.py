"""Generates tests/golden/step_golden.npz: whole decode steps of the REFERENCE's own CUDA path -- the unmodified
extension built by oracle/build_ref.sh + flash_attn_with_kvcache of the flash-attn wheel, driven exactly as
models/kv_cache.py:983-1176 / models/base.py:320-341 drive them (tests/refchain.py) -- at the BASELINE.json shapes
(tests/golden/step_cases.py).  Run on the GPU box:

    python tests/golden/make_step_golden.py gpurun_out/step_golden.npz

Per case and step the fixture keeps what a different implementation must reproduce: the resident chunk SET after the
step (sorted; slot order is an implementation choice), the hit counts, one K row of every 8th resident chunk (hits
carry the given state, misses are the reference's gather-GEMM + RoPE kernels' output), a permutation-invariant
checksum of the V bytes in the sparse region, and the attention output of flash_attn_with_kvcache."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.golden import step_cases as SC  # noqa: E402
from tests.golden.inputs import bits  # noqa: E402
from tests.refchain import ReferenceLayer, load_reference_extension  # noqa: E402


def sample_k_rows(k_buf, resident, ss):
    """k_buf [B,H,Lbuf,D], resident [B,H,S] (chunk id per slot) -> bf16 rows [B,H,S/8,D]: for every 8th chunk of the
    ascending resident set, row (id % 8) of that chunk."""
    B, H, S = resident.shape
    srt, slot = resident.sort(dim=-1)
    srt, slot = srt[..., ::SC.SAMPLE_EVERY], slot[..., ::SC.SAMPLE_EVERY]
    rows = ss + slot * 8 + (srt % 8)
    return k_buf.gather(2, rows.unsqueeze(-1).expand(-1, -1, -1, k_buf.shape[-1]))


def v_checksum(v_buf, ss, se):
    return v_buf[:, :, ss:se].contiguous().view(torch.int16).long().sum(dim=(2, 3))


def main(out_path):
    ref, flash = load_reference_extension()
    dev = torch.device("cuda:0")
    out = {}
    for name in SC.CASES:
        shape, B, ctx, cs, st = SC.build(name)
        lay = st["layout"]
        layer = ReferenceLayer(st, ctx, SC.BUDGET, dev)
        d_cs = cs.to(dev)
        for t, (q, k, v) in enumerate(SC.step_inputs(name, shape, B)):
            o = layer.step(ref, flash, q.to(dev), k.to(dev), v.to(dev), t, d_cs)
            torch.cuda.synchronize()
            assert int(layer.signals.abs().sum()) == 0
            res = layer.position_ids
            out[f"{name}/resident_sorted/{t}"] = res.sort(dim=-1).values.to(torch.int32).cpu().numpy()
            out[f"{name}/cnts/{t}"] = layer.cnts.cpu().numpy()
            out[f"{name}/k_rows/{t}"] = bits(sample_k_rows(layer.k_buf, res, lay.sparse_start))
            out[f"{name}/v_sum/{t}"] = v_checksum(layer.v_buf, lay.sparse_start, lay.sparse_end).cpu().numpy()
            out[f"{name}/attn/{t}"] = bits(o)
            print(name, "step", t, "hits", int(layer.cnts.sum()), "of", layer.cnts.numel() * lay.select, flush=True)
        del layer, st
        torch.cuda.empty_cache()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "step_golden.npz"))
