#!/usr/bin/env python
"""Launches each dense decode kernel of the Llama-3.1-8B layer once (batch 1) after a warm-up, for Nsight Compute:
    ncu --set full --clock-control none --import-source on -k regex:dense_streamk --launch-skip 10 --launch-count 5 \\
        -o gpurun_out/prof_dense python tools/profile_dense.py
Order of the profiled launches: qkv_rope, o_proj_residual, gate_up_silu_mul, down_residual, lm_head."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from shadowkv_b200 import dense, synth  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1


def w(*s):
    return (torch.randn(*s, generator=g, device=dev) * 0.02).to(torch.bfloat16)


hid, ffn, Hq, H, vocab = 4096, 14336, 32, 8, 128256
Wqkv, Wo, Wgu, Wdn, Wlm = w((Hq + 2 * H) * 128, hid), w(hid, hid), w(2 * ffn, hid), w(hid, ffn), w(vocab, hid)
x, xa, xf = w(B, hid) * 50, w(B, hid) * 15, w(B, ffn) * 15
res = torch.zeros(B, hid, dtype=torch.bfloat16, device=dev)
cs = synth.cos_sin_table(2048, device=dev)


def layer():
    dense.linear_qkv_rope(x, Wqkv, Hq, H, cs, 1000, static_weights=True)
    dense.linear(xa, Wo, residual=res, out=res, static_weights=True)
    dense.linear_silu_mul(x, Wgu, static_weights=True)
    dense.linear(xf, Wdn, residual=res, out=res, static_weights=True)
    dense.linear_f32(x, Wlm, static_weights=True)


for _ in range(2):
    layer()
torch.cuda.synchronize()
layer()
torch.cuda.synchronize()
print("done")
