"""Step-level golden cases: whole decode steps of one layer at the BASELINE.json shapes.  Inputs are pure functions of
seeds drawn on the CPU generator (identical wherever torch is the same build), so the generator on the GPU box
(make_step_golden.py) and the tests rebuild the very same decode state and queries."""
from __future__ import annotations

from shadowkv_b200 import synth

STEPS = 3
SAMPLE_EVERY = 8        # K rows: one row of every 8th chunk of the sorted resident set is kept in the fixture
NOISE = 0.3

# name -> (model, batch, context, seed)
CASES = {
    "llama31_8b_122k_b1": ("llama-3.1-8b", 1, 124928, 4001),      # BASELINE configs[1] (the headline shape)
    "glm4_9b_128k_b1": ("glm-4-9b-1m", 1, 131072, 4002),          # configs[3] geometry: 4 kv x 8 q heads, GLM RoPE
    "yi_9b_200k_b1": ("yi-9b-200k", 1, 204800, 4003),             # configs[4] geometry: 4 kv x 8 q heads, 200K
    "llama31_8b_32k_b3": ("llama-3.1-8b", 3, 32768, 4004),        # batch > 2 through the whole step
}
BUDGET, RANK = 2048, 160


def build(name: str):
    """Returns (shape, B, ctx, cos_sin [CPU], layer state dict [CPU tensors])."""
    model, B, ctx, seed = CASES[name]
    shape = synth.MODELS[model]
    cs = synth.cos_sin_table(ctx + 1024, theta=shape.rope_theta, glm=shape.glm_rope)
    st = synth.layer_state(shape, B, ctx, seed=seed, rank=RANK, sparse_budget=BUDGET, cos_sin=cs)
    return shape, B, ctx, cs, st


def step_inputs(name: str, shape, B):
    """[(q, k_new, v_new)] for STEPS decode steps (CPU bf16 tensors)."""
    _, _, _, seed = CASES[name]
    out, base = [], None
    for t in range(STEPS):
        q, k, v = synth.queries(shape, B, seed=seed * 10 + t, base=base, noise=NOISE)
        base = q if base is None else base
        out.append((q, k, v))
    return out
