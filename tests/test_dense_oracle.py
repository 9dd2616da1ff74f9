"""oracle/dense_oracle.py pinned against the LIBRARY kernels the reference's dense decode layer calls (flashinfer
rmsnorm, cuBLAS F.linear, vllm rotary_embedding / silu_and_mul), captured on a B200 by
tests/golden/make_dense_golden.py into tests/golden/dense_golden.npz.  CPU only.

Bars: element-wise kernels on identical inputs (rotary embedding, SiLU*mul, residual add) bit-exact; rmsnorm and linear
<= 1 bf16 ulp (their fp32 reduction order is the library's own)."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_oracle as DO
from tests.tolerances import assert_ulp_bf16

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "dense_golden.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(PATH)
    assert len(z["missing"]) == 0, f"fixture was generated without: {list(z['missing'])}"
    return z


def bf(a):
    return torch.from_numpy(np.ascontiguousarray(a)).view(torch.bfloat16)


def test_rmsnorm_matches_flashinfer(gold):
    got = DO.rmsnorm(bf(gold["x"]), bf(gold["norm_w"]), 1e-5)
    assert_ulp_bf16(got, bf(gold["rmsnorm"]), ulps=1, frac_exact=0.995, what="rmsnorm vs flashinfer")


def test_linear_matches_cublas(gold):
    x, W, b = bf(gold["lin_x"]), bf(gold["lin_W"]), bf(gold["lin_b"])
    assert_ulp_bf16(DO.linear(x, W), bf(gold["linear"]), ulps=1, frac_exact=0.98, what="F.linear")
    assert_ulp_bf16(DO.linear(x, W, b), bf(gold["linear_bias"]), ulps=1, frac_exact=0.98, what="F.linear + bias")
    # `.float()` of the bf16 result (base.py:157)
    assert np.array_equal(gold["logits_f32"], bf(gold["linear"]).float().numpy())


def test_residual_add_is_bit_exact(gold):
    assert torch.equal(DO.add_bf16(bf(gold["add_a"]), bf(gold["add_b"])), bf(gold["add"]))


@pytest.mark.parametrize("name,neox", [("neox", True), ("glm", False)])
def test_rotary_matches_vllm_bit_for_bit(gold, name, neox):
    row = bf(gold[f"cs_{name}_row"])
    for t in ("q", "k"):
        x = bf(gold[f"rope_{t}"])                      # [B, 1, heads*128]
        got = DO.rotary_vllm(x.view(x.shape[0], -1, 128), row, neox=neox).reshape(x.shape)
        assert torch.equal(got, bf(gold[f"rope_{t}_{name}"])), f"{t} {name}"


def test_silu_and_mul_matches_vllm(gold):
    got = DO.silu_and_mul(bf(gold["silu_in"]))
    want = bf(gold["silu_and_mul"])
    # expf on the device and torch.exp on the host may differ in the last fp32 bit: <= 1 bf16 ulp, almost all equal
    assert_ulp_bf16(got, want, ulps=1, frac_exact=0.999, what="silu_and_mul vs vllm")


def test_properties():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 8, 128, generator=g).to(torch.bfloat16)
    ang = torch.rand(64, generator=g) * 6.28
    row = torch.cat([ang.cos(), ang.sin()]).to(torch.bfloat16)
    y = DO.rotary_vllm(x, row, neox=True)
    # a rotation: pair norms are preserved up to bf16 rounding
    n0 = (x.float()[..., :64] ** 2 + x.float()[..., 64:] ** 2).sqrt()
    n1 = (y.float()[..., :64] ** 2 + y.float()[..., 64:] ** 2).sqrt()
    assert float(((n0 - n1).abs() / (n0 + 1e-3)).max()) < 0.03
    # GLM form leaves dims 64..127 alone
    z = DO.rotary_vllm(x, row[:64].contiguous(), neox=False)
    assert torch.equal(z[..., 64:], x[..., 64:])
    # rmsnorm output has unit mean square when w = 1
    h = DO.rmsnorm(x.reshape(32, 128), torch.ones(128, dtype=torch.bfloat16), 1e-6).float()
    assert float((h.pow(2).mean(-1) - 1).abs().max()) < 0.02
    # silu_and_mul against torch's own silu
    gu = torch.randn(3, 512, generator=g).to(torch.bfloat16)
    ref = torch.nn.functional.silu(gu[:, :256].float()) * gu[:, 256:].float()
    assert float((DO.silu_and_mul(gu).float() - ref).abs().max()) < 0.03


def test_prompt_path_rotary_embedding_equals_the_oracle_bit_for_bit():
    """shadowkv_b200.llama.rotary_embedding_rows (torch ops of the prompt path) against the
    oracle's restatement of vllm rotary_embedding at many positions, Llama (neox) and GLM (interleaved, 64 dims) forms;
    and the oracle's multi-position form against its single-position form that the library golden vectors pin."""
    import torch
    from oracle import dense_oracle as DO
    from shadowkv_b200 import synth
    from shadowkv_b200.llama import rotary_embedding_rows
    g = torch.Generator().manual_seed(4)
    L = 300
    x = (torch.randn(2, 3, L, 128, generator=g) * 2).to(torch.bfloat16)
    for glm in (False, True):
        cs = synth.cos_sin_table(5000, theta=10000.0 if glm else 500000.0, glm=glm)
        pos = torch.arange(4321, 4321 + L)
        rows = cs[pos]
        got = rotary_embedding_rows(x, rows, glm)
        want = DO.rotary_positions(x, rows, neox=not glm)
        assert torch.equal(got, want)
        for t in (0, 17, L - 1):
            one = DO.rotary_vllm(x[:, :, t], cs[pos[t]], neox=not glm)
            assert torch.equal(want[:, :, t], one)
        if glm:
            assert torch.equal(got[..., 64:], x[..., 64:])
