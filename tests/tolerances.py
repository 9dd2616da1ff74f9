"""Tolerances of the floating-point parity tests (north_star: "max-abs error 2e-2 and cosine similarity
above 0.999" for bf16 results; gathers and chunk-index selection bit-exact)."""
import torch

MAX_ABS = 2e-2          # attention output / logits-level quantities (O(1) magnitudes)
MIN_COS = 0.999
# a bf16 result computed with a different fp32 summation order can land on the neighbouring bf16 value:
# relative 2^-8 of the magnitude.  Rebuilt K entries are O(1)..O(4) -> up to 1 bf16 ulp = 2^-7 * 4.
BF16_REL = 2.0 ** -7


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
