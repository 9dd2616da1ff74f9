"""Deterministic inputs for the golden fixtures: bf16 bit patterns from integer hashing only (no libm, no
RNG state), so the generator on the GPU box and the tests anywhere rebuild the very same tensors."""
import numpy as np
import torch


def bf16_tensor(shape, seed, exp_base=0x7C, exp_span=5):
    """Values with sign/mantissa from a multiplicative hash; magnitudes 2^(exp_base-127) .. 2^(exp_base-127+exp_span)."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    h = (idx * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    sign = (h >> np.uint64(31)) & np.uint64(1)
    exp = np.uint64(exp_base) + ((h >> np.uint64(24)) & np.uint64(0x7F)) % np.uint64(exp_span)
    mant = (h >> np.uint64(8)) & np.uint64(0x7F)
    bits = ((sign << np.uint64(15)) | (exp << np.uint64(7)) | mant).astype(np.uint16)
    return torch.from_numpy(bits.view(np.int16).reshape(shape).copy()).view(torch.bfloat16)


def distinct_ids(rows, count, universe, seed):
    """`count` distinct ids in [0, universe) per row, in a hash-shuffled order."""
    out = np.zeros((rows, count), dtype=np.int64)
    for r in range(rows):
        mult = 2 * ((seed * 7919 + r * 104729) % (universe // 2)) + 1
        while np.gcd(mult, universe) != 1:
            mult += 2
        out[r] = (np.arange(count, dtype=np.int64) * mult + (seed * 31 + r * 17)) % universe
    return out


def bits(t):
    return t.contiguous().view(torch.int16).cpu().numpy()
