"""Pins the oracle's integer path against the reference's own known-answer vectors (CPU only)."""
import numpy as np
import torch

from oracle import shadowkv_oracle as O
from tests import refkat


def test_reorder_matches_reference_kat():
    keys, query = refkat.generate()
    gk, go, gc = refkat.golden_reorder(keys, query)
    nk, no, nc = O.reorder_keys_and_compute_offsets(keys, query)
    assert (gc == refkat.FAKE_HIT_CNT).all()
    np.testing.assert_array_equal(nc, gc)
    np.testing.assert_array_equal(nk, gk)
    np.testing.assert_array_equal(no, go)


def test_cached_copy_matches_reference_golden():
    # golden model of test_cached_gather_copy.cu:158-216 on a scaled-down store: slot i < cnt takes the OLD
    # buffer chunk offsets[i], slot i >= cnt takes host chunk offsets[i].
    blocks = 6
    keys, query = refkat.generate(blocks)
    _, off, cnt = O.reorder_keys_and_compute_offsets(keys, query)
    g = torch.Generator().manual_seed(7)
    # host store restricted to the chunks actually referenced (remap ids to a dense range)
    host = torch.randn(blocks, 128 * 1024 // 8, 16, generator=g).to(torch.bfloat16)      # 16 elems/chunk stand-in
    buf = torch.randn(blocks, refkat.MAP, 16, generator=g).to(torch.bfloat16)
    want = torch.empty_like(buf)
    for k in range(blocks):
        for i in range(refkat.MAP):
            want[k, i] = buf[k, off[k, i]] if i < cnt[k] else host[k, off[k, i]]
    got = buf.clone().view(blocks, refkat.MAP * 1, 16)
    # oracle works on [rows, S*C, D] with C*D elements per chunk: use C=1, D=16
    view = got.view(blocks, refkat.MAP, 16)
    O.gather_copy_with_offsets(host, view.view(blocks, refkat.MAP, 16), off, cnt)
    assert torch.equal(view, want)


def test_reorder_properties_random():
    rng = np.random.default_rng(0)
    for S in (128, 256, 512, 1024):
        rows = 5
        cached = np.stack([rng.choice(20000, S, replace=False) for _ in range(rows)]).astype(np.int64)
        cur = cached.copy()
        for r in range(rows):
            nmiss = int(rng.integers(0, S + 1))
            fresh = np.setdiff1d(rng.choice(40000, 3 * S, replace=False), cached[r])[:nmiss]
            pos = rng.choice(S, nmiss, replace=False)
            cur[r, pos] = fresh
            rng.shuffle(cur[r])
        nk, off, cnt = O.reorder_keys_and_compute_offsets(cached, cur)
        for r in range(rows):
            c = int(cnt[r])
            assert set(nk[r]) == set(cur[r])                       # same set of chunks
            assert (np.diff(off[r, :c]) > 0).all()                 # hits by ascending old slot
            assert (off[r, :c] >= np.arange(c)).all()              # forward moves only
            assert (cached[r][off[r, :c]] == nk[r, :c]).all()      # offset points at the resident copy
            assert (np.diff(nk[r, c:]) > 0).all()                  # misses by ascending id
            assert (off[r, c:] == nk[r, c:]).all()
            assert not set(nk[r, c:]) & set(cached[r])


def test_empty_table_is_all_misses():
    cached = np.full((2, 128), -1, dtype=np.int64)
    cur = np.stack([np.random.default_rng(i).choice(5000, 128, replace=False) for i in range(2)]).astype(np.int64)
    nk, off, cnt = O.reorder_keys_and_compute_offsets(cached, cur)
    assert (cnt == 0).all()
    np.testing.assert_array_equal(nk, np.sort(cur, axis=1))
    np.testing.assert_array_equal(off, np.sort(cur, axis=1).astype(np.int32))
