"""Bit-exact parity of the chunk-cache integer/byte path against the oracle (and through it the
reference's own known-answer vectors).  All calls go through the C ABI via shadowkv_b200.shadowkv."""
import numpy as np
import pytest
import torch

from oracle import shadowkv_oracle as O
from tests import refkat

pytestmark = pytest.mark.gpu


def _reorder_gpu(dev, cached, cur):
    from shadowkv_b200 import shadowkv
    rows, S = cached.shape
    c = torch.from_numpy(cached.copy()).to(dev)
    q = torch.from_numpy(cur.copy()).to(dev)
    off = torch.zeros(rows * S, dtype=torch.int32, device=dev)
    cnt = torch.zeros(rows, dtype=torch.int32, device=dev)
    shadowkv.reorder_keys_and_compute_offsets(c, q, off, cnt, rows, 1, S)
    torch.cuda.synchronize()
    return c.cpu().numpy(), off.cpu().numpy().reshape(rows, S), cnt.cpu().numpy()


def test_reorder_reference_kat(cuda):
    keys, query = refkat.generate()
    gk, go, gc = refkat.golden_reorder(keys, query)
    nk, no, nc = _reorder_gpu(cuda, keys, query)
    np.testing.assert_array_equal(nc, gc)
    np.testing.assert_array_equal(nk, gk)
    np.testing.assert_array_equal(no, go)


@pytest.mark.parametrize("S", [128, 256, 512, 1024])
def test_reorder_random_vs_oracle(cuda, S):
    rng = np.random.default_rng(S)
    rows = 24
    cached = np.stack([rng.choice(70000, S, replace=False) for _ in range(rows)]).astype(np.int64)
    cur = cached.copy()
    for r in range(rows):
        nmiss = [0, S, 1, S - 1][r] if r < 4 else int(rng.integers(0, S + 1))
        fresh = np.setdiff1d(rng.choice(140000, 3 * S, replace=False), cached[r])[:nmiss]
        cur[r, rng.choice(S, nmiss, replace=False)] = fresh
        rng.shuffle(cur[r])
    want = O.reorder_keys_and_compute_offsets(cached, cur)
    got = _reorder_gpu(cuda, cached, cur)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)


def test_reorder_empty_table_and_duplicates(cuda):
    S = 256
    cached = np.full((3, S), -1, dtype=np.int64)
    cur = np.stack([np.random.default_rng(i).choice(9000, S, replace=False) for i in range(3)]).astype(np.int64)
    cached[2, :S // 2] = cur[2, :S // 2]
    cached[2, S // 2:] = cur[2, :S // 2]        # every resident id present twice: lowest slot wins
    want = O.reorder_keys_and_compute_offsets(cached, cur)
    got = _reorder_gpu(cuda, cached, cur)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    assert (got[2][:2] == 0).all() and got[2][2] == S // 2


def test_unsupported_map_size_raises(cuda):
    from shadowkv_b200 import shadowkv
    t = torch.zeros(100, dtype=torch.int64, device=cuda)
    o = torch.zeros(100, dtype=torch.int32, device=cuda)
    with pytest.raises(ValueError):
        shadowkv.reorder_keys_and_compute_offsets(t, t.clone(), o, o.clone(), 1, 1, 100)


def _case(rng, rows, S, nchunks, hit_fraction):
    cached = np.stack([rng.choice(nchunks, S, replace=False) for _ in range(rows)]).astype(np.int64)
    cur = cached.copy()
    for r in range(rows):
        nmiss = int(round(S * (1 - hit_fraction[r % len(hit_fraction)])))
        fresh = np.setdiff1d(rng.choice(nchunks, min(nchunks, 4 * S), replace=False), cached[r])[:nmiss]
        cur[r, rng.choice(S, len(fresh), replace=False)] = fresh
        rng.shuffle(cur[r])
    return cached, cur


@pytest.mark.parametrize("S", [128, 256, 512, 1024])
def test_cached_gather_d2d_and_h2d(cuda, S):
    """gather_copy_d2d_with_offsets + gather_copy_with_offsets, driven exactly as kv_cache.py:1081-1150 does."""
    from shadowkv_b200 import shadowkv
    rng = np.random.default_rng(100 + S)
    B, H, nchunks = 2, 3, 4096
    rows = B * H
    cached, cur = _case(rng, rows, S, nchunks, [0.6, 0.0, 1.0, 0.95, 0.3, 0.6])
    _, off, cnt = O.reorder_keys_and_compute_offsets(cached, cur)
    sparse_start, tail = 448, 96
    Lbuf = sparse_start + S * 8 + tail
    g = torch.Generator().manual_seed(S)
    host = torch.randn(rows, nchunks, 1024, generator=g).to(torch.bfloat16).pin_memory()
    kbuf = torch.randn(rows, Lbuf, 128, generator=g).to(torch.bfloat16)
    vbuf = torch.randn(rows, Lbuf, 128, generator=g).to(torch.bfloat16)
    want_k, want_v = kbuf.clone(), vbuf.clone()
    ks = want_k[:, sparse_start:sparse_start + S * 8].clone()
    O.gather_copy_d2d(ks, off, cnt)
    want_k[:, sparse_start:sparse_start + S * 8] = ks
    vs = want_v[:, sparse_start:sparse_start + S * 8].clone()
    O.gather_copy_with_offsets(host, vs, off, cnt)
    want_v[:, sparse_start:sparse_start + S * 8] = vs

    d_off = torch.from_numpy(off.reshape(-1)).to(cuda)
    d_cnt = torch.from_numpy(cnt).to(cuda)
    d_k, d_v = kbuf.to(cuda), vbuf.to(cuda)
    temp = torch.zeros(rows, S, 1024, dtype=torch.bfloat16, device=cuda)
    signals = torch.zeros(rows, dtype=torch.int32, device=cuda)
    shadowkv.gather_copy_d2d_with_offsets(d_k, d_off, d_cnt, B, H, S * 8 * 128, sparse_start * 128, Lbuf * 128, S)
    shadowkv.gather_copy_with_offsets(host, d_v, temp, d_off, d_cnt, signals, B, H, nchunks * 1024, S * 8 * 128,
                                      sparse_start * 128, Lbuf * 128, S)
    torch.cuda.synchronize()
    assert torch.equal(d_k.cpu(), want_k), "K compaction differs"
    assert torch.equal(d_v.cpu(), want_v), "V compaction + fetch differs"
    assert int(signals.abs().sum()) == 0          # reference invariant: signals back at 0 (copy.cuh:250-263)


def test_uncached_gather_copy(cuda):
    from shadowkv_b200 import shadowkv
    B, H, S, nchunks = 1, 4, 256, 2048
    g = torch.Generator().manual_seed(1)
    host = torch.randn(B * H, nchunks, 1024, generator=g).to(torch.bfloat16).pin_memory()
    ids = torch.stack([torch.randperm(nchunks, generator=g)[:S] for _ in range(B * H)])
    dst = torch.zeros(B * H, S, 1024, dtype=torch.bfloat16, device=cuda)
    shadowkv.gather_copy(host, dst, ids.to(cuda), B, H, nchunks * 1024, S * 1024, S)
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), O.gather_copy(host, ids.numpy()))


def test_round_trip_full_size(cuda):
    """BASELINE config-2 sizes (B=1, H=8, S=256, 15616 chunks): after any sequence of reorders + cached
    gathers, every sparse slot holds exactly the store chunk its table entry names (size-independent
    property; no oracle needed)."""
    from shadowkv_b200 import shadowkv
    B, H, S, nchunks = 1, 8, 256, 15616
    rows = B * H
    g = torch.Generator().manual_seed(2)
    host = torch.randint(-2 ** 15, 2 ** 15, (rows, nchunks, 1024), generator=g, dtype=torch.int16).view(torch.bfloat16).pin_memory()
    Lbuf, ss = 2592, 448
    table = torch.stack([torch.randperm(nchunks, generator=g)[:S] for _ in range(rows)]).to(cuda)
    vbuf = torch.zeros(rows, Lbuf, 128, dtype=torch.bfloat16, device=cuda)
    vbuf[:, ss:ss + S * 8] = torch.stack([host[r, table[r].cpu()] for r in range(rows)]).view(rows, S * 8, 128).to(cuda)
    off = torch.zeros(rows * S, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(rows, dtype=torch.int32, device=cuda)
    for step in range(6):
        keep = int(S * [0.6, 0.9, 0.0, 1.0, 0.5, 0.75][step])
        new = []
        for r in range(rows):
            old = table[r].cpu()
            kept = old[torch.randperm(S, generator=g)[:keep]]
            pool = torch.randperm(nchunks, generator=g)
            fresh = pool[~torch.isin(pool, old)][:S - keep]
            sel = torch.cat([kept, fresh])
            new.append(sel[torch.randperm(S, generator=g)])
        cur = torch.stack(new).to(cuda)
        shadowkv.reorder_keys_and_compute_offsets(table, cur, off, cnt, B, H, S)
        shadowkv.gather_copy_with_offsets(host, vbuf, None, off, cnt, None, B, H, nchunks * 1024, S * 1024, ss * 128,
                                          Lbuf * 128, S)
        torch.cuda.synchronize()
        assert (cnt.cpu() == keep).all()
        assert torch.equal(table.sort(dim=1).values, cur.sort(dim=1).values)
        got = vbuf[:, ss:ss + S * 8].reshape(rows, S, 1024).cpu()
        want = torch.stack([host[r, table[r].cpu()] for r in range(rows)])
        assert torch.equal(got.view(torch.int16), want.view(torch.int16)), f"step {step}"
