"""Re-creation of the input generator and golden model of the reference's own known-answer test
kernels/test_cached_gather_copy.cu (main :224-275, golden :70-156 and :158-216).  The key stream comes
from glibc rand() seeded with srand(12345) (:224), exactly as the reference test draws it, so the vectors
are the reference's own.  BLOCK_NUM = 8*24, 256 slots, 154 forced hits (:58-61)."""
import ctypes
import ctypes.util

import numpy as np

BLOCK_NUM = 8 * 24
MAP = 256
CPU_V_LENGTH = 128 * 1024
FAKE_HIT_CNT = 154


def generate(block_num=BLOCK_NUM):
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.srand(12345)
    rand = libc.rand
    keys = np.zeros((block_num, MAP), dtype=np.int64)
    query = np.zeros((block_num, MAP), dtype=np.int64)
    nchunks = CPU_V_LENGTH // 8
    for k in range(block_num):
        prev_keys, prev_reorders, prev_hits = set(), set(), set()
        for i in range(MAP):
            while True:
                key = rand() % nchunks
                if key not in prev_keys:
                    break
            prev_keys.add(key)
            keys[k, i] = key
        keys[k].sort()
        for i in range(MAP):
            if i < FAKE_HIT_CNT:
                while True:
                    hr = rand() % MAP
                    if hr not in prev_hits:
                        break
                prev_hits.add(hr)
                query[k, i] = keys[k, hr]
            else:
                while True:
                    r = rand() % nchunks
                    if r not in prev_reorders and r not in prev_keys:
                        break
                query[k, i] = r
            prev_reorders.add(int(query[k, i]))
    return keys, query


def golden_reorder(keys, query):
    """test_cached_gather_copy.cu:70-156 (sorts keys and offsets independently -- valid because the
    resident keys are pre-sorted there, :251)."""
    blocks = keys.shape[0]
    gk = np.zeros_like(keys)
    go = np.zeros(keys.shape, dtype=np.int32)
    cnts = np.zeros(blocks, dtype=np.int32)
    for k in range(blocks):
        lut = {}
        for i in range(MAP):
            lut.setdefault(int(keys[k, i]), i)
        hk, ho, misses = [], [], []
        for i in range(MAP):
            q = int(query[k, i])
            if q in lut:
                hk.append(q)
                ho.append(lut[q])
            else:
                misses.append(q)
        hk.sort(); ho.sort(); misses.sort()
        n = len(hk)
        cnts[k] = n
        gk[k, :n] = hk
        go[k, :n] = ho
        gk[k, n:] = misses
        go[k, n:] = misses
    return gk, go, cnts
