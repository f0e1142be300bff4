"""The "slim" decode-attention loader (csrc/attention.cu, LOADER == 2) walks the page table incrementally -- (page entry,
offset) advanced by WARPS units per issue -- instead of dividing the unit's position by page_size in the issuing lane
for every 16-token unit.  This is the integer logic of `walk_slot0` / `walk_advance` restated next to `unit_slot0`
(the baseline loader), checked for every issue a CTA makes: all WARPS, split ranges, page sizes that are multiples of
the 16-token unit (the TMA path's precondition), and the contiguous-index cache mode.  The GPU parity tests
(tests/test_ops_gpu.py, run with the loader in use) cover page_size 16 and cache_mode 0; this covers the rest."""
import numpy as np
import pytest

UNIT, NSTAGE = 16, 3


def unit_slot0(cache_mode, page_size, cache_indices, b, max_pages, u):
    pos = u * UNIT
    if cache_mode == 0:
        return cache_indices[b] + pos
    return cache_indices[b * max_pages + pos // page_size] + pos % page_size


def issued_slots_walk(cache_mode, page_size, cache_indices, b, max_pages, u0, u1, warps, warp):
    """slots of the units warp `warp` issues, in issue order, by the incremental walk"""
    pos0 = (u0 + warp) * UNIT
    if cache_mode == 0:
        walk_page, walk_off, adv_page, adv_off = 0, pos0, 0, warps * UNIT
    else:
        walk_page, walk_off = pos0 // page_size, pos0 % page_size
        adv_page, adv_off = (warps * UNIT) // page_size, (warps * UNIT) % page_size
    out = []
    u_issue = u0 + warp
    n_iter = len(range(u0 + warp, u1, warps))
    for _ in range(NSTAGE - 1 + n_iter):           # prologue issues + one issue per consumed unit
        if u_issue < u1:
            if cache_mode == 0:
                out.append((u_issue, cache_indices[b] + walk_off))
            else:
                out.append((u_issue, cache_indices[b * max_pages + walk_page] + walk_off))
        walk_page += adv_page
        walk_off += adv_off
        if cache_mode != 0 and walk_off >= page_size:
            walk_off -= page_size
            walk_page += 1
        u_issue += warps
    return out


@pytest.mark.parametrize("page_size", [16, 32, 48, 64, 128, 256])
@pytest.mark.parametrize("warps", [1, 2, 4])
def test_incremental_page_walk_equals_division(page_size, warps):
    rng = np.random.default_rng(page_size * 8 + warps)
    for kv_len in [1, 15, 16, 17, 255, 256, 257, 512, 1000, 2048]:
        units_total = (kv_len + UNIT - 1) // UNIT
        max_pages = (kv_len + page_size - 1) // page_size + 1
        B = 3
        table = (rng.permutation(B * max_pages) * page_size).astype(np.int64)
        for nsplit in [1, 2, 3, 7]:
            ups = (units_total + nsplit - 1) // nsplit
            for split in range(nsplit):
                u0, u1 = split * ups, min(units_total, split * ups + ups)
                for warp in range(warps):
                    for b in (0, B - 1):
                        got = issued_slots_walk(1, page_size, table, b, max_pages, u0, u1, warps, warp)
                        want = [(u, unit_slot0(1, page_size, table, b, max_pages, u)) for u in range(u0 + warp, u1, warps)]
                        assert got == want, (kv_len, nsplit, split, warp)


@pytest.mark.parametrize("warps", [1, 2, 4])
def test_incremental_walk_contiguous_index_mode(warps):
    idx = np.array([1000, 5, 77777], dtype=np.int64)
    for kv_len in [1, 16, 333, 2048]:
        units_total = (kv_len + UNIT - 1) // UNIT
        for warp in range(warps):
            got = issued_slots_walk(0, 16, idx, 2, 0, 0, units_total, warps, warp)
            want = [(u, unit_slot0(0, 16, idx, 2, 0, u)) for u in range(warp, units_total, warps)]
            assert got == want
