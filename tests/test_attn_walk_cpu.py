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


def test_decode_plan_wave_efficiency():
    """host logic of the decode attention launch (csrc/attention.cu plan_decode), through the C ABI without a device: the
    kernel's unit of residency is a warp (148 SMs x 12 slots = 1776); a plan is (KV splits, warps per CTA).  Properties:
    big grids are left alone; shapes that fit ONE wave at 0.75 .. 1.0 of the slots take that cut (one-warp CTAs); other shapes
    that would run in a poorly filled last wave get split until the wave efficiency is >= 0.9 where the sequence length
    allows; at least 8 units (128 tokens) per warp; partials fit the workspace."""
    import ctypes as C
    import b200_import
    b200_import.load()
    from ppl_llm_serving_b200 import capi
    lib = capi.load_library()

    def plan(batch, nq, nkv, kv):
        n, w = C.c_int32(), C.c_int32()
        assert lib.b2llm_attention_decode_plan(batch, nq, nkv, kv, C.byref(n), C.byref(w)) == 0
        return n.value, w.value

    def eff(batch, nq, nkv, kv):
        n, w = plan(batch, nq, nkv, kv)
        gq = nq // nkv
        G = 1 if gq == 1 else (4 if gq <= 4 else 8)
        warps = nkv * -(-gq // G) * batch * n * w
        waves = warps / 1776
        return waves / -(-warps // 1776), n, w

    assert plan(1024, 32, 32, 512) == (1, 1)                      # the 1-GPU benchmark shape: 18.45 waves, untouched
    assert plan(1024, 16, 16, 512) == (1, 1)                      # TP = 2: 9.2 waves
    e, n, w = eff(1024, 4, 4, 512)                                # 7B TP = 8: 2.3 waves alone (0.77)
    assert e >= 0.9 and n * w > 1
    e, n, w = eff(256, 8, 1, 8192)                                # 70B TP = 8: 256 CTAs -> ONE wave, 1536 of 1776 slots
    assert (n, w) == (6, 1) and abs(e - 1536 / 1776) < 1e-9       # measured faster than 13 splits = 0.94 of two waves (run 45)
    assert 8192 // 16 // (n * w) >= 8
    assert 256 * n <= max(2 * 256 + 444, 4096)                    # split partials fit attention_workspace_rows
    assert plan(16, 32, 32, 4096) == (3, 1)                       # 512 base CTAs -> 1536 in one wave
    assert plan(16, 32, 32, 300) == (1, 2)                        # 19 units allow 2 cuts: 1024 CTAs < 0.75 of a wave -> the old search (two warps)
    e, n, w = eff(512, 10, 10, 2640)                              # 13B TP = 4: 2.9 waves, no single-wave cut -> efficiency search
    assert e >= 0.9
    for batch, nq, nkv, kv in [(1, 32, 32, 17), (3, 8, 2, 300), (48, 4, 4, 400), (7, 64, 8, 4096), (2000, 40, 40, 33)]:
        n, w = plan(batch, nq, nkv, kv)
        assert n >= 1 and w in (1, 2, 4)
        units = -(-kv // 16)
        assert n * w == 1 or units // (n * w) >= 8
        assert batch * n <= max(2 * batch + 444, 4096)
