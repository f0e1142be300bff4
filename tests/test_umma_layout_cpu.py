"""Shared-memory layout arithmetic of the tcgen05 prefill attention (csrc/attention_prefill_tc.cu), emulated byte for byte:
what TMA's 128-byte swizzle writes, what the softmax warps store for P, and what a UMMA reads through the kernel's
K-major / MN-major descriptors (canonical layouts of cute/atom/mma_traits_sm100.hpp make_umma_desc:
K-major SW128 ((8,m),(T,2)):((8T,SBO),(1,T)); MN-major SW128 ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)), T = 8 fp16,
Swizzle<3,4,3> = byte-address bits [4,7) ^= bits [7,10)).  With the kernel's start addresses, k-step advances, LBO and SBO
the emulated S = Q K^T and O = P V must equal numpy's -- a slip in a stride or an advance shows up here, not as garbage
on the device.  (The descriptor semantics themselves are CUTLASS's documentation, not something a CPU can confirm.)"""
import numpy as np

SLAB = 128 * 128          # 128 rows x 128 B
D, BQ, BN = 128, 128, 128
SBO = 1024


def swz(addr):
    return addr ^ (((addr >> 7) & 7) << 4)


def tma_box_write(smem, base, tile):
    """tile: [rows, 64] fp16 -> box of `rows` rows x 128 B at `base` (1024-aligned) with the 128 B swizzle"""
    rows = tile.shape[0]
    raw = tile.astype(np.float16).view(np.uint8).reshape(rows, 128)
    for r in range(rows):
        for c in range(8):
            a = swz(base + r * 128 + c * 16)
            smem[a:a + 16] = raw[r, c * 16:(c + 1) * 16]


def rd16(smem, addr):
    return smem[addr:addr + 2].view(np.float16)[0]


def umma_read_kmajor(smem, start, rows):
    """operand [rows, 16] of one K = 16 step through a K-major SW128 descriptor whose start address is `start`"""
    out = np.empty((rows, 16), np.float16)
    for r in range(rows):
        for k in range(16):
            out[r, k] = rd16(smem, swz(start + (r % 8) * 128 + (r // 8) * SBO + k * 2))
    return out


def umma_read_mnmajor(smem, start, n, lbo):
    """operand [n (MN), 16 (K)] of one K = 16 step through an MN-major SW128 descriptor (LBO between 64-element slabs)"""
    out = np.empty((n, 16), np.float16)
    for mn in range(n):
        for k in range(16):
            off = (mn % 8) * 2 + ((mn % 64) // 8) * 16 + (mn // 64) * lbo + (k % 8) * 128 + (k // 8) * SBO
            out[mn, k] = rd16(smem, swz(start + off))
    return out


def test_prefill_tc_operand_layouts():
    rng = np.random.default_rng(0)
    Q = rng.standard_normal((BQ, D)).astype(np.float16)
    K = rng.standard_normal((BN, D)).astype(np.float16)
    V = rng.standard_normal((BN, D)).astype(np.float16)
    P = rng.random((BQ, BN)).astype(np.float16)
    smem = np.zeros(7 * 2 * SLAB, np.uint8)
    sQ, sK, sV, sP = 0, 2 * SLAB, 4 * SLAB, 6 * SLAB
    for h in range(2):  # the producer's boxes: 64 head dims per slab
        tma_box_write(smem, sQ + h * SLAB, Q[:, 64 * h:64 * h + 64])
        tma_box_write(smem, sK + h * SLAB, K[:, 64 * h:64 * h + 64])
        tma_box_write(smem, sV + h * SLAB, V[:, 64 * h:64 * h + 64])
    # the softmax warps' P store: thread r, 32-key group c, 16-byte piece q -> slab c >> 1, chunk ((c & 1) * 4 + q) ^ (r & 7)
    rawP = P.view(np.uint8).reshape(BQ, 2 * BN)
    for r in range(BQ):
        for c in range(BN // 32):
            for q in range(4):
                c16 = (c & 1) * 4 + q
                a = sP + (c >> 1) * SLAB + r * 128 + ((c16 ^ (r & 7)) << 4)
                smem[a:a + 16] = rawP[r, (c * 32 + q * 8) * 2:(c * 32 + q * 8 + 8) * 2]

    # S = Q K^T: k-step k reads slab k >> 2 at +32 B * (k & 3)   (desc_kmajor(...) + 2 * (k & 3))
    S = np.zeros((BQ, BN), np.float32)
    for k in range(D // 16):
        a = umma_read_kmajor(smem, sQ + (k >> 2) * SLAB + 32 * (k & 3), BQ).astype(np.float32)
        b = umma_read_kmajor(smem, sK + (k >> 2) * SLAB + 32 * (k & 3), BN).astype(np.float32)
        S += a @ b.T
    assert np.array_equal(S, Q.astype(np.float32) @ K.astype(np.float32).T) or np.allclose(S, Q.astype(np.float32) @ K.astype(np.float32).T, atol=1e-3)

    # O = P V: k-step k (16 keys) reads P slab k >> 2 at +32 B * (k & 3) and V at +16 rows (2048 B), LBO = one slab
    O = np.zeros((BQ, D), np.float32)
    for k in range(BN // 16):
        a = umma_read_kmajor(smem, sP + (k >> 2) * SLAB + 32 * (k & 3), BQ).astype(np.float32)   # [q, 16 keys]
        b = umma_read_mnmajor(smem, sV + k * 16 * 128, D, SLAB).astype(np.float32)                 # [d, 16 keys]
        O += a @ b.T
    assert np.allclose(O, P.astype(np.float32) @ V.astype(np.float32), atol=1e-3)
