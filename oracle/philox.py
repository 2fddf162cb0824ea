"""numpy restatement of the Philox4x32-10 generator in csrc/common.cuh (test oracle only)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(seed, ctr_lo, ctr_hi):
    """seed: python int (64 bit); ctr_lo: uint64 array; ctr_hi: python int.  Returns uint32 [n,4]."""
    ctr_lo = np.asarray(ctr_lo, dtype=np.uint64)
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    c0 = ctr_lo & MASK
    c1 = ctr_lo >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(int(ctr_hi) & 0xFFFFFFFF))
    c3 = np.full_like(c0, np.uint64((int(ctr_hi) >> 32) & 0xFFFFFFFF))
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & MASK
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u53(a, b):
    a = a.astype(np.uint64)
    b = b.astype(np.uint64)
    return (((a >> np.uint64(5)) << np.uint64(26)) | (b >> np.uint64(6))).astype(np.float64) * (1.0 / 9007199254740992.0)


def u24(a):
    return (a >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


STREAM_HER, STREAM_EXPLORE, STREAM_RESET = 1, 2, 3
