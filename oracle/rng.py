"""oracle/rng.py -- TEST INFRASTRUCTURE.  numpy restatement of the production sample stream (Philox-4x32-10, Salmon et al.
2011, keyed by (seed, lane, column block)) so that the oracle can be fed exactly the uniforms the CUDA kernels draw."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniforms(seed, lane_offset, n, dims):
    """(n, dims) float32 uniforms in [0,1): column 4b+k is word k of philox(counter=(lane_lo, lane_hi, b, 0), key=seed)."""
    lane = np.arange(n, dtype=np.uint64) + np.uint64(lane_offset)
    out = np.empty((n, dims), np.float32)
    for b in range((dims + 3) // 4):
        w = philox4x32_10(lane & MASK, lane >> np.uint64(32), np.full(n, b, np.uint64), np.zeros(n, np.uint64), seed & 0xFFFFFFFF, seed >> 32)
        for k in range(4):
            if 4 * b + k < dims:
                out[:, 4 * b + k] = (w[k] >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return out
