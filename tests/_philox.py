"""numpy re-statement of the engine's counter-based normal generator (abr_kernels.cuh:
philox4x32 + philox_normal): Philox4x32-10, key = (seed_lo, seed_hi),
counter = (global_sample, index >> 2, problem, 0x5eed), Box-Muller on the pair (index >> 1) & 1,
cos branch for even index, sin for odd."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def normals(seed: int, sample, problem, index):
    """float32 normals for (global sample id, problem id, flat index t*nu+u); arrays broadcast."""
    sample, problem, index = np.broadcast_arrays(np.asarray(sample, dtype=np.uint32), np.asarray(problem, dtype=np.uint32),
                                                  np.asarray(index, dtype=np.uint32))
    w = philox4x32(sample, index >> np.uint32(2), problem, np.uint32(0x5EED), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    pr = ((index >> np.uint32(1)) & np.uint32(1)).astype(bool)
    a = np.where(pr, w[2], w[0]).astype(np.float32)
    b = np.where(pr, w[3], w[1]).astype(np.float32)
    scale = np.float32(2.3283064365386963e-10)
    u1 = (a + np.float32(0.5)) * scale
    u2 = (b + np.float32(0.5)) * scale
    r = np.sqrt(np.float32(-2.0) * np.log(u1, dtype=np.float32), dtype=np.float32)
    ang = np.float32(2.0) * u2
    odd = (index & np.uint32(1)).astype(bool)
    return np.where(odd, r * np.sin(np.pi * ang.astype(np.float64)).astype(np.float32),
                    r * np.cos(np.pi * ang.astype(np.float64)).astype(np.float32)).astype(np.float32)
