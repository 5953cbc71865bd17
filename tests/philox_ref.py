"""numpy restatement of the engine's sampler spec (test helper): Philox4x32-10 keyed by the run seed with counter
(sim lo, sim hi, row t, pair p) -> two 64-bit words -> Box-Muller pair for columns (2p, 2p+1)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def standard_normals(sim_ids, n_obs, n_assets, seed):
    """Z[len(sim_ids)][T][N] exactly as philox_normal_kernel defines it."""
    sim_ids = np.asarray(sim_ids, dtype=np.uint64)
    pairs = (n_assets + 1) // 2
    s, t, p = np.meshgrid(sim_ids, np.arange(n_obs, dtype=np.uint64), np.arange(pairs, dtype=np.uint64), indexing="ij")
    r0, r1, r2, r3 = philox4x32_10((s & MASK).astype(np.uint32), (s >> np.uint64(32)).astype(np.uint32),
                                   t.astype(np.uint32), p.astype(np.uint32), np.uint64(seed) & MASK,
                                   np.uint64(seed) >> np.uint64(32))
    a = (r1.astype(np.uint64) << np.uint64(32)) | r0.astype(np.uint64)
    b = (r3.astype(np.uint64) << np.uint64(32)) | r2.astype(np.uint64)
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
    u2 = (b >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(s.shape[:2] + (2 * pairs,))
    z[..., 0::2] = rad * np.cos(2.0 * np.pi * u2)
    z[..., 1::2] = rad * np.sin(2.0 * np.pi * u2)
    return z[..., :n_assets]
