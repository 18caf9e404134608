"""Philox4x32-10 restated in numpy (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

The product draws its throughput-run random numbers inside the CUDA kernels
(cv-nerf_b200/csrc/rng.cuh); this file restates the published generator (Salmon, Moraes, Dror,
Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 `philox4x32_R(10, ...)`) and the
mapping from (seed, stream, ray, value index) to a uniform / normal value, so that tests can (a) pin
the CUDA generator to the Random123 known-answer vectors and (b) hand the very numbers a kernel drew
to the CPU oracle.  The reference itself draws torch.rand / torch.randn on the host (main.py:188,233,
utils.py:23); which uniform numbers are consumed is not part of its contract.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)

# Random123 kat_vectors, philox4x32 10 rounds: (counter, key, expected)
KNOWN_ANSWERS = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(counter, key):
    """counter: 4 uint32 arrays (broadcastable), key: 2 uint32 values/arrays -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint32) for x in counter]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c


def draw_words(seed, stream, ray0, n, cols):
    """uint32 [n, 4*ceil(cols/4)]: word (i & 3) of group (i >> 2) of ray (ray0 + r), as csrc/rng.cuh."""
    groups = (cols + 3) // 4
    ray = (np.arange(n, dtype=np.uint64) + np.uint64(ray0))[:, None]
    grp = np.arange(groups, dtype=np.uint32)[None, :]
    lo, hi = (ray & MASK).astype(np.uint32), (ray >> np.uint64(32)).astype(np.uint32)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    w = philox4x32_10((lo, hi, grp, np.uint32(stream)), (seed & 0xFFFFFFFF, seed >> 32))
    return np.stack(w, -1).reshape(n, groups * 4)


def uniforms(seed, stream, ray0, n, cols):
    """float32 [n, cols] in [0, 1): (word >> 8) * 2^-24."""
    w = draw_words(seed, stream, ray0, n, cols)[:, :cols]
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))


def normals(seed, stream, ray0, n, cols):
    """float32 [n, cols]: Box-Muller over the word pairs (0,1) and (2,3) of every group."""
    w = draw_words(seed, stream, ray0, n, cols)
    a, b = w[:, 0::2], w[:, 1::2]
    u1 = ((a >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    r = np.sqrt(-2.0 * np.log(u1.astype(np.float32).astype(np.float64)))
    t = np.float32(6.283185307179586) * ((b >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))
    out = np.empty(w.shape, dtype=np.float32)
    out[:, 0::2] = (r * np.cos(t.astype(np.float64))).astype(np.float32)
    out[:, 1::2] = (r * np.sin(t.astype(np.float64))).astype(np.float32)
    return out[:, :cols]
