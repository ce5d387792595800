"""TEST INFRASTRUCTURE — numpy restatement of the in-kernel Philox4x32-10 noise streams
(sdrm_b200/csrc/philox.cuh).  Philox4x32-10 is the published algorithm of Salmon, Moraes, Dror, Shaw,
"Parallel random numbers: as easy as 1, 2, 3" (SC'11); the known-answer vectors of the Random123
distribution (kat_vectors) pin it in tests/test_philox.py.

The reference (SDRM) draws its noise from torch's global generator (train_SDRM.py:38,51,56,100), which
cannot be reproduced bit-for-bit on another device; parity of the arithmetic is therefore proven with
injected noise, and this module lets the tests inject EXACTLY the noise the kernel would generate.
"""
import numpy as np

STREAM_NORMAL, STREAM_MASK, STREAM_TRAIN_NOISE, STREAM_TRAIN_MASK = 0, 1, 2, 3
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _box_muller(a, b):
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)
    u2 = (b >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float32(6.283185307179586) * u2).astype(np.float32)
    return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)


def normals(seed, stream, rows, step, n_cols):
    """N(0,1) block [len(rows), n_cols] for global row ids `rows` at `step` (float32)."""
    rows = np.asarray(rows, dtype=np.uint64)
    nq = (n_cols + 3) // 4
    cq = np.arange(nq, dtype=np.uint32)[None, :]
    r_lo = (rows & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    c3 = (np.uint32(stream) | ((rows >> np.uint64(32)).astype(np.uint32) << np.uint32(8)))[:, None]
    x, y, z, w = philox4x32_10(cq, np.uint32(step), r_lo, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = _box_muller(x, y)
    z2, z3 = _box_muller(z, w)
    out = np.stack([z0, z1, z2, z3], axis=-1).reshape(len(rows), nq * 4)
    return out[:, :n_cols]


def keep_masks128(seed, stream, rows, step, n_cols):
    """Sampler dropout keep mask block [len(rows), n_cols] (uint8 0/1): one Philox call per 128 columns,
    column 128*b + 32*w + j is bit j of output word w (philox_mask128 in philox.cuh)."""
    rows = np.asarray(rows, dtype=np.uint64)
    nb = (n_cols + 127) // 128
    cb = np.arange(nb, dtype=np.uint32)[None, :]
    r_lo = (rows & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    c3 = (np.uint32(stream) | ((rows >> np.uint64(32)).astype(np.uint32) << np.uint32(8)))[:, None]
    words = philox4x32_10(cb, np.uint32(step), r_lo, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = np.stack(words, axis=-1)                                   # [rows, nb, 4]
    bits = (w[..., None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)   # [rows, nb, 4, 32]
    return bits.reshape(len(rows), nb * 128)[:, :n_cols].astype(np.uint8)


def _flat_counter(idx):
    idx = np.asarray(idx, dtype=np.uint64)
    return (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32)


def train_normals(seed, row_offset, B, L):
    """N(0,1) block [B, L] of the training step (noise_inputs_kernel): keyed by the GLOBAL flat element index
    i = (row_offset + b) * L + f; quad i // 4 is one Philox call, element i takes Box-Muller output i % 4."""
    i = np.uint64(row_offset) * np.uint64(L) + np.arange(B * L, dtype=np.uint64)
    q = np.unique(i >> np.uint64(2))
    c0, c1 = _flat_counter(q)
    x, y, z, w = philox4x32_10(c0, c1, np.uint32(0), np.uint32(STREAM_TRAIN_NOISE), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = _box_muller(x, y)
    z2, z3 = _box_muller(z, w)
    table = np.stack([z0, z1, z2, z3], axis=-1).reshape(-1)          # indexed by 4 * (q - q[0]) + lane
    pos = (i - (q[0] << np.uint64(2))).astype(np.int64)
    return table[pos].reshape(B, L)


def train_keep_masks(seed, row_offset, B, L):
    """The three dropout keep masks of the training step [3, B, L] (uint8): element i = 512 blk + 4 (32 j + l) + e takes
    bit (4 j + e) of output word k of Philox(c0|c1 = 32 blk + l, c2 = 0, c3 = STREAM_TRAIN_MASK)."""
    i = np.uint64(row_offset) * np.uint64(L) + np.arange(B * L, dtype=np.uint64)
    blk = i >> np.uint64(9)
    within = (i & np.uint64(511)).astype(np.int64)
    quad, e = within >> 2, within & 3
    j, lane = quad >> 5, quad & 31
    c0, c1 = _flat_counter(blk * np.uint64(32) + lane.astype(np.uint64))
    words = philox4x32_10(c0, c1, np.uint32(0), np.uint32(STREAM_TRAIN_MASK), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bit = (4 * j + e).astype(np.uint32)
    return np.stack([((words[k] >> bit) & np.uint32(1)).astype(np.uint8).reshape(B, L) for k in range(3)])


def sampler_noise(seed, row_offset, n, L, T):
    """The exact (x_T, z[T+1], keep[T+1]) tensors sdrm_sample generates in-kernel for rows row_offset..+n."""
    rows = np.arange(row_offset, row_offset + n, dtype=np.uint64)
    xT = normals(seed, STREAM_NORMAL, rows, 0, L)
    z = np.zeros((T + 1, n, L), dtype=np.float32)
    keep = np.zeros((T + 1, n, L), dtype=np.uint8)
    for i in range(1, T + 1):
        if i >= 2:
            z[i] = normals(seed, STREAM_NORMAL, rows, i, L)
        keep[i] = keep_masks128(seed, STREAM_MASK, rows, i, L)
    return xT, z, keep
