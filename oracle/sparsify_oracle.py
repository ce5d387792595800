"""TEST INFRASTRUCTURE — CPU restatement of the equal-sparsity thresholding step the reference performs on the host
(main.py:177-185, 259-262; hyperparameter_search.py:162-166) and a NumPy stand-in for kernel K4's histogram pass.
Only tests / smoke / the bench's cpu_baseline leg may import this; the product path never does.
"""
import numpy as np


def equal_sparsity_reference(scores, sparsity, lower=False):
    """Exactly the reference's lines: threshold = np.quantile(S.flatten(), q); (S >= threshold).astype(int)."""
    if lower:
        thr = np.quantile(scores.flatten(), 1 - sparsity)   # main.py:260
        return (scores <= thr).astype(int), thr             # main.py:262
    thr = np.quantile(scores.flatten(), sparsity)            # main.py:177
    return (scores >= thr).astype(int), thr                  # main.py:178


def score_keys(a):
    """Order-preserving uint32 keys of float32 values (sparsify.cu: score_key)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel()
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def numpy_histogram_fn(a):
    """hist_fn(prefix, prefix_bits, shift, bits) over the values of `a`, as sdrm_key_histogram computes it."""
    keys = score_keys(a)

    def hist_fn(prefix, prefix_bits, shift, bits):
        k = keys
        if prefix_bits:
            k = k[(k >> np.uint32(32 - prefix_bits)) == np.uint32(prefix)]
        digit = (k >> np.uint32(shift)) & np.uint32((1 << bits) - 1)
        h = np.zeros(2048, dtype=np.int64)
        h[: 1 << bits] = np.bincount(digit.astype(np.int64), minlength=1 << bits)
        return h
    return hist_fn


def pack_bits(dense01):
    """[rows, cols] 0/1 -> uint32 words, bit j of word w = column 32 w + j (the device layout)."""
    rows, cols = dense01.shape
    wpr = (cols + 31) // 32
    padded = np.zeros((rows, wpr * 32), dtype=np.uint8)
    padded[:, :cols] = dense01.astype(np.uint8)
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(rows, wpr)
