"""CPU: Philox4x32-10 restatement against the Random123 known-answer vectors + stream sanity."""
import numpy as np

from oracle import philox_ref as ph


def _one(c, k):
    r = ph.philox4x32_10(np.uint32(c[0]), np.uint32(c[1]), np.uint32(c[2]), np.uint32(c[3]), k[0], k[1])
    return [int(v) for v in r]


def test_known_answer_vectors():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _one([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _one([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _one([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normals_and_masks_statistics():
    z = ph.normals(12345, ph.STREAM_NORMAL, np.arange(2000), 3, 256).astype(np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert np.isfinite(z).all()
    m = ph.train_keep_masks(12345, 0, 2000, 250)
    assert m.shape == (3, 2000, 250) and set(np.unique(m)) == {0, 1}
    assert abs(m.mean() - 0.5) < 0.005 and not np.array_equal(m[0], m[1])
    zt = ph.train_normals(12345, 7, 500, 150).astype(np.float64)
    assert abs(zt.mean()) < 0.02 and abs(zt.std() - 1.0) < 0.02
    assert np.array_equal(ph.train_normals(12345, 107, 400, 150), ph.train_normals(12345, 7, 500, 150)[100:])
    assert np.array_equal(ph.train_keep_masks(12345, 107, 400, 150), ph.train_keep_masks(12345, 7, 500, 150)[:, 100:])
    m = ph.keep_masks128(12345, ph.STREAM_MASK, np.arange(2000), 3, 250)
    assert m.shape == (2000, 250) and set(np.unique(m)) == {0, 1}
    assert abs(m.mean() - 0.5) < 0.005
    # column 128*b + 32*w + j is bit j of word w of the Philox block b
    w = ph.philox4x32_10(np.uint32(1), np.uint32(3), np.uint32(7), np.uint32(ph.STREAM_MASK), 12345, 0)
    one = ph.keep_masks128(12345, ph.STREAM_MASK, np.array([7]), 3, 256)[0]
    assert all(int(one[128 + 32 * k + j]) == (int(w[k]) >> j) & 1 for k in range(4) for j in range(32))


def test_streams_are_keyed_by_global_row():
    a = ph.normals(7, ph.STREAM_NORMAL, np.arange(100, 200), 5, 40)
    b = ph.normals(7, ph.STREAM_NORMAL, np.arange(150, 160), 5, 40)
    assert np.array_equal(a[50:60], b)
    xT, z, keep = ph.sampler_noise(7, 100, 100, 40, 6)
    assert np.array_equal(z[5], a) and (z[1] == 0).all() and (z[0] == 0).all()
    assert not np.array_equal(ph.normals(7, ph.STREAM_NORMAL, np.arange(10), 4, 40),
                              ph.normals(8, ph.STREAM_NORMAL, np.arange(10), 4, 40))
