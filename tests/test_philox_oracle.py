"""Pin oracle/philox_np.py against the published Philox4x32-10 known-answer vectors (Random123
kat_vectors: philox4x32 10 rounds) and sanity-check the float32 transforms.  CPU only."""
import numpy as np

from oracle import philox_np as ph


def _hex(t):
    return [f"{int(v):08x}" for v in t]


def test_philox4x32_10_known_answers():
    assert _hex(ph.philox4x32_10(0, 0, 0, 0, 0, 0)) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    f = 0xFFFFFFFF
    assert _hex(ph.philox4x32_10(f, f, f, f, f, f)) == ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert _hex(ph.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)) == \
        ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]


def test_u01_open_interval_and_exact():
    edge = np.array([0, 1, 511, 512, 2**32 - 1], dtype=np.uint64).astype(np.uint32)
    u = ph.u01(edge)
    assert u.dtype == np.float32 and (u > 0).all() and (u < 1).all()
    assert u[0] == np.float32(2.0 ** -24) and u[-1] == np.float32(1 - 2.0 ** -24)


def test_transforms_accuracy_and_moments():
    x = np.random.default_rng(0).integers(0, 2**32, size=1_000_000, dtype=np.uint64).astype(np.uint32)
    u = ph.u01(x)
    ref = np.log(u.astype(np.float64))
    assert np.max(np.abs(ph._log_f32(u) - ref) / np.maximum(np.abs(ref), 1e-6)) < 1e-6
    s, c = ph._sincos_2pi_f32(u)
    assert np.max(np.abs(s - np.sin(2 * np.pi * u.astype(np.float64)))) < 5e-7
    assert np.max(np.abs(c - np.cos(2 * np.pi * u.astype(np.float64)))) < 5e-7
    z = np.concatenate(ph.normals4(1, np.arange(1_000_000), 3, ph.PURPOSE_OU)).astype(np.float64)
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 3e-3
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3) < 0.02


def test_streams_are_distinct_and_reproducible():
    a = ph.draw(1, np.arange(8), 0, ph.PURPOSE_OU)
    b = ph.draw(1, np.arange(8), 0, ph.PURPOSE_OU)
    c = ph.draw(1, np.arange(8), 1, ph.PURPOSE_OU)
    d = ph.draw(2, np.arange(8), 0, ph.PURPOSE_OU)
    e = ph.draw(1, np.arange(8), 0, ph.PURPOSE_LEADER_EXOG)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for other in (c, d, e):
        assert not any(np.array_equal(x, y) for x, y in zip(a, other))
    big = ph.draw(1, np.array([5, 5 + 2**32], dtype=np.uint64), 0, 0)   # id high word participates
    assert big[0][0] != big[0][1]


def test_replay_indices_range_and_layout():
    idx = ph.replay_indices(1, [0, 1, 2], 7, 64, 100)
    assert idx.shape == (3, 64) and idx.dtype == np.int64 and idx.min() >= 0 and idx.max() < 100
    assert np.array_equal(idx[1], ph.replay_indices(1, [1], 7, 64, 100)[0])     # ring streams independent of batch of rings
    assert (ph.replay_indices(1, [0], 7, 64, 1) == 0).all()
    h = np.bincount(ph.replay_indices(3, np.arange(4096), 0, 64, 10).ravel(), minlength=10) / (4096 * 64)
    assert np.abs(h - 0.1).max() < 0.005
