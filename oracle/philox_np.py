"""Host restatement of the counter-based RNG used by the CUDA path.  TEST INFRASTRUCTURE ONLY.

The reference draws everything from NumPy's global MT19937 stream (src/util.py:55-70,
src/replaybuffer.py:54), which is inherently serial.  SURVEY.md §7 therefore defines RNG
parity as: the CUDA generator and this independent host restatement agree BIT-EXACTLY, and
parity with the reference proper is obtained by injecting the reference's draws.

Generator: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123 v1.x).  Pinned here against the published known-answer vectors
(tests/test_philox_oracle.py).

Stream addressing (must match avddpg_b200/csrc/avd_rng.cuh):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (id_lo, id_hi, tick, purpose)
with ``purpose`` one of the PURPOSE_* constants below, ``id`` the global platoon / vehicle /
ring id (independent of how platoons are sharded over GPUs) and ``tick`` the step, episode or
update counter.

Transforms (all arithmetic is single IEEE-754 binary32 operations -- no fused multiply-add --
so NumPy float32 and CUDA ``__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn`` give identical bits):
    u01(x)      = ((x >> 9) + 0.5) * 2**-23                    in (0, 1), exact in fp32
    normal pair = Box-Muller with the polynomial log / sincos below
    index(x, n) = (x * n) >> 32                                 (uint64 multiply-shift)
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

PURPOSE_RESET_VEHICLE = 0   # id = global vehicle id (p*M+m), tick = episode index
PURPOSE_RESET_PLATOON = 1   # id = global platoon id,          tick = episode index
PURPOSE_OU = 2              # id = global vehicle id,          tick = step counter
PURPOSE_LEADER_EXOG = 3     # id = global platoon id,          tick = step counter
PURPOSE_REPLAY = 4          # id = ring id * (batch/4) + j/4,  tick = update counter
PURPOSE_INIT = 5            # id = parameter index / 4,        tick = tensor tag

_U32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable unsigned integers < 2**32.
    Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & _U32 for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _U32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _U32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def seed_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def draw(seed, ids, tick, purpose):
    """Four uint32 words for every id (ids may exceed 2**32: split into lo/hi)."""
    ids = np.asarray(ids, dtype=np.uint64)
    k0, k1 = seed_key(seed)
    return philox4x32_10(ids & _U32, ids >> np.uint64(32), tick, purpose, k0, k1)


# ----------------------------------------------------------------------------- transforms
F32 = np.float32


def u01(x):
    """uint32 -> float32 strictly inside (0,1); exact (23-bit grid + half step = 24 bits)."""
    x = np.asarray(x, dtype=np.uint32)
    return ((x >> np.uint32(9)).astype(F32) + F32(0.5)) * F32(2.0 ** -23)


def _log_f32(u):
    """Natural log of float32 u in (0,1] using only single fp32 ops.
    u = m * 2**e with m in [sqrt(1/2), sqrt(2)); ln m = 2*atanh(s), s=(m-1)/(m+1)."""
    u = np.asarray(u, dtype=F32)
    bits = u.view(np.uint32).astype(np.int64)
    e = ((bits >> 23) & 0xFF) - 127
    mbits = (bits & 0x007FFFFF) | 0x3F800000          # mantissa with exponent 0 -> [1,2)
    big = mbits >= 0x3FB504F3                           # m >= sqrt(2) (as float bits)
    mbits = np.where(big, mbits - 0x00800000, mbits)    # halve -> [sqrt(.5), sqrt(2))
    e = np.where(big, e + 1, e)
    m = mbits.astype(np.uint32).view(F32)
    s = (m - F32(1.0)) / (m + F32(1.0))
    z = s * s
    p = F32(2.0 / 9.0)
    p = p * z + F32(2.0 / 7.0)
    p = p * z + F32(2.0 / 5.0)
    p = p * z + F32(2.0 / 3.0)
    p = p * z + F32(2.0)
    lnm = s * p
    return e.astype(F32) * F32(0.6931471805599453) + lnm


def _sincos_2pi_f32(v):
    """(sin, cos) of 2*pi*v for float32 v in (0,1), single fp32 ops only."""
    v = np.asarray(v, dtype=F32)
    t = v * F32(4.0)                                   # exact
    k = np.floor(t + F32(0.5))                         # nearest quadrant 0..4
    f = t - k                                          # exact, in [-0.5, 0.5]
    x = f * F32(1.5707963267948966)
    z = x * x
    # sin(x) ~ x*(1 + z*(-1/6 + z*(1/120 + z*(-1/5040 + z/362880))))
    ps = F32(1.0 / 362880.0)
    ps = ps * z + F32(-1.0 / 5040.0)
    ps = ps * z + F32(1.0 / 120.0)
    ps = ps * z + F32(-1.0 / 6.0)
    ps = ps * z + F32(1.0)
    sx = x * ps
    # cos(x) ~ 1 + z*(-1/2 + z*(1/24 + z*(-1/720 + z*(1/40320 - z/3628800))))
    pc = F32(-1.0 / 3628800.0)
    pc = pc * z + F32(1.0 / 40320.0)
    pc = pc * z + F32(-1.0 / 720.0)
    pc = pc * z + F32(1.0 / 24.0)
    pc = pc * z + F32(-0.5)
    cx = pc * z + F32(1.0)
    q = k.astype(np.int64) & 3
    s = np.where(q == 0, sx, np.where(q == 1, cx, np.where(q == 2, -sx, -cx)))
    c = np.where(q == 0, cx, np.where(q == 1, -sx, np.where(q == 2, -cx, sx)))
    return s.astype(F32), c.astype(F32)


def normal_pair(xa, xb):
    """Two independent N(0,1) float32 from two uint32 words (Box-Muller)."""
    u1 = u01(xa)
    u2 = u01(xb)
    r = np.sqrt(F32(-2.0) * _log_f32(u1))
    s, c = _sincos_2pi_f32(u2)
    return (r * c).astype(F32), (r * s).astype(F32)


def normals4(seed, ids, tick, purpose):
    """Four N(0,1) float32 per id: (z0,z1) from words (0,1), (z2,z3) from words (2,3)."""
    w0, w1, w2, w3 = draw(seed, ids, tick, purpose)
    z0, z1 = normal_pair(w0, w1)
    z2, z3 = normal_pair(w2, w3)
    return z0, z1, z2, z3


def uniform_sym(x, bound):
    """U(-bound, bound) float32 from a uint32 word: (2*u01-1)*bound."""
    return ((u01(x) * F32(2.0)) - F32(1.0)) * F32(bound)


def index_from_word(x, n):
    """Uniform integer in [0, n) by 64-bit multiply-shift (bias <= n / 2**32)."""
    return ((np.asarray(x, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def replay_indices(seed, ring_ids, update_tick, batch, record_range):
    """Replay sample indices for each ring: shape [len(ring_ids), batch], int64.
    Word j%4 of the Philox block with id = ring_id*(batch/4) + j//4."""
    assert batch % 4 == 0
    ring_ids = np.asarray(ring_ids, dtype=np.uint64).reshape(-1, 1)
    blocks = np.arange(batch // 4, dtype=np.uint64).reshape(1, -1)
    ids = ring_ids * np.uint64(batch // 4) + blocks
    words = draw(seed, ids, update_tick, PURPOSE_REPLAY)          # 4 x [R, batch/4]
    w = np.stack(words, axis=-1).reshape(ring_ids.shape[0], batch)  # j = 4*block + word
    return index_from_word(w, record_range)
