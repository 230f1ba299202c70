"""Host restatement of how the CUDA path maps Philox draws onto resets, OU noise and leader inputs.
TEST INFRASTRUCTURE ONLY.

The *distributions* are the reference's (src/environment.py:284-301, 520-559; src/noise.py:14-23;
workers/trainer.py:292-295; src/util.py:55-70); the *stream* is the counter-based one defined in
oracle/philox_np.py because the reference's global MT19937 order cannot be reproduced in parallel
(SURVEY.md §7).  All arithmetic is float32 with one rounding per operation, so the CUDA kernels
(csrc/avd_env.cu, csrc/avd_rng.cuh) must match these functions BIT FOR BIT.
"""
from __future__ import annotations

import numpy as np

from . import philox_np as ph
from .platoon_np import EnvParams

F32 = np.float32


def _first(words, bound, uniform):
    w0, w1 = words
    if uniform:
        return ph.uniform_sym(w0, bound)
    z0, _ = ph.normal_pair(w0, w1)
    return z0 * F32(bound)


def reset_draws(prm: EnvParams, P, M, seed, platoon_id_base=0, episode=0, reset_mode=0, fixed=None):
    """-> x0[P,M,4] float32, front_accel[P], front_u[P].  ``episode`` may be an int or an array[P]."""
    uni = prm.rand_gen == "uniform"
    gp = np.arange(P, dtype=np.uint64) + np.uint64(platoon_id_base)
    ep = np.broadcast_to(np.asarray(episode, dtype=np.uint64), (P,))
    wp = ph.draw(seed, gp, ep, ph.PURPOSE_RESET_PLATOON)
    if uni:
        fa = ph.uniform_sym(wp[0], prm.pl_leader_reset_a)
        fu = ph.uniform_sym(wp[1], prm.reset_max_u)
    else:
        z0, z1 = ph.normal_pair(wp[0], wp[1])
        fa, fu = z0 * F32(prm.pl_leader_reset_a), z1 * F32(prm.reset_max_u)
    x = np.zeros((P, M, 4), dtype=F32)
    if reset_mode == 0:
        gv = gp[:, None] * np.uint64(M) + np.arange(M, dtype=np.uint64)[None, :]
        wv = ph.draw(seed, gv, ep[:, None], ph.PURPOSE_RESET_VEHICLE)
        if uni:
            x[..., 0] = ph.uniform_sym(wv[0], prm.reset_ep_max)
            x[..., 1] = ph.uniform_sym(wv[1], prm.reset_max_ev)
            x[..., 2] = ph.uniform_sym(wv[2], prm.reset_max_a)
        else:
            z0, z1 = ph.normal_pair(wv[0], wv[1])
            z2, _ = ph.normal_pair(wv[2], wv[3])
            x[..., 0] = z0 * F32(prm.reset_ep_max)
            x[..., 1] = z1 * F32(prm.reset_max_ev)
            x[..., 2] = z2 * F32(prm.reset_max_a)
    else:
        x[..., 0], x[..., 1], x[..., 2] = (F32(v) for v in fixed)
    x[:, 0, 3] = fa
    x[:, 1:, 3] = x[:, :-1, 2]
    return x, fa.astype(F32), fu.astype(F32)


def ou_advance(prm: EnvParams, state, seed, M, platoon_id_base=0, tick=0, mean=0.0):
    """state[P,M] float32 -> new state, float32 ops in the kernel's order."""
    state = np.asarray(state, dtype=F32)
    P = state.shape[0]
    gv = (np.arange(P, dtype=np.uint64)[:, None] + np.uint64(platoon_id_base)) * np.uint64(M) + np.arange(M, dtype=np.uint64)[None, :]
    w = ph.draw(seed, gv, tick, ph.PURPOSE_OU)
    z0, _ = ph.normal_pair(w[0], w[1])
    c = F32(prm.std_dev) * np.sqrt(F32(prm.ou_dt))
    drift = (F32(prm.theta) * (F32(mean) - state)) * F32(prm.ou_dt)
    return ((state + drift) + c * z0).astype(F32), z0


def noisy_clipped_action(prm: EnvParams, mu, noise):
    return np.clip(np.asarray(mu, dtype=F32) + np.asarray(noise, dtype=F32), F32(prm.action_low), F32(prm.action_high)).astype(F32)


def leader_exog(prm: EnvParams, P, seed, platoon_id_base=0, tick=0):
    gp = np.arange(P, dtype=np.uint64) + np.uint64(platoon_id_base)
    w = ph.draw(seed, gp, tick, ph.PURPOSE_LEADER_EXOG)
    return _first((w[0], w[1]), prm.reset_max_u, prm.rand_gen == "uniform").astype(F32)
