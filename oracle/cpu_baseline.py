"""CPU timing of the reference algorithm's hot loop (the ``cpu_baseline`` / ``--impl reference`` legs of
bench.py).  TEST/BENCH INFRASTRUCTURE ONLY -- never on the product path.

/root/reference is a Python tree that does not exist on the GPU box, so what is timed here is the
*port* in oracle/platoon_np.py (``SerialPlatoon`` etc.), which reproduces the reference bit for bit
(tests/test_oracle_env_vs_golden.py) with the same per-object structure: a Python loop over platoons and
followers, three tiny NumPy ``dot``s per vehicle, one global-RNG draw per OU sample, one ring write per
agent -- i.e. workers/trainer.py:282-296 (minus the actor forward) + 316-319.  The reference is
single-threaded by construction (src/rand.py:14-15); "all cores" replicates independent platoons over
processes, which is the most favourable reading of how it could use a multi-core host.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import platoon_np as onp


def _env_worker(args):
    M, steps, seed, with_replay = args
    prm = onp.EnvParams()
    np.random.seed(seed)
    pl = onp.SerialPlatoon(M, prm)
    ous = [onp.SerialOUNoise(prm) for _ in range(M)]
    rbs = [onp.SerialReplay(4096, 64, 4, 1) for _ in range(M)] if with_replay else None
    prev = pl.reset()
    mu = np.zeros(M)
    t0 = time.perf_counter()
    done_steps = 0
    for k in range(steps):
        acts = [np.squeeze(onp.clip_action(mu[m], ous[m](), prm.action_low, prm.action_high)) for m in range(M)]
        st, rw, dn = pl.step(acts, np.random.normal(0, prm.reset_max_u))
        if rbs is not None:
            for m in range(M):
                rbs[m].add(prev[m], acts[m], rw[m], st[m])
        prev = st
        done_steps += 1
        if dn:
            prev = pl.reset()
    return done_steps, time.perf_counter() - t0


class EnvLoopPool:
    """Persistent worker pool so a multi-step reference-arm run pays process start-up once."""

    def __init__(self, M: int = 4, cores: int | None = None, with_replay: bool = True):
        self.M, self.with_replay = M, with_replay
        self.cores = cores or os.cpu_count() or 1
        n0, t0 = _env_worker((M, 300, 1, with_replay))
        self.per_step = t0 / n0                      # single-core seconds per platoon-step
        self.pool = mp.get_context("fork").Pool(self.cores) if self.cores > 1 else None
        self._round = 0

    def run(self, seconds: float):
        """Every worker steps its own platoon for ~seconds.  -> (vehicle-steps/s aggregate, steps per worker, slowest s)."""
        steps = max(50, int(seconds / self.per_step))
        jobs = [(self.M, steps, 1 + self._round * self.cores + i, self.with_replay) for i in range(self.cores)]
        self._round += 1
        res = self.pool.map(_env_worker, jobs) if self.pool else [_env_worker(jobs[0])]
        slowest = max(r[1] for r in res)
        return sum(r[0] for r in res) * self.M / slowest, steps, slowest

    def describe(self, steps, slowest):
        return (f"{self.cores} process(es) x 1 platoon x {self.M} followers x {steps} steps of act(OU+clip)+Platoon.step"
                f"{'+ReplayBuffer.add' if self.with_replay else ''} (float64 NumPy port of the reference loop, {slowest:.1f} s each)")

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def time_env_steps(M: int = 4, target_seconds: float = 12.0, cores: int | None = None, with_replay: bool = True):
    """-> dict(value=vehicle env-steps/s over all workers, cores=..., sample=...)."""
    pool = EnvLoopPool(M, cores, with_replay)
    try:
        v, steps, slowest = pool.run(target_seconds)
    finally:
        pool.close()
    return dict(value=v, unit="platoon-vehicle env-steps/s", cores=pool.cores, kind="port", sample=pool.describe(steps, slowest),
                single_core_vehicle_steps_per_s=M / pool.per_step)
