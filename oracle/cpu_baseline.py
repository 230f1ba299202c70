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

from . import ddpg_np as D
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


def _train_worker(args):
    """The reference's full hot loop for ONE platoon (workers/trainer.py:251-271): per step and follower
    actor(state) -> policy(+OU, clip) -> Platoon.step -> ReplayBuffer.add -> [sample -> learn -> Adam x2 -> Polyak]."""
    M, steps, seed, _ = args
    prm = onp.EnvParams()
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    pl = onp.SerialPlatoon(M, prm)
    ous = [onp.SerialOUNoise(prm) for _ in range(M)]
    rbs = [onp.SerialReplay(100000, 64, 4, 1) for _ in range(M)]
    nets = []
    for m in range(M):
        ac, cr = D.init_actor(rng), D.init_critic(rng)
        z = lambda p, names: {k: np.zeros_like(p[k]) for k in names}
        nets.append(dict(ac=ac, cr=cr, ta={k: v.copy() for k, v in ac.items()}, tc={k: v.copy() for k, v in cr.items()},
                         am=z(ac, D.ACTOR_TRAINABLE), av=z(ac, D.ACTOR_TRAINABLE), cm=z(cr, D.CRITIC_TRAINABLE),
                         cv=z(cr, D.CRITIC_TRAINABLE), t=0))
    for rb in rbs:                      # pre-fill past batch_size so every timed step learns (steady state)
        for _ in range(65):
            rb.add(rng.normal(size=4), rng.normal(size=1), -rng.random(), rng.normal(size=4))
    prev = pl.reset()
    t0 = time.perf_counter()
    for k in range(steps):
        acts = []
        for m in range(M):
            mu, _ = D.actor_forward(nets[m]["ac"], np.asarray(prev[m], np.float32)[None])
            acts.append(np.squeeze(onp.clip_action(np.squeeze(mu), ous[m](), prm.action_low, prm.action_high)))
        st, rw, dn = pl.step(acts, np.random.normal(0, prm.reset_max_u))
        for m in range(M):
            rbs[m].add(prev[m], acts[m], rw[m], st[m])
            n = nets[m]
            cg, ag, _ = D.learn(n["ac"], n["cr"], n["ta"], n["tc"], rbs[m].sample())
            n["t"] += 1
            D.adam_apply(n["cr"], cg, n["cm"], n["cv"], n["t"], 5e-4, D.CRITIC_TRAINABLE)
            D.adam_apply(n["ac"], ag, n["am"], n["av"], n["t"], 5e-5, D.ACTOR_TRAINABLE)
            n["tc"] = D.polyak(n["tc"], n["cr"], 0.001, D.CRITIC_WEIGHTS)
            n["ta"] = D.polyak(n["ta"], n["ac"], 0.001, D.ACTOR_WEIGHTS)
        prev = pl.reset() if dn else st
    return steps, time.perf_counter() - t0


class EnvLoopPool:
    """Persistent worker pool so a multi-step reference-arm run pays process start-up once."""

    def __init__(self, M: int = 4, cores: int | None = None, with_replay: bool = True, with_learn: bool = False):
        self.M, self.with_replay, self.with_learn = M, with_replay, with_learn
        self.worker = _train_worker if with_learn else _env_worker
        self.cores = cores or os.cpu_count() or 1
        # one single-threaded process per core (the reference pins TF to 1 thread, src/rand.py:14-15); "spawn" so the
        # children start with BLAS limited to 1 thread instead of inheriting a forked multi-threaded BLAS state
        saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
        os.environ.update({k: "1" for k in saved})
        try:
            self.pool = mp.get_context("spawn").Pool(self.cores)
            n0, t0 = self.pool.apply(self.worker, ((M, 20 if with_learn else 300, 1, with_replay),))
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        self.per_step = t0 / n0                      # single-core seconds per platoon-step
        self._round = 0

    def run(self, seconds: float):
        """Every worker steps its own platoon for ~seconds.  -> (vehicle-steps/s aggregate, steps per worker, slowest s)."""
        steps = max(5 if self.with_learn else 50, int(seconds / self.per_step))
        jobs = [(self.M, steps, 1 + self._round * self.cores + i, self.with_replay) for i in range(self.cores)]
        self._round += 1
        res = self.pool.map(self.worker, jobs)
        slowest = max(r[1] for r in res)
        return sum(r[0] for r in res) * self.M / slowest, steps, slowest

    def describe(self, steps, slowest):
        what = ("actor fwd+OU+clip+Platoon.step+ReplayBuffer.add+sample+learn(batch 64)+Adam x2+Polyak per agent" if self.with_learn
                else "act(OU+clip)+Platoon.step" + ("+ReplayBuffer.add" if self.with_replay else ""))
        return (f"{self.cores} process(es) x 1 platoon x {self.M} followers x {steps} steps of {what} "
                f"(NumPy port of the reference loop, 1 thread per process, {slowest:.1f} s each)")

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def time_env_steps(M: int = 4, target_seconds: float = 12.0, cores: int | None = None, with_replay: bool = True,
                   with_learn: bool = False):
    """-> dict(value=vehicle env-steps/s over all workers, cores=..., sample=...)."""
    pool = EnvLoopPool(M, cores, with_replay, with_learn)
    try:
        v, steps, slowest = pool.run(target_seconds)
    finally:
        pool.close()
    return dict(value=v, unit="platoon-vehicle env-steps/s", cores=pool.cores, kind="port", sample=pool.describe(steps, slowest),
                single_core_vehicle_steps_per_s=M / pool.per_step)
