"""CPU restatement of the reference's training hot loop for a small population, driven by the counter-based draws of the
CUDA path.  TEST INFRASTRUCTURE ONLY (tests/test_gpu_loop.py): never imported by the product.

Follows workers/trainer.py:251-271 step by step:
    advance_environment (282-302): actor(prev_state) -> policy(+OU, clip) -> Platoon.step(actions, leader_exog ~ N(0, reset_max_u))
    train_all_models   (304-359): ReplayBuffer.add; once buffer_counter > batch_size: sample -> learn -> Adam x2 -> update_target
    federated gradients round (400-431) when an FRL method is configured: mean (or weighted mean) of the members' gradients per
    system, applied by every member's own Adam, then update_target -- instead of the local update (345).
Each piece is one of the pinned restatements: oracle/platoon_np.py (env, float64, bit-exact vs the reference), oracle/ddpg_np.py
(learn / Adam / Polyak / FedAvg), oracle/philox_env_np.py + philox_np.py (the draws).  Agent a = m*G + g is follower m of
platoon-group g; with E environments per group its minibatch is the concatenation of its E platoons' 64-row samples, which is the
reference's interfrl round over E platoons that share weights.
"""
from __future__ import annotations

import numpy as np

from . import ddpg_np as D
from . import philox_env_np as penv
from . import philox_np as ph
from . import platoon_np as onp

F32 = np.float32


class TrainLoopOracle:
    def __init__(self, conf, G, E, M, nets, *, seed=1, ring_capacity=64, fed=None):
        """nets: list over agents a = m*G + g of [actor, critic, target_actor, target_critic] parameter dicts (copied).
        fed: None (local updates), "interfrl" or "intrafrl" (gradients mode, every step, unweighted)."""
        self.conf, self.G, self.E, self.M, self.P = conf, G, E, M, G * E
        self.prm = onp.EnvParams.from_config(conf)
        self.seed, self.cap, self.fed = seed, ring_capacity, fed
        self.batch = int(conf.batch_size)
        self.nets = [[{k: v.copy() for k, v in n.items()} for n in four] for four in nets]
        z = lambda p, names: {k: np.zeros_like(p[k]) for k in names}
        self.opt = [dict(am=z(n[0], D.ACTOR_TRAINABLE), av=z(n[0], D.ACTOR_TRAINABLE), cm=z(n[1], D.CRITIC_TRAINABLE),
                         cv=z(n[1], D.CRITIC_TRAINABLE), t=0) for n in self.nets]
        x0, fa, fu = penv.reset_draws(self.prm, self.P, M, seed)
        self.env = onp.BatchedPlatoons(self.P, M, self.prm)
        self.obs = self.env.set_state(x0.astype(np.float64), front_accel=fa)          # [P, M, 4]
        self.ou = np.zeros((self.P, M), dtype=F32)
        self.ring = np.zeros((ring_capacity, M, self.P, 10), dtype=F32)
        self.count, self.tick, self.update_tick = 0, 0, 0
        self.losses = []          # per learn step: [A][2]
        self.rewards = []         # per step: [P, M]

    def agent(self, m, g):
        return m * self.G + g

    def step(self):
        conf, prm, G, E, M, P = self.conf, self.prm, self.G, self.E, self.M, self.P
        s_prev = self.obs.astype(F32)                                                  # what the fp32 device state holds
        mu = np.zeros((P, M), dtype=F32)
        for m in range(M):
            for g in range(G):
                rows = slice(g * E, (g + 1) * E)
                out, _ = D.actor_forward(self.nets[self.agent(m, g)][0], s_prev[rows, m], conf.action_high)      # trainer.py:287-289
                mu[rows, m] = out[:, 0]
        self.ou, _ = penv.ou_advance(prm, self.ou, self.seed, M, 0, tick=self.tick)
        act = penv.noisy_clipped_action(prm, mu, self.ou)                              # ddpgagent.py:18-29
        obs2, rew, done = self.env.step(act, penv.leader_exog(prm, P, self.seed, 0, tick=self.tick))      # trainer.py:291-296
        slot = self.count % self.cap                                                   # replaybuffer.py:40-47
        rec = self.ring[slot]
        rec[..., 0:4] = np.transpose(s_prev, (1, 0, 2))
        rec[..., 4] = act.T
        rec[..., 5] = rew.T.astype(F32)
        rec[..., 6:10] = np.transpose(obs2.astype(F32), (1, 0, 2))
        self.count += 1
        self.tick += 1
        self.obs = obs2
        self.rewards.append(rew.copy())
        if self.count > self.batch:                                                    # trainer.py:322
            self._learn()
        return obs2, rew, done

    def _learn(self):
        conf, G, E, M, P = self.conf, self.G, self.E, self.M, self.P
        n = min(self.count, self.cap)
        idx = ph.replay_indices(self.seed, np.arange(M * P), self.update_tick, self.batch, n)          # [M*P, batch]
        self.update_tick += 1
        grads, losses = [None] * (M * G), []
        for m in range(M):
            for g in range(G):
                a = self.agent(m, g)
                rows = []
                for e in range(E):
                    p = g * E + e
                    rows.append(self.ring[idx[m * P + p], m, p])                         # [batch, 10]
                b = np.concatenate(rows, axis=0)
                batch = (b[:, 0:4], b[:, 4:5], b[:, 5:6], b[:, 6:10])
                ac, cr, ta, tc = self.nets[a]
                cg, ag, info = D.learn(ac, cr, ta, tc, batch, gamma=conf.gamma, high=conf.action_high)      # trainer.py:325-326
                grads[a] = (cg, ag)
                losses.append((info["critic_loss"], info["actor_loss"]))
        self.losses.append(losses)
        if self.fed is not None:                                                       # trainer.py:400-431 (unweighted: federated.py:47-63)
            systems = ([[self.agent(m, g) for g in range(G)] for m in range(M)] if self.fed == "interfrl"
                       else [[self.agent(m, g) for m in range(M)] for g in range(G)])
            for members in systems:
                cavg = {k: np.mean(np.stack([grads[a][0][k] for a in members]), axis=0).astype(F32) for k in D.CRITIC_TRAINABLE}
                aavg = {k: np.mean(np.stack([grads[a][1][k] for a in members]), axis=0).astype(F32) for k in D.ACTOR_TRAINABLE}
                for a in members:
                    grads[a] = (cavg, aavg)
        for a in range(M * G):
            ac, cr, ta, tc = self.nets[a]
            o = self.opt[a]
            o["t"] += 1
            D.adam_apply(cr, grads[a][0], o["cm"], o["cv"], o["t"], conf.critic_lr, D.CRITIC_TRAINABLE)      # trainer.py:348-349 / 420-425
            D.adam_apply(ac, grads[a][1], o["am"], o["av"], o["t"], conf.actor_lr, D.ACTOR_TRAINABLE)
            self.nets[a][3] = D.polyak(tc, cr, conf.tau, D.CRITIC_WEIGHTS)              # trainer.py:352-356 / 428-431
            self.nets[a][2] = D.polyak(ta, ac, conf.tau, D.ACTOR_WEIGHTS)
