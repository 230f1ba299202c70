"""The reference's training loop written against the PER-OBJECT drop-ins of this package -- the proof that
``environment.Platoon``, ``noise.OUActionNoise``, ``replaybuffer.ReplayBuffer``, ``model.get_actor / get_critic``,
``ddpgagent.policy / update_target``, ``trainer.Trainer.learn``, ``optimizers.Adam`` and ``server.federated.Server`` compose the way
workers/trainer.py composes the originals (initialize 61-179, run 223-280, advance_environment 282-302, train_all_models 304-359,
train_all_models_federated_gradients 400-431, train_all_models_federated_weights 433-456).

/root/reference does not exist on the GPU box and this package has no CPU path, so the reference's own workers/trainer.py cannot be
executed against these modules anywhere; this class makes the same calls in the same order on the same kind of objects (one
Platoon per platoon, one actor / critic / target pair, optimizer pair, OU process and replay buffer per (platoon, follower)) and
tests/test_gpu_dropin.py checks it against the batched loop.  It is the slow path by construction: one tiny launch per object and step.

Only what the hot loop needs is here: no CSV / plotting / .h5 / rendering (out of scope).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import ddpgagent, environment, model, noise, optimizers, replaybuffer
from .server import federated
from .trainer import (Trainer, is_fed_enabled, is_valid_step_for_federated_training_with_gradients,
                      is_valid_step_for_federated_training_with_weights, is_valid_update_step, is_weighted_fed_enabled)


class PerObjectTrainer:
    def __init__(self, conf, leader_exog_fn: Optional[Callable[[int, int], float]] = None):
        """leader_exog_fn(platoon, step) -> the leader's exogenous input; default: N(0, reset_max_u) from NumPy like
        util.get_random_val (workers/trainer.py:292-295)."""
        self.conf = conf
        self.P, self.M = int(conf.num_platoons), int(conf.pl_size)
        self.low_bound, self.high_bound = conf.action_low, conf.action_high
        self.leader_exog_fn = leader_exog_fn or (lambda p, k: float(np.random.normal(0, conf.reset_max_u)))
        self.learner = Trainer(conf=conf)
        self.step_count = 0

    # -------------------------------------------------------------------------------------------- trainer.py:61-179
    def initialize(self):
        conf, P, M = self.conf, self.P, self.M
        conf.fed_enabled = is_fed_enabled(conf)
        self.fed_server = federated.Server("AVDDPG", False) if conf.fed_enabled else None
        self.envs, self.ou, self.actors, self.critics, self.t_actors, self.t_critics = [], [], [], [], [], []
        self.actor_opt, self.critic_opt, self.rbuffers, self.ep_rewards = [], [], [], []
        init_a = init_c = None
        for p in range(P):
            env = environment.Platoon(M, conf, p, rand_states=conf.rand_states)
            self.envs.append(env)
            row = {k: [] for k in ("ou", "a", "c", "ta", "tc", "ao", "co", "rb")}
            for m in range(M):
                row["ou"].append(noise.OUActionNoise(mean=np.zeros(1), config=conf, stream_id=p * M + m))
                a = model.get_actor(env.num_states, env.num_actions, self.high_bound, seed_int=conf.random_seed, hidd_mult=env.hidden_multiplier,
                                    layer1_size=conf.actor_layer1_size, layer2_size=conf.actor_layer2_size)
                c = model.get_critic(env.num_states, env.num_actions, hidd_mult=env.hidden_multiplier, layer1_size=conf.critic_layer1_size,
                                     layer2_size=conf.critic_layer2_size, action_layer_size=conf.critic_act_layer_size)
                ta = model.get_actor(env.num_states, env.num_actions, self.high_bound, seed_int=conf.random_seed, hidd_mult=env.hidden_multiplier,
                                     layer1_size=conf.actor_layer1_size, layer2_size=conf.actor_layer2_size)
                tc = model.get_critic(env.num_states, env.num_actions, hidd_mult=env.hidden_multiplier, layer1_size=conf.critic_layer1_size,
                                      layer2_size=conf.critic_layer2_size, action_layer_size=conf.critic_act_layer_size)
                if init_a is None:                       # every (p, m) starts from platoon 0 / model 0 (121-128)
                    init_a, init_c = a.get_weights(), c.get_weights()
                else:
                    a.set_weights(init_a)
                    c.set_weights(init_c)
                ta.set_weights(a.get_weights())          # 130-131
                tc.set_weights(c.get_weights())
                row["a"].append(a); row["c"].append(c); row["ta"].append(ta); row["tc"].append(tc)
                row["co"].append(optimizers.Adam(conf.critic_lr))        # 138-139
                row["ao"].append(optimizers.Adam(conf.actor_lr))
                row["rb"].append(replaybuffer.ReplayBuffer(conf.buffer_size, conf.batch_size, env.num_states, env.num_actions, conf.pl_size,
                                                           seed=conf.random_seed, ring_id=m * P + p))
            self.ou.append(row["ou"]); self.actors.append(row["a"]); self.critics.append(row["c"])
            self.t_actors.append(row["ta"]); self.t_critics.append(row["tc"])
            self.actor_opt.append(row["ao"]); self.critic_opt.append(row["co"]); self.rbuffers.append(row["rb"])
            self.ep_rewards.append([[] for _ in range(M)])
        self._reset_fed_lists()
        self.actions = np.zeros((P, M, 1))
        return self

    def _reset_fed_lists(self):
        S, X = (self.M, self.P) if self.conf.fed_method == self.conf.interfrl else (self.P, self.M)
        mk = lambda: [[None] * X for _ in range(S)]
        self.actor_grads, self.critic_grads, self.actor_weights, self.critic_weights = mk(), mk(), mk(), mk()
        self.fed_weights = np.zeros((S, X), dtype=np.float32)

    # -------------------------------------------------------------------------------------------- trainer.py:223-280
    def run(self, episodes: int, steps_per_episode: Optional[int] = None):
        conf = self.conf
        spe = int(conf.steps_per_episode if steps_per_episode is None else steps_per_episode)
        for ep in range(int(episodes)):
            self.counters = [np.zeros(self.M, dtype=np.float32) for _ in range(self.P)]
            prev = [env.reset() for env in self.envs]
            for i in range(spe):
                states, rewards, terminals = [], [], []
                for p in range(self.P):
                    s, r, t = self.advance_environment(p, prev)
                    states.append(s); rewards.append(r); terminals.append(t)
                self.train_all_models(rewards, states, prev, ep, i)
                if is_valid_step_for_federated_training_with_gradients(conf, ep, i):
                    self.federated_gradients(ep)
                if is_valid_step_for_federated_training_with_weights(conf, ep, i):
                    self.federated_weights(ep)
                self.step_count += 1
                if True in terminals:
                    break
                prev = states
            for p in range(self.P):
                for m in range(self.M):
                    self.ep_rewards[p][m].append(self.counters[p][m])
        return self

    # -------------------------------------------------------------------------------------------- trainer.py:282-302
    def advance_environment(self, p, prev):
        for m in range(self.M):
            out = self.actors[p][m](np.asarray(prev[p][m], dtype=np.float32)[None])
            self.actions[p][m] = ddpgagent.policy(out, self.ou[p][m], self.low_bound, self.high_bound)[0]
        return self.envs[p].step(self.actions[p].flatten(), self.leader_exog_fn(p, self.step_count))

    # -------------------------------------------------------------------------------------------- trainer.py:304-359
    def train_all_models(self, rewards, states, prev, ep, i):
        conf = self.conf
        for p in range(self.P):
            for m in range(self.M):
                rb = self.rbuffers[p][m]
                rb.add((prev[p][m], self.actions[p][m], rewards[p][m], states[p][m]))
                self.counters[p][m] += np.float32(rewards[p][m])
                if rb.buffer_counter <= conf.batch_size:
                    continue
                critic_grad, actor_grad = self.learner.learn(rb, self.actors[p][m], self.critics[p][m], self.t_actors[p][m], self.t_critics[p][m])
                if conf.fed_enabled:
                    w = self._weight(p, m) if is_weighted_fed_enabled(conf, ep) else None
                    s, x = (m, p) if conf.fed_method == conf.interfrl else (p, m)
                    scale = (lambda ts: [t * w for t in ts]) if w is not None else (lambda ts: list(ts))
                    self.actor_grads[s][x], self.critic_grads[s][x] = scale(actor_grad), scale(critic_grad)
                    self.actor_weights[s][x], self.critic_weights[s][x] = scale(self.actors[p][m].weights), scale(self.critics[p][m].weights)
                    if w is not None:
                        self.fed_weights[s][x] = w
                if not conf.fed_enabled or not is_valid_update_step(conf, i):      # 345: local update unless FRL step
                    self.critic_opt[p][m].apply_gradients(zip(critic_grad, self.critics[p][m].trainable_variables))
                    self.actor_opt[p][m].apply_gradients(zip(actor_grad, self.actors[p][m].trainable_variables))
                    self._soft_update(p, m)

    def _weight(self, p, m):       # trainer.py:385-398
        return np.float32(abs(1 / np.mean(self.ep_rewards[p][m][-self.conf.weighted_window:])))

    def _soft_update(self, p, m):  # trainer.py:352-356 / 428-431
        tc, ta = ddpgagent.update_target(self.conf.tau, self.t_critics[p][m].weights, self.critics[p][m].weights,
                                         self.t_actors[p][m].weights, self.actors[p][m].weights)
        self.t_actors[p][m].set_weights(ta)
        self.t_critics[p][m].set_weights(tc)

    def _filled(self):
        return all(rb.buffer_counter > self.conf.batch_size for row in self.rbuffers for rb in row)

    def _avg(self, lists, ep):
        if is_weighted_fed_enabled(self.conf, ep):
            return self.fed_server.get_weighted_avg_params(lists, self.fed_weights.sum(axis=1))
        return self.fed_server.get_avg_params(lists)

    # -------------------------------------------------------------------------------------------- trainer.py:400-431
    def federated_gradients(self, ep):
        if not self._filled():
            return
        conf = self.conf
        a_avg, c_avg = self._avg(self.actor_grads, ep), self._avg(self.critic_grads, ep)
        for p in range(self.P):
            for m in range(self.M):
                if conf.fed_method == conf.intrafrl and m == 0 and conf.intra_directional_averaging:
                    continue
                s = m if conf.fed_method == conf.interfrl else p
                self.actor_opt[p][m].apply_gradients(zip(a_avg[s], self.actors[p][m].trainable_variables))
                self.critic_opt[p][m].apply_gradients(zip(c_avg[s], self.critics[p][m].trainable_variables))
                self._soft_update(p, m)

    # -------------------------------------------------------------------------------------------- trainer.py:433-456
    def federated_weights(self, ep):
        if not self._filled():
            return
        conf = self.conf
        a_avg, c_avg = self._avg(self.actor_weights, ep)[0], self._avg(self.critic_weights, ep)[0]       # `[0]`: system 0's average for everyone
        for p in range(self.P):
            for m in range(self.M):
                if conf.fed_method == conf.intrafrl and m == 0 and conf.intra_directional_averaging:
                    continue
                for net, avg in ((self.actors, a_avg), (self.critics, c_avg), (self.t_actors, a_avg), (self.t_critics, c_avg)):
                    net[p][m].set_weights(avg)
