"""Noise-free evaluation rollout -- the arithmetic of the reference's ``workers/evaluator.py:40-96,145`` on the GPU
(SURVEY.md §8f N2); plotting, .h5 loading and figure saving are out of scope.

``run`` follows the reference step by step: seed with ``evaluation_seed``, build the platoon(s) with the fixed evaluator
initial state ``[reset_ep_eval_max, reset_ev_eval_max, reset_a_eval_max, a_lead]`` (environment.py:534-539), pre-draw the
leader's exogenous inputs ~ N(0, reset_max_u) for every time step (evaluator.py:55-56), then loop
``policy(actor(state))`` without noise (81-83) -> ``env.step`` (84) -> ``get_jerk`` (85) while accumulating float32
episodic rewards (93), and return ``round(mean(episodic rewards), 3)`` (145, 158).  It is batched over P platoons that
share the actors, so one call evaluates many leader-input realisations at once.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .environment import BatchedPlatoons
from .trainer import DDPGPopulation


def get_number_of_timesteps_for_plot(conf, manual_timestep_override: Optional[int]):
    return conf.steps_per_episode if manual_timestep_override is None else manual_timestep_override


def run(conf, population: DDPGPopulation, *, num_platoons: int = 1, leader_inputs=None, manual_timestep_override=None,
        seed: bool = True, precision: Optional[int] = None, return_traces: bool = False):
    """-> pl_rew (float, like the reference) or (pl_rew, traces) with traces = dict(states[T,P,M,ns], inputs[T,P,M],
    jerks[T,P,M], rewards[P,M]).  `population` must have G == 1 (the platoons share the M actors).
    leader_inputs: optional [T, P] array (e.g. the reference's own pre-drawn list) -- otherwise Philox draws keyed by
    evaluation_seed are used."""
    _lib.require_device()
    if population.G != 1:
        raise ValueError("evaluation shares one set of actors: population.G must be 1")
    M, P = population.M, int(num_platoons)
    T = get_number_of_timesteps_for_plot(conf, manual_timestep_override)
    ev_seed = int(conf.evaluation_seed) if seed else int(getattr(conf, "random_seed", 1))
    env = BatchedPlatoons(P, M, conf, seed=ev_seed, evaluator_states_enabled=True, track_kinematics=True, collect_stats=False)
    env.reset()
    old_prec = population.precision
    if precision is not None:
        population.precision = int(precision)
    ep_reward = torch.zeros(P, M, dtype=torch.float32, device=env.device)     # np.float32 counters (evaluator.py:67)
    traces = None
    if return_traces:
        traces = dict(states=torch.zeros(T, P, M, env.num_states, device=env.device), inputs=torch.zeros(T, P, M, device=env.device),
                      jerks=torch.zeros(T, P, M, device=env.device))
    if leader_inputs is not None:
        li = torch.as_tensor(np.asarray(leader_inputs, dtype=np.float32), device=env.device).reshape(T, -1).expand(T, P).contiguous()
    try:
        for i in range(T):
            population.act(env.native_state, env.action_mu, envs_per_group=P)
            if leader_inputs is not None:
                env.leader_exog.copy_(li[i])
                env.step_native(explore=False, clip=True, leader_exog=True)
            else:
                env.step_native(explore=False, clip=True, gen_exog=True)
            ep_reward += env.reward
            if return_traces:
                traces["states"][i] = env.obs
                traces["inputs"][i] = env.action_out.t()
                traces["jerks"][i] = env.jerk.t()
    finally:
        population.precision = old_prec
    per_platoon = ep_reward.mean(dim=1)                      # np.average(episodic_reward_counters), evaluator.py:145
    pl_rew = round(float(per_platoon.mean().item()), 3)
    if return_traces:
        traces["rewards"] = ep_reward
        traces["pl_rew_per_platoon"] = per_platoon
        return pl_rew, traces
    return pl_rew
