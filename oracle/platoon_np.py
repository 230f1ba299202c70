"""CPU restatement (float64 NumPy) of the reference platoon environment.  TEST INFRASTRUCTURE ONLY.

Follows, function by function:
  * discretised dynamics matrices ........ /root/reference/src/environment.py:390-451
  * one vehicle step (reward, terminal,
    kinematic observables, x <- Ax+Bu+Ce) . /root/reference/src/environment.py:460-518
  * platoon step / exogenous wiring ...... /root/reference/src/environment.py:209-241, 253-269
  * centralized reward ................... /root/reference/src/environment.py:271-282
  * reset ................................ /root/reference/src/environment.py:284-301, 520-559
  * OU exploration noise ................. /root/reference/src/noise.py:14-29
  * action clip .......................... /root/reference/agent/ddpgagent.py:18-29
  * replay ring .......................... /root/reference/src/replaybuffer.py:31-63

Two statements of the same algorithm live here:
  ``SerialPlatoon``  one platoon, follower-by-follower, drawing from NumPy's *global legacy*
                     RNG in exactly the reference's draw order -- so with the same
                     ``np.random.seed`` it reproduces the reference bit for bit (this is how the
                     restatement is pinned: tests/test_oracle_vs_golden.py), and it is the
                     ``cpu_baseline`` "port" that bench.py times.
  ``BatchedPlatoons`` the same arithmetic vectorised over P platoons with *injected* actions,
                     exogenous inputs and initial states; used to check the CUDA path at sizes
                     where a Python loop over platoons would take minutes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

STANDSTILL_M = 8.0  # environment.py:343 (hard-coded standstill distance r)


@dataclass
class EnvParams:
    """The scalars the hot path reads from the reference Config (src/config.py)."""
    sample_rate: float = 0.1        # T           config.py:86
    timegap: float = 1.0            # h           config.py:49
    dyn_coeff: float = 0.1          # tau         config.py:50
    pl_leader_tau: float = 0.1      #             config.py:44
    method: str = "euler"           #             config.py:45-47
    model: str = "ModelB"           #             config.py:7
    framework: str = "decentralized"  #           config.py:26
    reward_ep_coeff: float = 0.4    #             config.py:52-55
    reward_ev_coeff: float = 0.2
    reward_u_coeff: float = 0.2
    reward_jerk_coeff: float = 0.2
    max_ep: float = 20.0            #             config.py:57-58
    max_ev: float = 20.0
    action_high: float = 2.5        #             config.py:68-69
    action_low: float = -2.5
    re_scalar: float = 1.0          #             config.py:71
    terminal_reward: float = 0.5    #             config.py:72
    can_terminate: bool = True      #             config.py:75
    reset_ep_max: float = 1.5       #             config.py:60-62
    reset_max_ev: float = 1.5
    reset_max_a: float = 0.05
    reset_ep_eval_max: float = 1.0  #             config.py:64-66
    reset_ev_eval_max: float = 1.0
    reset_a_eval_max: float = 0.03
    pl_leader_reset_a: float = 0.0  #             config.py:41
    reset_max_u: float = 0.1        #             config.py:42
    rand_gen: str = "normal"        #             config.py:81
    rand_states: bool = True        #             config.py:82
    std_dev: float = 0.02           # OU sigma    config.py:101
    theta: float = 0.15             #             config.py:102
    ou_dt: float = 1e-2             #             config.py:103

    @classmethod
    def from_config(cls, conf) -> "EnvParams":
        return cls(**{k: getattr(conf, k) for k in cls.__dataclass_fields__ if hasattr(conf, k)})


# --------------------------------------------------------------------------- system matrices
def system_matrices(T, h, tau, tau_lead, method):
    """A (4x4), B (4), C (4) of x_{k+1} = A x_k + B u_k + C w_k.  environment.py:390-451."""
    if method == "euler":
        A = np.array([[1.0, T, -h * T, 0.0],
                      [0.0, 1.0, -T, T],
                      [0.0, 0.0, 1.0 - T / tau, 0.0],
                      [0.0, 0.0, 0.0, 1.0 - T / tau_lead]])
        B = np.array([0.0, 0.0, T / tau, 0.0])
        C = np.array([0.0, 0.0, 0.0, T / tau_lead])
    elif method == "exact":
        e = np.exp(-T / tau)
        el = np.exp(-T / tau_lead)
        a13 = -h * tau + h * tau * e - tau * T + tau ** 2 - (tau ** 2) * e
        a14 = tau_lead * T - tau_lead ** 2 + (tau_lead ** 2) * el
        a23 = -tau + tau * e
        a24 = tau_lead - tau_lead * el
        A = np.array([[1.0, T, a13, a14],
                      [0.0, 1.0, a23, a24],
                      [0.0, 0.0, e, 0.0],
                      [0.0, 0.0, 0.0, el]])
        b11 = (-h * T + h * tau * e - h * tau - (T ** 2) / 2 + tau * T + (tau ** 2) * e - tau ** 2)
        b21 = -T - tau * e + tau
        B = np.array([b11, b21, -e + 1.0, 0.0])
        c11 = (T ** 2) / 2 - tau_lead * T - (tau_lead ** 2) * el + tau_lead ** 2
        c21 = T + tau_lead * el - tau_lead
        C = np.array([c11, c21, 0.0, -el + 1.0])
    else:
        raise ValueError(f"unknown discretisation {method!r}")
    return A, B, C


def follower_matrices(prm: EnvParams, M: int):
    """Per-follower (A,B,C): follower 0 sees the leader's tau, follower m>0 its predecessor's
    (environment.py:57,61; every follower's own tau is dyn_coeff, environment.py:347)."""
    out = []
    for m in range(M):
        tl = prm.pl_leader_tau if m == 0 else prm.dyn_coeff
        out.append(system_matrices(prm.sample_rate, prm.timegap, prm.dyn_coeff, tl, prm.method))
    return out


def num_states_of(prm: EnvParams) -> int:
    return 3 if prm.model == "ModelA" else 4   # environment.py:47-52


# --------------------------------------------------------------------------- RNG helper
def _draw(prm: EnvParams, bound, size=None, mode=None):
    """util.get_random_val (src/util.py:55-70) on the global legacy NumPy stream."""
    mode = prm.rand_gen if mode is None else mode
    if mode == "uniform":
        return np.random.uniform(-1 * bound, bound)
    return np.random.normal(0, bound, size=size)


# --------------------------------------------------------------------------- serial port
class SerialVehicle:
    def __init__(self, prm: EnvParams, idx: int, tau_lead: float, a_lead, rand_states=True,
                 eval_states=False):
        self.prm, self.idx = prm, idx
        self.tau = prm.dyn_coeff
        self.rand_states, self.eval_states = rand_states, eval_states
        self.cum_accel = 0.0
        self.velocity = self.desired_headway = self.headway = 0.0
        self.reward = 0.0
        self.u = 0.0
        self.exog = 0.0
        self.reset(a_lead)                                     # environment.py:385 (draws first)
        self.A, self.B, self.C = system_matrices(prm.sample_rate, prm.timegap, self.tau, tau_lead,
                                                 prm.method)

    def reset(self, a_lead):
        p = self.prm
        self.u = 0.0
        self.cum_accel = 0.0
        self.desired_headway = self.headway = 0.0
        self.jerk = 0.0
        if self.eval_states:                                   # environment.py:534-544
            head = ([p.reset_ep_eval_max, p.reset_ev_eval_max, p.reset_a_eval_max]
                    if self.rand_states else [p.reset_ep_max, p.reset_max_ev, p.reset_max_a])
        elif self.rand_states:                                 # environment.py:546-550
            head = [_draw(p, p.reset_ep_max), _draw(p, p.reset_max_ev), _draw(p, p.reset_max_a)]
        else:                                                  # environment.py:551-555
            head = [p.reset_ep_max, p.reset_max_ev, p.reset_max_a]
        self.x = np.array(head + [a_lead])
        self.prev_x = self.x
        return self.x

    def step(self, u, exog):
        p = self.prm
        x = self.x
        self.u, self.exog = u, exog
        n_ep = abs(x[0]) / p.max_ep                             # environment.py:473-477
        n_ev = abs(x[1]) / p.max_ev
        n_u = abs(u) / abs(p.action_high)
        n_jerk = abs(x[2] - self.prev_x[2]) / (2 * p.action_high)
        self.jerk = (x[2] - self.prev_x[2]) / p.sample_rate
        self.cum_accel += x[2]                                  # environment.py:500-503
        self.velocity = self.cum_accel * p.sample_rate
        self.desired_headway = STANDSTILL_M + p.timegap * self.velocity
        self.headway = x[0] + self.desired_headway
        terminal = bool((abs(x[0]) > p.max_ep or abs(x[1]) > p.max_ev) and p.can_terminate)
        if terminal:                                            # environment.py:505-510
            self.reward = p.terminal_reward * p.re_scalar
        else:
            self.reward = (p.reward_ep_coeff * n_ep + p.reward_ev_coeff * n_ev
                           + p.reward_u_coeff * n_u + p.reward_jerk_coeff * n_jerk) * p.re_scalar
        self.prev_x = x
        self.x = self.A.dot(x) + self.B.dot(u) + self.C.dot(exog)   # environment.py:513
        return self.x, -self.reward, terminal


class SerialPlatoon:
    """One platoon, global-RNG draw order identical to the reference:
    ctor  : N(0,reset_a_leader), N(0,reset_max_u), then per follower 3 state draws
    reset : N(0,reset_a_leader), then per follower [N(0,reset_max_u), 3 state draws]."""

    def __init__(self, length: int, prm: EnvParams, rand_states=True, eval_states=False):
        self.prm, self.length = prm, length
        self.ns = num_states_of(prm)
        self.front_accel = _draw(prm, prm.pl_leader_reset_a)   # environment.py:24
        self.front_u = _draw(prm, prm.reset_max_u)             # environment.py:32
        self.followers = []
        for m in range(length):                                # environment.py:55-63
            tl = prm.pl_leader_tau if m == 0 else self.followers[m - 1].tau
            al = self.front_accel if m == 0 else self.followers[m - 1].x[2]
            self.followers.append(SerialVehicle(prm, m, tl, al, rand_states, eval_states))

    def reset(self):
        p = self.prm
        self.front_accel = _draw(p, p.pl_leader_reset_a)       # environment.py:286
        states = []
        for m, f in enumerate(self.followers):
            self.front_u = _draw(p, p.reset_max_u)             # environment.py:289
            al = self.front_accel if m == 0 else self.followers[m - 1].x[2]
            states.append(f.reset(al)[: self.ns])
        if p.framework == "centralized":
            states = [list(np.concatenate(states).flat)]
        return states

    def exog_for(self, m, leader_exog):                        # environment.py:253-269
        if self.prm.model == "ModelB":
            if m == 0:
                return self.front_u if leader_exog is None else leader_exog
            return self.followers[m - 1].u
        if m == 0:
            return self.front_accel if leader_exog is None else leader_exog
        return self.followers[m - 1].x[2]

    def step(self, actions: Sequence[float], leader_exog=None):
        states, rewards, terms = [], [], []
        for m, u in enumerate(actions):
            x, r, t = self.followers[m].step(u, self.exog_for(m, leader_exog))
            states.append(x[: self.ns])
            rewards.append(r)
            terms.append(t)
        if self.prm.framework == "centralized":                # environment.py:234-236
            states = [list(np.concatenate(states).flat)]
            rewards = [(1 / self.length) * sum(rewards)]
        return states, rewards, any(terms)

    def jerks(self):                                           # environment.py:243-251
        return [[f.jerk] for f in self.followers]


class SerialOUNoise:
    """noise.py:3-29.  One scalar process; state starts at x_init or 0 and is never reset by
    the trainer between episodes."""

    def __init__(self, prm: EnvParams, mean=None, x_init=None):
        self.prm = prm
        self.mean = np.zeros(1) if mean is None else mean
        self.sigma = float(prm.std_dev) * np.ones(1)
        self.x_init = x_init
        self.reset()

    def reset(self):
        self.x_prev = self.x_init if self.x_init is not None else np.zeros_like(self.mean)

    def __call__(self):
        p = self.prm
        z = np.random.normal(0, 1.0, size=self.mean.shape)      # util.py:69-70 via noise.py:18
        x = self.x_prev + p.theta * (self.mean - self.x_prev) * p.ou_dt + self.sigma * np.sqrt(p.ou_dt) * z
        self.x_prev = x
        return x


def clip_action(mu, noise, lo, hi):
    """ddpgagent.policy (agent/ddpgagent.py:18-29) minus the tensor plumbing."""
    a = np.asarray(mu, dtype=np.float64) if noise is None else np.asarray(mu) + noise
    return np.clip(a, lo, hi)


class SerialReplay:
    """replaybuffer.py:5-63 (ring of float64 rows; sample = choice with replacement)."""

    def __init__(self, capacity, batch, ns, na):
        self.capacity, self.batch, self.count = capacity, batch, 0
        self.s = np.zeros((capacity, ns))
        self.a = np.zeros((capacity, na))
        self.r = np.zeros((capacity, 1))
        self.s2 = np.zeros((capacity, ns))

    def add(self, s, a, r, s2):
        i = self.count % self.capacity
        self.s[i], self.a[i], self.r[i], self.s2[i] = s, a, r, s2
        self.count += 1

    def sample_indices(self):
        return np.random.choice(min(self.count, self.capacity), self.batch)   # replaybuffer.py:52-54

    def gather(self, idx):
        return self.s[idx], self.a[idx], self.r[idx].astype(np.float32), self.s2[idx]

    def sample(self):
        return self.gather(self.sample_indices())


# --------------------------------------------------------------------------- batched statement
@dataclass
class BatchedState:
    x: np.ndarray                    # [P, M, 4] float64
    prev_a: np.ndarray               # [P, M]   prev_x[2]
    cum_accel: np.ndarray            # [P, M]
    front_accel: np.ndarray          # [P]
    front_u: np.ndarray              # [P]
    jerk: np.ndarray = field(default=None)
    velocity: np.ndarray = field(default=None)
    headway: np.ndarray = field(default=None)
    desired_headway: np.ndarray = field(default=None)


class BatchedPlatoons:
    """Vectorised float64 statement of Platoon.step / reset with injected randomness."""

    def __init__(self, P: int, M: int, prm: EnvParams):
        self.P, self.M, self.prm = P, M, prm
        mats = follower_matrices(prm, M)
        self.A = np.stack([m[0] for m in mats])      # [M,4,4]
        self.B = np.stack([m[1] for m in mats])      # [M,4]
        self.C = np.stack([m[2] for m in mats])      # [M,4]
        self.ns = num_states_of(prm)
        self.st: Optional[BatchedState] = None

    def set_state(self, x0, front_accel=None, front_u=None):
        """x0[P,M,3 or 4]: (ep, ev, a[, a_lead]).  When only 3 columns are given a_lead is
        chained as in reset (leader accel for m=0, predecessor's fresh x[2] otherwise:
        environment.py:291-294)."""
        P, M = self.P, self.M
        x0 = np.asarray(x0, dtype=np.float64)
        fa = np.zeros(P) if front_accel is None else np.asarray(front_accel, dtype=np.float64)
        fu = np.zeros(P) if front_u is None else np.asarray(front_u, dtype=np.float64)
        x = np.zeros((P, M, 4))
        x[..., : x0.shape[-1]] = x0
        if x0.shape[-1] == 3:
            x[:, 0, 3] = fa
            x[:, 1:, 3] = x[:, :-1, 2]
        self.st = BatchedState(x=x, prev_a=x[..., 2].copy(), cum_accel=np.zeros((P, M)),
                               front_accel=fa, front_u=fu)
        return x[..., : self.ns].copy()

    def step(self, actions, leader_exog=None):
        """actions[P,M]; leader_exog[P] or None -> (obs[P,M,ns], reward[P,M], done[P]).
        Centralized framework: reward[P,1] = mean over followers, obs still [P,M,ns]."""
        p, st = self.prm, self.st
        u = np.asarray(actions, dtype=np.float64).reshape(self.P, self.M)
        x = st.x
        n_ep = np.abs(x[..., 0]) / p.max_ep
        n_ev = np.abs(x[..., 1]) / p.max_ev
        n_u = np.abs(u) / abs(p.action_high)
        n_jerk = np.abs(x[..., 2] - st.prev_a) / (2 * p.action_high)
        st.jerk = (x[..., 2] - st.prev_a) / p.sample_rate
        st.cum_accel = st.cum_accel + x[..., 2]
        st.velocity = st.cum_accel * p.sample_rate
        st.desired_headway = STANDSTILL_M + p.timegap * st.velocity
        st.headway = x[..., 0] + st.desired_headway
        term = ((np.abs(x[..., 0]) > p.max_ep) | (np.abs(x[..., 1]) > p.max_ev)) & bool(p.can_terminate)
        shaped = (p.reward_ep_coeff * n_ep + p.reward_ev_coeff * n_ev + p.reward_u_coeff * n_u
                  + p.reward_jerk_coeff * n_jerk) * p.re_scalar
        reward = np.where(term, p.terminal_reward * p.re_scalar, shaped)
        xn = np.empty_like(x)
        if p.model == "ModelB":
            lead = st.front_u if leader_exog is None else np.asarray(leader_exog, dtype=np.float64)
            exog = np.concatenate([lead.reshape(self.P, 1), u[:, :-1]], axis=1)
            xn = np.einsum("mij,pmj->pmi", self.A, x) + self.B[None] * u[..., None] + self.C[None] * exog[..., None]
        else:  # Model A: predecessor's *post-update* acceleration (environment.py:263-267)
            lead = st.front_accel if leader_exog is None else np.asarray(leader_exog, dtype=np.float64)
            w = lead
            for m in range(self.M):
                xn[:, m] = x[:, m] @ self.A[m].T + np.outer(u[:, m], self.B[m]) + np.outer(w, self.C[m])
                w = xn[:, m, 2]
        st.prev_a = x[..., 2].copy()
        st.x = xn
        rew_out = -reward
        if p.framework == "centralized":
            rew_out = rew_out.mean(axis=1, keepdims=True)
        return xn[..., : self.ns].copy(), rew_out, term.any(axis=1)


def ou_step(x_prev, z, prm: EnvParams, mean=0.0):
    """Vectorised OU update with injected N(0,1) draws z (noise.py:14-23)."""
    return x_prev + prm.theta * (mean - x_prev) * prm.ou_dt + prm.std_dev * np.sqrt(prm.ou_dt) * z
