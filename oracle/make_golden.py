"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE (container only).

TEST INFRASTRUCTURE ONLY.  Usage (from the repo root, in the build container where
/root/reference is mounted):

    python -m oracle.make_golden

The reference has no tests and no golden vectors of its own, so the fixtures are
outputs of the unmodified reference modules imported through oracle/ref_import.py.  Each case
records its inputs (seed, injected actions / exogenous inputs / initial states) and the
reference's outputs in float64.  tests/ compares (a) the CPU restatement in oracle/ and (b) the
CUDA path against these files; nothing at run time on the GPU box touches /root/reference.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _conf(ref, **over):
    c = ref.config.Config()
    for k, v in over.items():
        setattr(c, k, v)
    return c


def _rollout(ref, conf, M, steps, act_seed, act_scale=0.8, init=None, seed=1, leader_none=False):
    """Reference rollout with injected actions / leader exog.  Returns dict of arrays."""
    np.random.seed(seed)
    pl = ref_import.make_platoon(ref, M, conf, 0)
    s0 = pl.reset()
    if init is not None:
        for m, f in enumerate(pl.followers):
            f.x = np.array(init[m], dtype=np.float64)
            f.prev_x = f.x
        s0 = [f.x[: pl.num_states] for f in pl.followers]
    rs = np.random.RandomState(act_seed)
    acts = np.clip(rs.normal(0, act_scale, size=(steps, M)), conf.action_low, conf.action_high)
    exog = rs.normal(0, conf.reset_max_u, size=steps)
    ns = pl.def_num_states
    full0 = np.stack([f.x.copy() for f in pl.followers])
    S = np.zeros((steps, M, ns)); R = np.zeros((steps, M if conf.framework == conf.dcntrl else 1))
    X = np.zeros((steps, M, 4)); D = np.zeros(steps, dtype=bool)
    J = np.zeros((steps, M)); V = np.zeros((steps, M)); H = np.zeros((steps, M)); DH = np.zeros((steps, M))
    for k in range(steps):
        st, rw, dn = pl.step(acts[k], None if leader_none else exog[k])
        S[k] = np.reshape(st, (M, ns))
        R[k] = np.asarray(rw, dtype=np.float64).reshape(-1)
        D[k] = dn
        X[k] = np.stack([f.x for f in pl.followers])
        J[k] = np.reshape(pl.get_jerk(), M)
        V[k] = [f.velocity for f in pl.followers]
        H[k] = [f.headway for f in pl.followers]
        DH[k] = [f.desired_headway for f in pl.followers]
    return dict(x0=full0, reset_obs=np.reshape(s0, (M, -1)), actions=acts, exog=exog, obs=S, reward=R,
                done=D, x=X, jerk=J, velocity=V, headway=H, desired_headway=DH,
                front_u=np.float64(pl.front_u), front_accel=np.float64(pl.front_accel),
                A=np.stack([f.A for f in pl.followers]), B=np.stack([f.B for f in pl.followers]),
                C=np.stack([f.C for f in pl.followers]))


def case_default_trace(ref):
    """Default Config, seed 1, 2 followers: ctor, reset, OU, 60 trainer-ordered steps
    (M OU draws then one leader-exog draw per platoon-step: trainer.py:286-296)."""
    conf = _conf(ref)
    np.random.seed(1)
    pl = ref_import.make_platoon(ref, 2, conf, 0)
    ctor_x = np.stack([f.x.copy() for f in pl.followers])
    reset_obs = np.stack(pl.reset())
    ous = [ref.noise.OUActionNoise(mean=np.zeros(1), config=conf) for _ in range(2)]
    mu_rs = np.random.RandomState(7)
    steps = 60
    mu = mu_rs.normal(0, 0.5, size=(steps, 2))
    acts = np.zeros((steps, 2)); noises = np.zeros((steps, 2)); exog = np.zeros(steps)
    obs = np.zeros((steps, 2, 4)); rew = np.zeros((steps, 2)); done = np.zeros(steps, dtype=bool)
    for k in range(steps):
        for m in range(2):
            nz = ous[m]()
            noises[k, m] = nz[0]
            # ddpgagent.policy needs a tensor-like with .numpy(): use the shim
            import tensorflow as tf  # the NumPy-backed shim from ref_import
            acts[k, m] = ref.ddpgagent.policy(tf.convert_to_tensor(np.float32([[mu[k, m]]])), ous_wrap(nz), conf.action_low, conf.action_high)[0]
        exog[k] = ref.util.get_random_val(conf.rand_gen, conf.reset_max_u, std_dev=conf.reset_max_u, config=conf)
        st, rw, dn = pl.step(acts[k], exog[k])
        obs[k] = np.stack(st); rew[k] = rw; done[k] = dn
    return dict(ctor_x=ctor_x, reset_obs=reset_obs, mu=mu, ou=noises, actions=acts, exog=exog, obs=obs,
                reward=rew, done=done)


class ous_wrap:
    """policy() calls noise_object() once; hand it the sample we already drew so the OU stream
    is advanced exactly once per agent-step like the trainer does."""

    def __init__(self, v):
        self.v = v

    def __call__(self):
        return self.v


def case_terminal(ref):
    conf = _conf(ref)
    np.random.seed(1)
    pl = ref_import.make_platoon(ref, 2, conf, 0)
    pl.reset()
    pl.followers[0].x = np.array([25.0, 0.0, 0.0, 0.0]); pl.followers[0].prev_x = pl.followers[0].x
    x1 = pl.followers[1].x.copy()
    st, rw, dn = pl.step([0.0, 0.0], 0.0)
    return dict(x_before=np.stack([np.array([25.0, 0, 0, 0]), x1]), obs=np.stack(st), reward=np.array(rw), done=np.array(dn))


def case_multi_platoon(ref, P=48, M=4, steps=200):
    """P independent reference platoons with injected initial states and actions."""
    conf = _conf(ref, can_terminate=True)
    rs = np.random.RandomState(2024)
    x0 = np.zeros((P, M, 3))
    x0[..., 0] = rs.normal(0, 6.0, (P, M)); x0[..., 1] = rs.normal(0, 6.0, (P, M)); x0[..., 2] = rs.normal(0, 0.5, (P, M))
    fa = rs.normal(0, 0.2, P)
    acts = np.clip(rs.normal(0, 1.5, (steps, P, M)), -2.5, 2.5)
    exog = rs.normal(0, 0.1, (steps, P))
    obs = np.zeros((steps, P, M, 4)); rew = np.zeros((steps, P, M)); done = np.zeros((steps, P), dtype=bool)
    np.random.seed(3)
    for p in range(P):
        pl = ref_import.make_platoon(ref, M, conf, p)
        for m, f in enumerate(pl.followers):
            al = fa[p] if m == 0 else x0[p, m - 1, 2]
            f.x = np.array([x0[p, m, 0], x0[p, m, 1], x0[p, m, 2], al]); f.prev_x = f.x
        for k in range(steps):
            st, rw, dn = pl.step(acts[k, p], exog[k, p])
            obs[k, p] = np.stack(st); rew[k, p] = rw; done[k, p] = dn
    return dict(x0=x0, front_accel=fa, actions=acts, exog=exog, obs=obs, reward=rew, done=done)


def case_resets(ref):
    out = {}
    for tag, over, kw in [("normal", {}, {}), ("uniform", {"rand_gen": "uniform"}, {}),
                          ("fixed", {}, {"rand_states": False}),
                          ("eval", {}, {"evaluator_states_enabled": True})]:
        conf = _conf(ref, **over)
        np.random.seed(11)
        pl = ref_import.make_platoon(ref, 3, conf, 0, **kw)
        a = np.stack(pl.reset()); b = np.stack(pl.reset())
        out[f"{tag}_first"] = a; out[f"{tag}_second"] = b
        out[f"{tag}_front_u"] = np.float64(pl.front_u); out[f"{tag}_front_accel"] = np.float64(pl.front_accel)
    return out


def case_ou(ref):
    conf = _conf(ref)
    np.random.seed(5)
    ou = ref.noise.OUActionNoise(mean=np.zeros(1), config=conf)
    xs = np.array([ou()[0] for _ in range(1000)])
    np.random.seed(5)
    z = np.array([np.random.normal(0, 1.0, size=(1,))[0] for _ in range(1000)])
    return dict(samples=xs, z=z)


def case_replay(ref):
    np.random.seed(9)
    rb = ref.replaybuffer.ReplayBuffer(128, 16, 4, 1, 2)
    rs = np.random.RandomState(1)
    S = rs.normal(size=(300, 4)); A = rs.normal(size=(300, 1)); R = rs.normal(size=(300,)); S2 = rs.normal(size=(300, 4))
    idx_log, batches = [], []
    for i in range(300):
        rb.add((S[i], A[i], R[i], S2[i]))
        if i in (10, 127, 128, 299):
            np.random.seed(100 + i)
            rng_state_idx = np.random.choice(min(rb.buffer_counter, rb.buffer_capacity), rb.batch_size)
            np.random.seed(100 + i)
            s, a, r, s2 = rb.sample()
            idx_log.append(rng_state_idx)
            batches.append(np.concatenate([np.asarray(s), np.asarray(a), np.asarray(r, dtype=np.float64), np.asarray(s2)], axis=1))
    np.random.seed(1)
    choice_kat = np.random.choice(100, 8)
    return dict(S=S, A=A, R=R, S2=S2, at=np.array([10, 127, 128, 299]), idx=np.stack(idx_log), batch=np.stack(batches),
                ring_s=rb.state_buffer, ring_a=rb.action_buffer, ring_r=rb.reward_buffer, ring_s2=rb.next_state_buffer,
                counter=np.int64(rb.buffer_counter), choice_kat=choice_kat)


def case_polyak(ref):
    rs = np.random.RandomState(4)
    shapes = [(4, 256), (256,), (256,), (256,), (256,), (256,), (256, 128), (128,), (128, 1), (1,)]
    cw = [rs.normal(size=s).astype(np.float32) for s in shapes]
    tcw = [rs.normal(size=s).astype(np.float32) for s in shapes]
    aw = [rs.normal(size=s).astype(np.float32) for s in shapes[:6]]
    taw = [rs.normal(size=s).astype(np.float32) for s in shapes[:6]]
    tc_new, ta_new = ref.ddpgagent.update_target(0.001, tcw, cw, taw, aw)
    d = {}
    for i, (a, b, c) in enumerate(zip(cw, tcw, tc_new)):
        d[f"c{i}"] = a; d[f"tc{i}"] = b; d[f"tc_new{i}"] = np.asarray(c)
    for i, (a, b, c) in enumerate(zip(aw, taw, ta_new)):
        d[f"a{i}"] = a; d[f"ta{i}"] = b; d[f"ta_new{i}"] = np.asarray(c)
    d["tau"] = np.float64(0.001); d["n_c"] = np.int64(len(cw)); d["n_a"] = np.int64(len(aw))
    return d


def case_fedavg(ref):
    """The reference's own demo inputs (src/server/test_federated.py:26-42, interfrl, weighted) run
    through the reference Server, plus a larger random case."""
    srv = ref.federated.Server("golden", False)
    f32 = np.float32
    pl = [[[f32([1, 2, 3]), f32([1, 2]), f32([3, 4])], [f32([7, 8, 9]), f32([5, 6]), f32([7, 8])]],
          [[f32([10, 11, 12]), f32([9, 10]), f32([11, 12])], [f32([13, 14, 15]), f32([13, 14]), f32([15, 16])]]]
    w = [[2, 0.5], [1, 6]]
    P, M = 2, 2
    sys_params = [[None] * P for _ in range(M)]; fw = [[0.0] * P for _ in range(M)]
    raw = [[None] * P for _ in range(M)]
    for p in range(P):
        for m in range(M):
            arr = np.array(pl[p][m], dtype=object)
            sys_params[m][p] = w[p][m] * arr; fw[m][p] = w[p][m]; raw[m][p] = arr
    sums = np.sum(np.array(fw), axis=1)
    wavg = srv.get_weighted_avg_params(sys_params, sums)
    avg = srv.get_avg_params(raw)
    d = {"kat_weights": np.array(w, dtype=np.float64), "kat_sums": sums}
    for m in range(M):
        for l in range(3):
            d[f"kat_wavg_{m}_{l}"] = np.asarray(wavg[m][l]); d[f"kat_avg_{m}_{l}"] = np.asarray(avg[m][l])
            for p in range(P):
                d[f"kat_in_{p}_{m}_{l}"] = pl[p][m][l]
    # larger random case: S systems x X members x L layers
    rs = np.random.RandomState(8)
    S, X = 3, 5
    shapes = [(4, 16), (16,), (16, 8), (8,), (8, 1)]
    members = [[[rs.normal(size=s).astype(np.float32) for s in shapes] for _ in range(X)] for _ in range(S)]
    wts = rs.uniform(0.1, 3.0, size=(S, X))
    weighted = [[np.multiply(np.array(members[s][x], dtype=object), wts[s][x]) for x in range(X)] for s in range(S)]
    plain = [[np.array(members[s][x], dtype=object) for x in range(X)] for s in range(S)]
    wavg = srv.get_weighted_avg_params(weighted, np.sum(wts, axis=1))
    avg = srv.get_avg_params(plain)
    d["rnd_weights"] = wts
    for s in range(S):
        for l in range(len(shapes)):
            d[f"rnd_wavg_{s}_{l}"] = np.asarray(wavg[s][l]); d[f"rnd_avg_{s}_{l}"] = np.asarray(avg[s][l])
            for x in range(X):
                d[f"rnd_in_{s}_{x}_{l}"] = members[s][x][l]
    d["rnd_shape"] = np.array([S, X, len(shapes)])
    return d


def main():
    assert ref_import.reference_available(), "run in the build container (needs /root/reference)"
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "default_trace": case_default_trace(ref),
        "terminal": case_terminal(ref),
        "rollout_euler_M4": _rollout(ref, _conf(ref, can_terminate=False), 4, 1000, 21),
        "rollout_exact_M4": _rollout(ref, _conf(ref, can_terminate=False, method="exact"), 4, 1000, 22),
        "rollout_exact_hetero_M3": _rollout(ref, _conf(ref, can_terminate=False, method="exact", pl_leader_tau=0.25,
                                                        timegap=1.3, dyn_coeff=0.15), 3, 300, 23),
        "rollout_terminating_M4": _rollout(ref, _conf(ref, max_ep=4.0, max_ev=4.0), 4, 300, 24, act_scale=2.0),
        "rollout_modelA_M3": _rollout(ref, _conf(ref, model="ModelA", can_terminate=False), 3, 300, 25),
        "rollout_M8": _rollout(ref, _conf(ref, can_terminate=False), 8, 200, 26),
        "rollout_central_M3": _rollout(ref, _conf(ref, framework="centralized", can_terminate=False), 3, 100, 27),
        "rollout_leader_none_M2": _rollout(ref, _conf(ref, can_terminate=False), 2, 50, 28, leader_none=True),
        "multi_platoon": case_multi_platoon(ref),
        "resets": case_resets(ref),
        "ou": case_ou(ref),
        "replay": case_replay(ref),
        "polyak": case_polyak(ref),
        "fedavg": case_fedavg(ref),
    }
    manifest = {}
    for name, arrs in cases.items():
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **arrs)
        manifest[name] = {k: list(np.shape(v)) for k, v in arrs.items()}
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "python -m oracle.make_golden", "reference": "cboin1996/avddpg @ /root/reference",
                   "numpy": np.__version__, "cases": manifest}, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()


# The evaluator fixture (tests/golden/evaluator.npz) is produced by tools/make_evaluator_golden.py: the reference's
# Platoon(evaluator_states_enabled=True) + its own pre-drawn leader inputs + ddpgagent.policy, driven by the oracle's
# actor forward (the reference's Keras actors need TensorFlow).
