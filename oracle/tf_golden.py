"""Reader / checker for tests/golden/tf_learn.npz, the TensorFlow-pinned fixture of the learn step that tools/make_tf_golden.py
writes where TensorFlow 2.4.1 exists.  TEST INFRASTRUCTURE ONLY.

The file stores Keras-ordered lists (`.weights`, `trainable_variables`); this module maps them onto the named tensors of
oracle/ddpg_np.py (ACTOR_WEIGHTS / CRITIC_WEIGHTS are exactly the Keras orders of agent/model.py's functional models, SURVEY.md
section 3.3) after checking every shape, replays the recorded steps with the NumPy oracle and reports the differences.  Until a
maintainer has generated the file the oracle's forward / backward / Adam stay "parity unpinned" (DESIGN.md section 5); the schema and
this consumer are exercised by a synthetic file written from the oracle itself (write_synthetic), which pins nothing.
"""
from __future__ import annotations

import numpy as np

from . import ddpg_np as D

NETS = (("actor", D.ACTOR_WEIGHTS, D.ACTOR_TRAINABLE, D.actor_shapes), ("critic", D.CRITIC_WEIGHTS, D.CRITIC_TRAINABLE, D.critic_shapes),
        ("t_actor", D.ACTOR_WEIGHTS, D.ACTOR_TRAINABLE, D.actor_shapes), ("t_critic", D.CRITIC_WEIGHTS, D.CRITIC_TRAINABLE, D.critic_shapes))


def _named(z, prefix, names, shapes):
    out = {}
    for i, n in enumerate(names):
        arr = np.asarray(z[f"{prefix}_{i:02d}"], dtype=np.float32)
        want = shapes[n]
        if tuple(arr.shape) != tuple(want):
            raise ValueError(f"{prefix}_{i:02d}: shape {arr.shape} does not match {n} {want}: the Keras order differs from oracle/ddpg_np.py")
        out[n] = arr
    return out


def load(path):
    """-> dict(init={net: named weights}, batch=(s, a, r, s2), steps=[dict(critic_grad, actor_grad, losses, weights={net: named})], hyper)"""
    with np.load(path, allow_pickle=False) as f:
        z = {k: f[k] for k in f.files}
    ash, csh = D.actor_shapes(), D.critic_shapes()
    shapes = {"actor": ash, "critic": csh, "t_actor": ash, "t_critic": csh}
    g = {"init": {tag: _named(z, f"init_{tag}", wn, shapes[tag]) for tag, wn, _, _ in NETS},
         "batch": (z["batch_s"], z["batch_a"], z["batch_r"], z["batch_s2"]),
         "hyper": {k: float(z[k]) for k in ("gamma", "tau", "actor_lr", "critic_lr", "high")},
         "tf_version": str(z["tf_version"]), "steps": []}
    for k in range(int(z["steps"])):
        g["steps"].append({"critic_grad": _named(z, f"step{k}_critic_grad", D.CRITIC_TRAINABLE, csh),
                           "actor_grad": _named(z, f"step{k}_actor_grad", D.ACTOR_TRAINABLE, ash),
                           "critic_loss": float(z[f"step{k}_critic_loss"]), "actor_loss": float(z[f"step{k}_actor_loss"]),
                           "weights": {tag: _named(z, f"step{k}_{tag}", wn, shapes[tag]) for tag, wn, _, _ in NETS}})
    return g


def _rel(got, want):
    got, want = np.asarray(got, np.float64).ravel(), np.asarray(want, np.float64).ravel()
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


def replay_with_oracle(g):
    """Runs the recorded steps with oracle/ddpg_np.py from the recorded initial weights.  -> worst rel-L2 differences
    dict(grad=..., loss=..., weights=...) against what TensorFlow produced."""
    nets = {tag: {k: v.copy() for k, v in g["init"][tag].items()} for tag in g["init"]}
    h = g["hyper"]
    z = lambda p, names: {k: np.zeros_like(p[k]) for k in names}
    am, av = z(nets["actor"], D.ACTOR_TRAINABLE), z(nets["actor"], D.ACTOR_TRAINABLE)
    cm, cv = z(nets["critic"], D.CRITIC_TRAINABLE), z(nets["critic"], D.CRITIC_TRAINABLE)
    worst = {"grad": 0.0, "loss": 0.0, "weights": 0.0}
    for k, st in enumerate(g["steps"], start=1):
        cg, ag, info = D.learn(nets["actor"], nets["critic"], nets["t_actor"], nets["t_critic"], g["batch"], gamma=h["gamma"], high=h["high"])
        for n in D.CRITIC_TRAINABLE:
            worst["grad"] = max(worst["grad"], _rel(cg[n].reshape(st["critic_grad"][n].shape), st["critic_grad"][n]))
        for n in D.ACTOR_TRAINABLE:
            worst["grad"] = max(worst["grad"], _rel(ag[n].reshape(st["actor_grad"][n].shape), st["actor_grad"][n]))
        worst["loss"] = max(worst["loss"], abs(info["critic_loss"] - st["critic_loss"]) / max(abs(st["critic_loss"]), 1e-12),
                            abs(info["actor_loss"] - st["actor_loss"]) / max(abs(st["actor_loss"]), 1e-12))
        D.adam_apply(nets["critic"], cg, cm, cv, k, h["critic_lr"], D.CRITIC_TRAINABLE)
        D.adam_apply(nets["actor"], ag, am, av, k, h["actor_lr"], D.ACTOR_TRAINABLE)
        nets["t_critic"] = D.polyak(nets["t_critic"], nets["critic"], h["tau"], D.CRITIC_WEIGHTS)
        nets["t_actor"] = D.polyak(nets["t_actor"], nets["actor"], h["tau"], D.ACTOR_WEIGHTS)
        for tag, wn, _, _ in NETS:
            for n in wn:
                worst["weights"] = max(worst["weights"], _rel(nets[tag][n], st["weights"][tag][n]))
    return worst


def write_synthetic(path, seed=3, batch=64, steps=3):
    """A file with the schema of tools/make_tf_golden.py whose numbers come from the ORACLE (tf_version = 'oracle-synthetic').
    It exists to exercise load() / the consuming tests; it pins nothing."""
    rng = np.random.default_rng(seed)
    nets = {"actor": D.init_actor(rng), "critic": D.init_critic(rng)}
    D.randomize_bn(nets["actor"], rng, [("g1", "be1", "mu1", "var1"), ("g2", "be2", "mu2", "var2")])
    D.randomize_bn(nets["critic"], rng, [("gs", "bes", "mus", "vars"), ("ga", "bea", "mua", "vara"), ("g2", "be2", "mu2", "var2")])
    nets["actor"]["W3"] *= 50
    nets["critic"]["W3"] *= 300
    nets["t_actor"] = {k: (v + rng.normal(0, 0.01, v.shape)).astype(np.float32) for k, v in nets["actor"].items()}
    nets["t_critic"] = {k: (v + rng.normal(0, 0.01, v.shape)).astype(np.float32) for k, v in nets["critic"].items()}
    s = rng.normal(0, 2, (batch, 4)).astype(np.float32)
    b = (s, rng.uniform(-2.5, 2.5, (batch, 1)).astype(np.float32), -rng.uniform(0, 0.5, (batch, 1)).astype(np.float32),
         (s + rng.normal(0, 0.2, (batch, 4))).astype(np.float32))
    h = dict(gamma=0.99, tau=0.001, actor_lr=5e-5, critic_lr=5e-4, high=2.5)
    out = {"batch_s": b[0], "batch_a": b[1], "batch_r": b[2], "batch_s2": b[3], "tf_version": np.array("oracle-synthetic"), "steps": np.int64(steps)}
    out.update({k: np.float64(v) for k, v in h.items()})
    for tag, wn, tn, _ in NETS:
        out[f"{tag}_weight_names"], out[f"{tag}_trainable_names"] = np.array(wn), np.array(tn)
        for i, n in enumerate(wn):
            out[f"init_{tag}_{i:02d}"] = nets[tag][n]
    z = lambda p, names: {k: np.zeros_like(p[k]) for k in names}
    am, av = z(nets["actor"], D.ACTOR_TRAINABLE), z(nets["actor"], D.ACTOR_TRAINABLE)
    cm, cv = z(nets["critic"], D.CRITIC_TRAINABLE), z(nets["critic"], D.CRITIC_TRAINABLE)
    for k in range(steps):
        cg, ag, info = D.learn(nets["actor"], nets["critic"], nets["t_actor"], nets["t_critic"], b, gamma=h["gamma"], high=h["high"])
        out[f"step{k}_critic_loss"], out[f"step{k}_actor_loss"] = np.float64(info["critic_loss"]), np.float64(info["actor_loss"])
        for i, n in enumerate(D.CRITIC_TRAINABLE):
            out[f"step{k}_critic_grad_{i:02d}"] = cg[n].reshape(D.critic_shapes()[n])
        for i, n in enumerate(D.ACTOR_TRAINABLE):
            out[f"step{k}_actor_grad_{i:02d}"] = ag[n].reshape(D.actor_shapes()[n])
        D.adam_apply(nets["critic"], cg, cm, cv, k + 1, h["critic_lr"], D.CRITIC_TRAINABLE)
        D.adam_apply(nets["actor"], ag, am, av, k + 1, h["actor_lr"], D.ACTOR_TRAINABLE)
        nets["t_critic"] = D.polyak(nets["t_critic"], nets["critic"], h["tau"], D.CRITIC_WEIGHTS)
        nets["t_actor"] = D.polyak(nets["t_actor"], nets["actor"], h["tau"], D.ACTOR_WEIGHTS)
        for tag, wn, _, _ in NETS:
            for i, n in enumerate(wn):
                out[f"step{k}_{tag}_{i:02d}"] = nets[tag][n]
    np.savez_compressed(path, **out)
    return path
