"""CPU restatement of the DDPG learn step, TF-Keras Adam, Polyak update and FedAvg.
TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against real TensorFlow: the reference computes these with tensorflow==2.4.1
(requirements.txt:2), which cannot be installed here (no network) and is not vendored, and the
reference has no tests or fixtures at this boundary.  What follows restates the
reference's call sites with the published TF-Keras semantics, and tests/test_oracle_ddpg.py checks
the hand-written backward pass against an independent torch-autograd statement of the same networks.
Polyak and FedAvg ARE pinned: the reference's own code for them runs under the NumPy-backed shim
(oracle/ref_import.py) and its outputs are in tests/golden/{polyak,fedavg}.npz.

Follows:
  actor network   /root/reference/agent/model.py:4-38    Dense(256,relu)->BN->Dense(128,relu)->BN->Dense(1,tanh)*high
  critic network  /root/reference/agent/model.py:41-85   (s->Dense(256,relu)->BN || a->Dense(48,relu)->BN)->concat
                                                          ->Dense(128,relu)->BN->Dense(1)
  learn           /root/reference/workers/trainer.py:489-506
  Adam            tf.keras.optimizers.Adam as constructed at workers/trainer.py:138-139 (beta1=.9, beta2=.999,
                  epsilon=1e-7, non-amsgrad):  lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)
  Polyak          /root/reference/agent/ddpgagent.py:31-55
  FedAvg          /root/reference/src/server/federated.py:18-122 and its use at workers/trainer.py:400-456

TF-Keras semantics that matter: the models are always called without training=True, so
BatchNormalization is the inference affine  y = gamma*(x-moving_mean)/sqrt(moving_var+1e-3)+beta  with
trainable gamma/beta and frozen moving statistics; kernel_regularizer='l2' never contributes (model.losses is
never added); there is no terminal mask in the TD target; both gradients use the pre-update weights.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3          # tf.keras.layers.BatchNormalization default epsilon
F32 = np.float32

ACTOR_WEIGHTS = ["W1", "b1", "g1", "be1", "mu1", "var1", "W2", "b2", "g2", "be2", "mu2", "var2", "W3", "b3"]
ACTOR_TRAINABLE = ["W1", "b1", "g1", "be1", "W2", "b2", "g2", "be2", "W3", "b3"]
CRITIC_WEIGHTS = ["Ws", "bs", "Wa", "ba", "gs", "bes", "mus", "vars", "ga", "bea", "mua", "vara",
                  "W2", "b2", "g2", "be2", "mu2", "var2", "W3", "b3"]
CRITIC_TRAINABLE = ["Ws", "bs", "Wa", "ba", "gs", "bes", "ga", "bea", "W2", "b2", "g2", "be2", "W3", "b3"]


def actor_shapes(ns=4, l1=256, l2=128, na=1):
    return {"W1": (ns, l1), "b1": (l1,), "g1": (l1,), "be1": (l1,), "mu1": (l1,), "var1": (l1,),
            "W2": (l1, l2), "b2": (l2,), "g2": (l2,), "be2": (l2,), "mu2": (l2,), "var2": (l2,),
            "W3": (l2, na), "b3": (na,)}


def critic_shapes(ns=4, l1=256, la=48, l2=128, na=1):
    return {"Ws": (ns, l1), "bs": (l1,), "Wa": (na, la), "ba": (la,),
            "gs": (l1,), "bes": (l1,), "mus": (l1,), "vars": (l1,),
            "ga": (la,), "bea": (la,), "mua": (la,), "vara": (la,),
            "W2": (l1 + la, l2), "b2": (l2,), "g2": (l2,), "be2": (l2,), "mu2": (l2,), "var2": (l2,),
            "W3": (l2, 1), "b3": (1,)}


def init_actor(rng: np.random.Generator, ns=4, l1=256, l2=128):
    """model.py:19-25: U(+-1/sqrt(l1)), U(+-1/sqrt(l2)), U(+-3e-3); zero biases; BN gamma=1, beta=0, mean=0, var=1."""
    sh = actor_shapes(ns, l1, l2)
    p = {k: np.zeros(v, F32) for k, v in sh.items()}
    p["W1"] = rng.uniform(-1 / np.sqrt(l1), 1 / np.sqrt(l1), sh["W1"]).astype(F32)
    p["W2"] = rng.uniform(-1 / np.sqrt(l2), 1 / np.sqrt(l2), sh["W2"]).astype(F32)
    p["W3"] = rng.uniform(-3e-3, 3e-3, sh["W3"]).astype(F32)
    for g, v in (("g1", "var1"), ("g2", "var2")):
        p[g][:] = 1
        p[v][:] = 1
    return p


def init_critic(rng: np.random.Generator, ns=4, l1=256, la=48, l2=128):
    """model.py:53-60: state layer U(+-1/sqrt(l1)); action layer AND layer 2 use the layer-2 bound (model.py:70,76);
    head U(+-3e-4)."""
    sh = critic_shapes(ns, l1, la, l2)
    p = {k: np.zeros(v, F32) for k, v in sh.items()}
    p["Ws"] = rng.uniform(-1 / np.sqrt(l1), 1 / np.sqrt(l1), sh["Ws"]).astype(F32)
    p["Wa"] = rng.uniform(-1 / np.sqrt(l2), 1 / np.sqrt(l2), sh["Wa"]).astype(F32)
    p["W2"] = rng.uniform(-1 / np.sqrt(l2), 1 / np.sqrt(l2), sh["W2"]).astype(F32)
    p["W3"] = rng.uniform(-3e-4, 3e-4, sh["W3"]).astype(F32)
    for g, v in (("gs", "vars"), ("ga", "vara"), ("g2", "var2")):
        p[g][:] = 1
        p[v][:] = 1
    return p


def randomize_bn(p, rng, names):
    """Give BN parameters non-trivial values so tests exercise the general affine (not just gamma=1, var=1)."""
    for g, b, mu, var in names:
        p[g] = rng.uniform(0.5, 1.5, p[g].shape).astype(F32)
        p[b] = rng.uniform(-0.2, 0.2, p[b].shape).astype(F32)
        p[mu] = rng.uniform(-0.1, 0.1, p[mu].shape).astype(F32)
        p[var] = rng.uniform(0.5, 1.5, p[var].shape).astype(F32)
    return p


# ----------------------------------------------------------------------------------- forward / backward
def _bn(x, g, b, mu, var):
    inv = (F32(1) / np.sqrt(var + F32(BN_EPS))).astype(F32)
    xh = (x - mu) * inv
    return g * xh + b, xh, inv


def actor_forward(p, s, high=2.5):
    s = np.asarray(s, F32)
    z1 = s @ p["W1"] + p["b1"]; r1 = np.maximum(z1, 0)
    h1, xh1, inv1 = _bn(r1, p["g1"], p["be1"], p["mu1"], p["var1"])
    z2 = h1 @ p["W2"] + p["b2"]; r2 = np.maximum(z2, 0)
    h2, xh2, inv2 = _bn(r2, p["g2"], p["be2"], p["mu2"], p["var2"])
    t = np.tanh(h2 @ p["W3"] + p["b3"])
    out = (t * F32(high)).astype(F32)
    return out, dict(s=s, z1=z1, xh1=xh1, inv1=inv1, h1=h1, z2=z2, xh2=xh2, inv2=inv2, h2=h2, t=t, high=F32(high))


def _ident(x):
    return x


def actor_backward(p, c, dout, absolute=False):
    """dout[B,1] = dL/d(actor output).  Returns grads in ACTOR_TRAINABLE order (dict).
    absolute=True: the same sums over the batch with every row's contribution replaced by its magnitude, sum_n |g_n| -- the
    scale against which a rounding error of a gradient tensor is judged when the signed sum nearly cancels (tests only)."""
    A = np.abs if absolute else _ident
    g = {}
    dpre3 = dout * c["high"] * (F32(1) - c["t"] ** 2)
    g["W3"] = A(c["h2"]).T @ A(dpre3); g["b3"] = A(dpre3).sum(0)
    dh2 = dpre3 @ p["W3"].T
    g["g2"] = A(dh2 * c["xh2"]).sum(0); g["be2"] = A(dh2).sum(0)
    dz2 = dh2 * p["g2"] * c["inv2"] * (c["z2"] > 0)
    g["W2"] = A(c["h1"]).T @ A(dz2); g["b2"] = A(dz2).sum(0)
    dh1 = dz2 @ p["W2"].T
    g["g1"] = A(dh1 * c["xh1"]).sum(0); g["be1"] = A(dh1).sum(0)
    dz1 = dh1 * p["g1"] * c["inv1"] * (c["z1"] > 0)
    g["W1"] = A(c["s"]).T @ A(dz1); g["b1"] = A(dz1).sum(0)
    return {k: v.astype(F32) for k, v in g.items()}


def critic_forward(p, s, a):
    s = np.asarray(s, F32); a = np.asarray(a, F32).reshape(len(s), -1)
    zs = s @ p["Ws"] + p["bs"]; rs = np.maximum(zs, 0)
    hs, xhs, invs = _bn(rs, p["gs"], p["bes"], p["mus"], p["vars"])
    za = a @ p["Wa"] + p["ba"]; ra = np.maximum(za, 0)
    ha, xha, inva = _bn(ra, p["ga"], p["bea"], p["mua"], p["vara"])
    cat = np.concatenate([hs, ha], axis=1)
    z2 = cat @ p["W2"] + p["b2"]; r2 = np.maximum(z2, 0)
    h2, xh2, inv2 = _bn(r2, p["g2"], p["be2"], p["mu2"], p["var2"])
    q = (h2 @ p["W3"] + p["b3"]).astype(F32)
    return q, dict(s=s, a=a, zs=zs, xhs=xhs, invs=invs, za=za, xha=xha, inva=inva, cat=cat, z2=z2, xh2=xh2, inv2=inv2, h2=h2)


def critic_backward(p, c, dq, want_params=True, absolute=False):
    """dq[B,1] = dL/dq.  Returns (param grads dict or None, dL/da[B,1]).  absolute: see actor_backward."""
    A = np.abs if absolute else _ident
    g = {}
    dh2 = dq @ p["W3"].T
    dz2 = dh2 * p["g2"] * c["inv2"] * (c["z2"] > 0)
    dcat = dz2 @ p["W2"].T
    l1 = c["zs"].shape[1]
    dhs, dha = dcat[:, :l1], dcat[:, l1:]
    dzs = dhs * p["gs"] * c["invs"] * (c["zs"] > 0)
    dza = dha * p["ga"] * c["inva"] * (c["za"] > 0)
    da = dza @ p["Wa"].T
    if want_params:
        g["W3"] = A(c["h2"]).T @ A(dq); g["b3"] = A(dq).sum(0)
        g["g2"] = A(dh2 * c["xh2"]).sum(0); g["be2"] = A(dh2).sum(0)
        g["W2"] = A(c["cat"]).T @ A(dz2); g["b2"] = A(dz2).sum(0)
        g["gs"] = A(dhs * c["xhs"]).sum(0); g["bes"] = A(dhs).sum(0)
        g["ga"] = A(dha * c["xha"]).sum(0); g["bea"] = A(dha).sum(0)
        g["Ws"] = A(c["s"]).T @ A(dzs); g["bs"] = A(dzs).sum(0)
        g["Wa"] = A(c["a"]).T @ A(dza); g["ba"] = A(dza).sum(0)
        g = {k: v.astype(F32) for k, v in g.items()}
    return (g if want_params else None), da.astype(F32)


def learn(actor, critic, t_actor, t_critic, batch, gamma=0.99, high=2.5, with_abs=False):
    """Trainer.learn (trainer.py:489-506).  batch = (s[B,ns], a[B,1], r[B,1], s2[B,ns]).
    Returns (critic_grads, actor_grads, info) with grads as dicts keyed like *_TRAINABLE."""
    s, a, r, s2 = (np.asarray(x, F32) for x in batch)
    B = F32(len(s))
    a2, _ = actor_forward(t_actor, s2, high)                               # trainer.py:493
    q2, _ = critic_forward(t_critic, s2, a2)
    y = r.reshape(-1, 1) + F32(gamma) * q2                                   # trainer.py:494
    q, cc = critic_forward(critic, s, a)                                    # trainer.py:495
    critic_loss = np.mean((y - q) ** 2)                                     # trainer.py:496
    dq = (F32(2) * (q - y) / B).astype(F32)
    cg, _ = critic_backward(critic, cc, dq)                                 # trainer.py:498
    pi, ca = actor_forward(actor, s, high)                                  # trainer.py:502
    qpi, cc2 = critic_forward(critic, s, pi)                                # trainer.py:503
    actor_loss = -np.mean(qpi)                                              # trainer.py:504
    _, dpi = critic_backward(critic, cc2, np.full_like(qpi, -1.0 / B), want_params=False)
    ag = actor_backward(actor, ca, dpi)                                     # trainer.py:506
    info = dict(critic_loss=float(critic_loss), actor_loss=float(actor_loss), y=y, q=q, pi=pi)
    if with_abs:      # sum over the batch of the magnitudes of the per-row contributions, per tensor
        info["critic_abs"], _ = critic_backward(critic, cc, dq, absolute=True)
        info["actor_abs"] = actor_backward(actor, ca, dpi, absolute=True)
    return cg, ag, info


# ----------------------------------------------------------------------------------- optimiser / targets
def adam_apply(params, grads, m, v, t, lr, names, b1=0.9, b2=0.999, eps=1e-7):
    """One tf.keras Adam step, in place on params/m/v (dicts).  t is the 1-based step AFTER increment."""
    lr_t = F32(lr) * np.sqrt(F32(1) - F32(b2) ** F32(t)) / (F32(1) - F32(b1) ** F32(t))
    for k in names:
        g = grads[k].astype(F32)
        m[k] = (m[k] + (g - m[k]) * F32(1 - b1)).astype(F32)
        v[k] = (v[k] + (g * g - v[k]) * F32(1 - b2)).astype(F32)
        params[k] = (params[k] - lr_t * m[k] / (np.sqrt(v[k]) + F32(eps))).astype(F32)
    return params


def polyak(target, online, tau, names):
    """ddpgagent.update_target: theta' <- tau*theta + (1-tau)*theta' over ALL weights (incl. BN stats)."""
    return {k: (online[k] * F32(tau) + target[k] * F32(1 - tau)).astype(F32) for k in names}


# ----------------------------------------------------------------------------------- FedAvg
def fed_average(system_params):
    """Server.get_avg_params (federated.py:47-63): per system, per layer, mean over members.
    system_params[system][member] = list of layer arrays."""
    out = []
    for members in system_params:
        out.append([np.mean(np.stack([mem[l] for mem in members], 0), axis=0) for l in range(len(members[0]))])
    return out


def fed_weighted_average(weighted_params, weight_sums):
    """Server.get_weighted_avg_params (federated.py:99-118): members arrive pre-multiplied by their weight;
    result = float32(1/sum_w) * sum."""
    out = []
    for members, wsum in zip(weighted_params, weight_sums):
        scale = F32(1 / wsum)
        out.append([scale * np.sum(np.stack([mem[l] for mem in members], 0), axis=0) for l in range(len(members[0]))])
    return out


def frl_weight(ep_rewards, window):
    """Trainer.get_weight (trainer.py:385-395): |1/mean(last `window` episodic rewards)|."""
    return abs(1 / np.mean(ep_rewards[-window:]))
