"""Checks of the DDPG/Adam/Polyak/FedAvg restatement (oracle/ddpg_np.py).  CPU only.

 * hand-written backward pass vs an independent torch-autograd statement of the same networks
   (TensorFlow is not installable here, so this -- not TF -- is what the learn step is checked against:
   "parity unpinned" vs real TF, see the oracle header);
 * Adam vs the closed form after one step and vs torch's formula where the two coincide;
 * Polyak and FedAvg vs golden vectors produced by the REFERENCE's own code.
"""
import numpy as np
import pytest
import torch

from oracle import ddpg_np as D


def _t(p):
    return {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}


def _bn_t(x, g, b, mu, var):
    return g * (x - mu) / torch.sqrt(var + D.BN_EPS) + b


def torch_actor(p, s, high=2.5):
    h1 = _bn_t(torch.relu(s @ p["W1"] + p["b1"]), p["g1"], p["be1"], p["mu1"], p["var1"])
    h2 = _bn_t(torch.relu(h1 @ p["W2"] + p["b2"]), p["g2"], p["be2"], p["mu2"], p["var2"])
    return torch.tanh(h2 @ p["W3"] + p["b3"]) * high


def torch_critic(p, s, a):
    hs = _bn_t(torch.relu(s @ p["Ws"] + p["bs"]), p["gs"], p["bes"], p["mus"], p["vars"])
    ha = _bn_t(torch.relu(a @ p["Wa"] + p["ba"]), p["ga"], p["bea"], p["mua"], p["vara"])
    h2 = _bn_t(torch.relu(torch.cat([hs, ha], 1) @ p["W2"] + p["b2"]), p["g2"], p["be2"], p["mu2"], p["var2"])
    return h2 @ p["W3"] + p["b3"]


def make_nets(seed, random_bn=True):
    rng = np.random.default_rng(seed)
    nets = [D.init_actor(rng), D.init_critic(rng), D.init_actor(rng), D.init_critic(rng)]
    if random_bn:
        for p in (nets[0], nets[2]):
            D.randomize_bn(p, rng, [("g1", "be1", "mu1", "var1"), ("g2", "be2", "mu2", "var2")])
        for p in (nets[1], nets[3]):
            D.randomize_bn(p, rng, [("gs", "bes", "mus", "vars"), ("ga", "bea", "mua", "vara"), ("g2", "be2", "mu2", "var2")])
        for p in nets:   # biases are zero-initialised in the reference; make them non-trivial for the check
            for k in p:
                if k.startswith("b") and not k.startswith("be"):
                    p[k] = rng.normal(0, 0.05, p[k].shape).astype(np.float32)
        nets[0]["W3"] *= 50; nets[1]["W3"] *= 300; nets[2]["W3"] *= 50; nets[3]["W3"] *= 300
    return nets


def make_batch(seed, B=64):
    rng = np.random.default_rng(seed)
    s = rng.normal(0, 2, (B, 4)).astype(np.float32); a = rng.uniform(-2.5, 2.5, (B, 1)).astype(np.float32)
    r = -rng.uniform(0, 0.5, (B, 1)).astype(np.float32); s2 = (s + rng.normal(0, 0.2, (B, 4))).astype(np.float32)
    return s, a, r, s2


def test_parameter_counts_match_survey():
    a, c = D.actor_shapes(), D.critic_shapes()
    n = lambda sh, names: sum(int(np.prod(sh[k])) for k in names)
    assert n(a, D.ACTOR_TRAINABLE) == 35_073 and n(a, D.ACTOR_WEIGHTS) == 35_841
    assert n(c, D.CRITIC_TRAINABLE) == 41_409 and n(c, D.CRITIC_WEIGHTS) == 42_273
    assert len(D.ACTOR_TRAINABLE) == 10 and len(D.ACTOR_WEIGHTS) == 14
    assert len(D.CRITIC_TRAINABLE) == 14 and len(D.CRITIC_WEIGHTS) == 20


@pytest.mark.parametrize("seed", [0, 1])
def test_learn_gradients_vs_torch_autograd(seed):
    actor, critic, t_actor, t_critic = make_nets(seed)
    s, a, r, s2 = make_batch(seed)
    cg, ag, info = D.learn(actor, critic, t_actor, t_critic, (s, a, r, s2), gamma=0.99)
    ta, tc, tta, ttc = _t(actor), _t(critic), _t(t_actor), _t(t_critic)
    S, A, R, S2 = (torch.tensor(x, dtype=torch.float64) for x in (s, a, r, s2))
    y = R + 0.99 * torch_critic(ttc, S2, torch_actor(tta, S2))
    closs = torch.mean((y - torch_critic(tc, S, A)) ** 2)
    cgt = torch.autograd.grad(closs, [tc[k] for k in D.CRITIC_TRAINABLE])
    aloss = -torch.mean(torch_critic(tc, S, torch_actor(ta, S)))
    agt = torch.autograd.grad(aloss, [ta[k] for k in D.ACTOR_TRAINABLE])
    assert abs(info["critic_loss"] - closs.item()) < 1e-5 * max(1, abs(closs.item()))
    assert abs(info["actor_loss"] - aloss.item()) < 1e-5
    for k, gt in zip(D.CRITIC_TRAINABLE, cgt):
        scale = max(gt.abs().max().item(), 1e-8)
        assert np.max(np.abs(cg[k].reshape(gt.shape) - gt.numpy())) / scale < 2e-4, k
    for k, gt in zip(D.ACTOR_TRAINABLE, agt):
        scale = max(gt.abs().max().item(), 1e-10)
        assert np.max(np.abs(ag[k].reshape(gt.shape) - gt.numpy())) / scale < 2e-4, k
    assert max(np.abs(v).max() for v in ag.values()) > 0   # gradient actually flows through the critic to the actor


def test_adam_matches_tf_keras_formula():
    rng = np.random.default_rng(3)
    p = {"w": rng.normal(size=100).astype(np.float32)}
    g = {"w": rng.normal(size=100).astype(np.float32)}
    m = {"w": np.zeros(100, np.float32)}; v = {"w": np.zeros(100, np.float32)}
    p0 = p["w"].copy()
    D.adam_apply(p, g, m, v, 1, 5e-4, ["w"])
    # after one step m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1)  => step = lr*g/(|g| + eps/sqrt(1-b2))
    expect = p0 - 5e-4 * g["w"] / (np.abs(g["w"]) + 1e-7 / np.sqrt(1 - 0.999))
    np.testing.assert_allclose(p["w"], expect, rtol=2e-5, atol=1e-8)
    # epsilon placement differs from torch.optim.Adam (eps added to sqrt(v_hat)): the two must NOT be identical
    tw = torch.tensor(p0.copy(), requires_grad=True); opt = torch.optim.Adam([tw], lr=5e-4, eps=1e-7)
    tw.grad = torch.tensor(g["w"]); opt.step()
    small = np.abs(g["w"]) < 1e-3
    assert np.allclose(tw.detach().numpy()[~small], p["w"][~small], rtol=1e-3, atol=1e-7)
    for t in range(2, 6):
        D.adam_apply(p, g, m, v, t, 5e-4, ["w"])
    assert np.isfinite(p["w"]).all()


def test_polyak_vs_reference_golden(golden):
    g = golden("polyak")
    for prefix, n in (("c", int(g["n_c"])), ("a", int(g["n_a"]))):
        names = [str(i) for i in range(n)]
        online = {str(i): g[f"{prefix}{i}"] for i in range(n)}
        target = {str(i): g[f"t{prefix}{i}"] for i in range(n)}
        new = D.polyak(target, online, float(g["tau"]), names)
        for i in range(n):
            np.testing.assert_allclose(new[str(i)], g[f"t{prefix}_new{i}"], rtol=1e-6, atol=1e-7)


def test_fedavg_vs_reference_golden(golden):
    g = golden("fedavg")
    # the reference's own demo inputs (src/server/test_federated.py:26-42)
    P, M, L = 2, 2, 3
    w = g["kat_weights"]
    weighted = [[[w[p][m] * g[f"kat_in_{p}_{m}_{l}"] for l in range(L)] for p in range(P)] for m in range(M)]
    plain = [[[g[f"kat_in_{p}_{m}_{l}"] for l in range(L)] for p in range(P)] for m in range(M)]
    wavg = D.fed_weighted_average(weighted, g["kat_sums"]); avg = D.fed_average(plain)
    for m in range(M):
        for l in range(L):
            np.testing.assert_allclose(wavg[m][l], g[f"kat_wavg_{m}_{l}"], rtol=1e-6)
            np.testing.assert_allclose(avg[m][l], g[f"kat_avg_{m}_{l}"], rtol=1e-6)
    # values quoted in SURVEY.md §8c
    np.testing.assert_allclose(wavg[0][0], [4, 5, 6], rtol=1e-6)
    np.testing.assert_allclose(wavg[1][1], [12.3846, 13.3846], rtol=1e-5)
    np.testing.assert_allclose(avg[1][0], [10, 11, 12])
    S, X, L = (int(v) for v in g["rnd_shape"])
    wts = g["rnd_weights"]
    weighted = [[[np.float32(wts[s][x]) * g[f"rnd_in_{s}_{x}_{l}"] for l in range(L)] for x in range(X)] for s in range(S)]
    plain = [[[g[f"rnd_in_{s}_{x}_{l}"] for l in range(L)] for x in range(X)] for s in range(S)]
    wavg = D.fed_weighted_average(weighted, wts.sum(1)); avg = D.fed_average(plain)
    for s in range(S):
        for l in range(L):
            np.testing.assert_allclose(wavg[s][l], g[f"rnd_wavg_{s}_{l}"], rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(avg[s][l], g[f"rnd_avg_{s}_{l}"], rtol=1e-6, atol=1e-7)
