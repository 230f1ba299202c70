"""GPU parity tests of the DDPG learn path (C ABI: avd_ddpg_learn, avd_actor_forward, avd_critic_forward,
avd_adam_apply, avd_polyak_update, avd_fed_*) against oracle/ddpg_np.py and the reference-generated goldens.

Tolerances (fp32 SIMT kernels, `precision=0`): forward values 1e-5 relative; gradients 2e-4 normwise per
tensor (the reductions over the batch run in a different order than NumPy's); parameters after Adam/Polyak
steps 1e-5.  The tensor-core modes (`precision=1` bf16 operands, `precision=2` fp16 operands -- the bench default) have their
own bars below (test_learn_gradients_tensor_core_mode, test_learn_gradients_fp16_mode).
"""
import numpy as np
import pytest
import torch

from oracle import ddpg_np as D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from avddpg_b200 import _lib, ddpgagent, model, replaybuffer, trainer
    from avddpg_b200.config import Config
    from avddpg_b200.server import federated
    _lib.require_device()
    return dict(lib=_lib, model=model, trainer=trainer, Config=Config, fed=federated, agent=ddpgagent, rb=replaybuffer)


def _nrm(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-12)


def _l2(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)


def make_nets(seed):
    rng = np.random.default_rng(seed)
    nets = [D.init_actor(rng), D.init_critic(rng), D.init_actor(rng), D.init_critic(rng)]
    for p in (nets[0], nets[2]):
        D.randomize_bn(p, rng, [("g1", "be1", "mu1", "var1"), ("g2", "be2", "mu2", "var2")])
    for p in (nets[1], nets[3]):
        D.randomize_bn(p, rng, [("gs", "bes", "mus", "vars"), ("ga", "bea", "mua", "vara"), ("g2", "be2", "mu2", "var2")])
    for p in nets:
        for k in p:
            if k.startswith("b") and not k.startswith("be"):
                p[k] = rng.normal(0, 0.05, p[k].shape).astype(np.float32)
    nets[0]["W3"] *= 50; nets[1]["W3"] *= 300; nets[2]["W3"] *= 50; nets[3]["W3"] *= 300
    return nets


def make_batch(seed, B):
    rng = np.random.default_rng(seed)
    s = rng.normal(0, 2, (B, 4)).astype(np.float32); a = rng.uniform(-2.5, 2.5, (B, 1)).astype(np.float32)
    r = -rng.uniform(0, 0.5, (B, 1)).astype(np.float32); s2 = (s + rng.normal(0, 0.2, (B, 4))).astype(np.float32)
    return s, a, r, s2


def build_population(mods, A, R, seeds, G=None, M=None):
    conf = mods["Config"]()
    G = A if G is None else G
    M = 1 if M is None else M
    pop = mods["trainer"].DDPGPopulation(G, M, conf, rows_per_agent=R)
    nets = [make_nets(sd) for sd in seeds]
    for a, (ac, cr, ta, tc) in enumerate(nets):
        pop.actor.load_named(a, ac); pop.critic.load_named(a, cr); pop.t_actor.load_named(a, ta); pop.t_critic.load_named(a, tc)
    batches = [make_batch(100 + sd, R) for sd in seeds]
    cat = lambda i, w: torch.as_tensor(np.concatenate([b[i].reshape(R, w) for b in batches]), device="cuda").contiguous()
    s, a, r, s2 = cat(0, 4), cat(1, 1).reshape(-1), cat(2, 1).reshape(-1), cat(3, 4)
    return conf, pop, nets, batches, (s, a, r, s2)


def test_forward_vs_oracle(mods):
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, 3, 70, [1, 2, 3])
    lib, d = mods["lib"], pop.dims
    n = 3 * 70
    out = torch.empty(n, device="cuda"); q = torch.empty(n, device="cuda")
    ws = torch.empty(n * (d.l1 + d.la + d.l2) * 4 + (d.l1 + d.la) * d.l2 * 2 * 3 + 1024, dtype=torch.uint8, device="cuda")
    lib.check(lib.load().avd_actor_forward(d, 3, 70, lib.ptr(pop.actor.flat), lib.ptr(s), 4, 1, 2.5, lib.ptr(out), lib.ptr(ws), ws.numel(), 0, lib.current_stream()))
    lib.check(lib.load().avd_critic_forward(d, 3, 70, lib.ptr(pop.critic.flat), lib.ptr(s), lib.ptr(a), lib.ptr(q), lib.ptr(ws), ws.numel(), 0, lib.current_stream()))
    for i in range(3):
        ref_a, _ = D.actor_forward(nets[i][0], batches[i][0])
        ref_q, _ = D.critic_forward(nets[i][1], batches[i][0], batches[i][1])
        assert _nrm(out[i * 70:(i + 1) * 70].cpu().numpy(), ref_a.ravel()) < 1e-5
        assert _nrm(q[i * 70:(i + 1) * 70].cpu().numpy(), ref_q.ravel()) < 1e-5


@pytest.mark.parametrize("A,R", [(1, 64), (3, 64), (2, 200), (2, 1000)])
def test_learn_gradients_vs_oracle(mods, A, R):
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, A, R, list(range(10, 10 + A)))
    cg, ag = pop.learn(s, a, r, s2, apply_updates=False)
    torch.cuda.synchronize()
    for i in range(A):
        ocg, oag, info = D.learn(nets[i][0], nets[i][1], nets[i][2], nets[i][3], batches[i], gamma=conf.gamma, high=conf.action_high)
        for name in pop.critic.trainable_names:
            got = pop.critic.view(name, i, pop.critic.grad).cpu().numpy()
            assert _nrm(got, ocg[name].reshape(got.shape)) < 2e-4, ("critic", name)
        for name in pop.actor.trainable_names:
            got = pop.actor.view(name, i, pop.actor.grad).cpu().numpy()
            assert _nrm(got, oag[name].reshape(got.shape)) < 2e-4, ("actor", name)
        loss = pop.loss[i].cpu().numpy()
        assert abs(loss[0] - info["critic_loss"]) < 1e-4 * max(1, abs(info["critic_loss"]))
        assert abs(loss[1] - info["actor_loss"]) < 1e-4 * max(1, abs(info["actor_loss"]))


@pytest.mark.parametrize("precision", [0, 2])
def test_learn_three_state_model_a(mods, precision):
    """Model A agents see 3 state words (environment.py:47-49); their batches still travel in the replay gather's 4-word rows
    (avd_learn_io.s_stride = 4, dims.ns = 3) and the 4th word -- a_lead, whatever it holds -- must not leak into the networks."""
    conf = mods["Config"](model="ModelA")
    R = 1000
    pop = mods["trainer"].DDPGPopulation(1, 1, conf, num_states=3, rows_per_agent=R, precision=precision)
    rng = np.random.default_rng(3)
    nets = [D.init_actor(rng, ns=3), D.init_critic(rng, ns=3), D.init_actor(rng, ns=3), D.init_critic(rng, ns=3)]
    nets[0]["W3"] *= 50; nets[1]["W3"] *= 300; nets[2]["W3"] *= 50; nets[3]["W3"] *= 300
    for bank, net in zip((pop.actor, pop.critic, pop.t_actor, pop.t_critic), nets):
        bank.load_named(0, net)
    s, a, r, s2 = make_batch(5, R)
    s3, s23 = s[:, :3].copy(), s2[:, :3].copy()
    s[:, 3], s2[:, 3] = 1e3, -1e3                      # garbage in the unused word
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    pop.learn(dev(s), dev(a.reshape(-1)), dev(r.reshape(-1)), dev(s2), apply_updates=False)
    torch.cuda.synchronize()
    ocg, oag, info = D.learn(nets[0], nets[1], nets[2], nets[3], (s3, a, r, s23), gamma=conf.gamma, high=conf.action_high, with_abs=True)
    for bank, ref, ab in ((pop.critic, ocg, info["critic_abs"]), (pop.actor, oag, info["actor_abs"])):
        for name in bank.trainable_names:
            got = bank.view(name, 0, bank.grad).cpu().numpy().astype(np.float64)
            want = ref[name].reshape(got.shape).astype(np.float64)
            err = np.linalg.norm(got - want) / max(np.linalg.norm(ab[name].astype(np.float64)), 1e-300)
            assert err < (2e-4 if precision == 0 else 8e-3), (bank.kind, name, err)


def test_learn_apply_updates_sequence(mods):
    """3 consecutive local updates (Adam x2 + Polyak) stay on the oracle's trajectory."""
    A, R = 2, 64
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, A, R, [21, 22])
    state = []
    for ac, cr, ta, tc in nets:
        z = lambda p, names: {k: np.zeros_like(p[k]) for k in names}
        state.append(dict(am=z(ac, D.ACTOR_TRAINABLE), av=z(ac, D.ACTOR_TRAINABLE), cm=z(cr, D.CRITIC_TRAINABLE), cv=z(cr, D.CRITIC_TRAINABLE)))
    for step in range(1, 4):
        pop.learn(s, a, r, s2, apply_updates=True)
        for i, (ac, cr, ta, tc) in enumerate(nets):
            ocg, oag, _ = D.learn(ac, cr, ta, tc, batches[i], gamma=conf.gamma, high=conf.action_high)
            D.adam_apply(cr, ocg, state[i]["cm"], state[i]["cv"], step, conf.critic_lr, D.CRITIC_TRAINABLE)
            D.adam_apply(ac, oag, state[i]["am"], state[i]["av"], step, conf.actor_lr, D.ACTOR_TRAINABLE)
            nets[i][3] = tc = D.polyak(tc, cr, conf.tau, D.CRITIC_WEIGHTS)
            nets[i][2] = ta = D.polyak(ta, ac, conf.tau, D.ACTOR_WEIGHTS)
    torch.cuda.synchronize()
    assert pop.actor.step.tolist() == [3, 3] and pop.critic.step.tolist() == [3, 3]
    for i, (ac, cr, ta, tc) in enumerate(nets):
        for bank, ref in ((pop.actor, ac), (pop.critic, cr), (pop.t_actor, ta), (pop.t_critic, tc)):
            for name in bank.weight_names:
                got = bank.view(name, i).cpu().numpy()
                np.testing.assert_allclose(got, ref[name].reshape(got.shape), rtol=2e-4, atol=2e-6, err_msg=f"{bank.kind}.{name}")


def test_adam_and_polyak_kernels(mods, golden):
    lib = mods["lib"]
    rng = np.random.default_rng(5)
    A, n, stride = 3, 1000, 1100
    p = rng.normal(size=(A, stride)).astype(np.float32); g = rng.normal(size=(A, n)).astype(np.float32)
    tp, tg = torch.as_tensor(p, device="cuda"), torch.as_tensor(g, device="cuda")
    m, v = torch.zeros(A, n, device="cuda"), torch.zeros(A, n, device="cuda")
    step = torch.tensor([0, 4, 9], dtype=torch.int32, device="cuda")
    mask = torch.tensor([1, 0, 1], dtype=torch.uint8, device="cuda")
    lib.check(lib.load().avd_adam_apply(lib.ptr(tp), stride, lib.ptr(tg), n, lib.ptr(m), lib.ptr(v), lib.ptr(step), lib.ptr(mask), A, n,
                                        5e-4, 0.9, 0.999, 1e-7, lib.current_stream()))
    assert step.tolist() == [1, 4, 10]
    for a, t in ((0, 1), (2, 10)):
        P = {"w": p[a, :n].copy()}; M = {"w": np.zeros(n, np.float32)}; V = {"w": np.zeros(n, np.float32)}
        D.adam_apply(P, {"w": g[a]}, M, V, t, 5e-4, ["w"])
        np.testing.assert_allclose(tp[a, :n].cpu().numpy(), P["w"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(tp[1].cpu().numpy(), p[1]) and np.array_equal(tp[:, n:].cpu().numpy(), p[:, n:])
    # Polyak against the REFERENCE's update_target outputs
    gp = golden("polyak")
    tcw = [gp[f"tc{i}"] for i in range(int(gp["n_c"]))]; cw = [gp[f"c{i}"] for i in range(int(gp["n_c"]))]
    taw = [gp[f"ta{i}"] for i in range(int(gp["n_a"]))]; aw = [gp[f"a{i}"] for i in range(int(gp["n_a"]))]
    tc_new, ta_new = mods["agent"].update_target(float(gp["tau"]), tcw, cw, taw, aw)
    for i, t in enumerate(tc_new):
        np.testing.assert_allclose(t.cpu().numpy(), gp[f"tc_new{i}"], rtol=1e-6, atol=1e-7)
    for i, t in enumerate(ta_new):
        np.testing.assert_allclose(t.cpu().numpy(), gp[f"ta_new{i}"], rtol=1e-6, atol=1e-7)


def test_policy_dropin(mods):
    out = mods["agent"].policy(torch.tensor([[3.1]], device="cuda"), None, -2.5, 2.5)
    assert isinstance(out, list) and float(out[0]) == 2.5
    out = mods["agent"].policy(torch.tensor([[0.5]], device="cuda"), lambda: np.array([0.25]), -2.5, 2.5)
    assert abs(float(out[0]) - 0.75) < 1e-7


def test_server_dropin_vs_reference_golden(mods, golden):
    g = golden("fedavg")
    srv = mods["fed"].Server("t", False)
    P, M, L = 2, 2, 3
    w = g["kat_weights"]
    weighted = [[[np.float32(w[p][m]) * g[f"kat_in_{p}_{m}_{l}"] for l in range(L)] for p in range(P)] for m in range(M)]
    plain = [[[g[f"kat_in_{p}_{m}_{l}"] for l in range(L)] for p in range(P)] for m in range(M)]
    wavg = srv.get_weighted_avg_params(weighted, g["kat_sums"]); avg = srv.get_avg_params(plain)
    for m in range(M):
        for l in range(L):
            np.testing.assert_allclose(wavg[m][l].cpu().numpy(), g[f"kat_wavg_{m}_{l}"], rtol=1e-6)
            np.testing.assert_allclose(avg[m][l].cpu().numpy(), g[f"kat_avg_{m}_{l}"], rtol=1e-6)
    S, X, L = (int(v) for v in g["rnd_shape"])
    plain = [[[g[f"rnd_in_{s}_{x}_{l}"] for l in range(L)] for x in range(X)] for s in range(S)]
    avg = srv.get_avg_params(plain)
    for s in range(S):
        for l in range(L):
            assert avg[s][l].shape == g[f"rnd_avg_{s}_{l}"].shape
            np.testing.assert_allclose(avg[s][l].cpu().numpy(), g[f"rnd_avg_{s}_{l}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("method", ["interfrl", "intrafrl"])
@pytest.mark.parametrize("weighted", [False, True])
def test_aggregator_gradients(mods, method, weighted):
    G, M, R = 3, 2, 64
    A = G * M
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, A, R, list(range(30, 30 + A)), G=G, M=M)
    conf.fed_method = method
    pop.learn(s, a, r, s2, apply_updates=False)
    ag0, cg0 = pop.actor.grad.cpu().numpy().copy(), pop.critic.grad.cpu().numpy().copy()
    agg = mods["fed"].FederatedAggregator(pop, conf)
    rng = np.random.default_rng(1)
    S, X = (M, G) if method == "interfrl" else (G, M)
    w = rng.uniform(0.2, 2.0, (S, X)).astype(np.float32) if weighted else None
    before = pop.actor.flat.clone()
    agg.aggregate_gradients(weights=w, apply=True)
    member = (lambda s_, x: s_ * G + x) if method == "interfrl" else (lambda s_, x: x * G + s_)
    for s_ in range(S):
        rows = [member(s_, x) for x in range(X)]
        if weighted:
            ea = (w[s_][:, None] * ag0[rows]).sum(0) * np.float32(1 / w[s_].sum()); ec = (w[s_][:, None] * cg0[rows]).sum(0) * np.float32(1 / w[s_].sum())
        else:
            ea, ec = ag0[rows].mean(0), cg0[rows].mean(0)
        for row in rows:
            np.testing.assert_allclose(pop.actor.grad[row].cpu().numpy(), ea, rtol=2e-5, atol=1e-9)
            np.testing.assert_allclose(pop.critic.grad[row].cpu().numpy(), ec, rtol=2e-5, atol=1e-9)
    assert pop.actor.step.tolist() == [1] * A and not torch.equal(before, pop.actor.flat)


@pytest.mark.parametrize("method,weighted,directional", [("interfrl", False, False), ("interfrl", True, False), ("intrafrl", True, True)])
def test_fused_frl_round_matches_unfused(mods, method, weighted, directional):
    """The fused consumer (avd_fed_apply_gradients: division by the divisor, tf.keras Adam, Polyak, step counters in ONE kernel after
    avd_fed_reduce2) leaves weights, targets, Adam moments and step counters where reduce -> finalize -> broadcast -> Adam/Polyak
    (trainer.py:400-431 spelled out) leaves them; three rounds so the Adam bias correction uses t = 1, 2, 3."""
    G, M, R = 3, 2, 64
    A = G * M
    pops = []
    for _ in range(2):
        conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, A, R, list(range(30, 30 + A)), G=G, M=M)
        conf.fed_method, conf.intra_directional_averaging = method, directional
        pops.append((conf, pop))
    rng = np.random.default_rng(3)
    S, X = (M, G) if method == "interfrl" else (G, M)
    aggs = [mods["fed"].FederatedAggregator(pop, conf) for conf, pop in pops]
    for rnd in range(3):
        w = rng.uniform(0.2, 2.0, (S, X)).astype(np.float32) if weighted else None
        for (conf, pop), agg, fused in zip(pops, aggs, (True, False)):
            pop.learn(s, a, r, s2, apply_updates=False)
            if fused:
                assert agg.aggregate_gradients(weights=w, apply=True, write_back=False) is None     # two launches, nothing materialised
            else:
                agg.aggregate_gradients(weights=w, apply=False)
                pop.apply_gradients_and_soft_update(agg.apply_mask)
        if weighted:
            np.testing.assert_allclose(aggs[0].last_weight_sums.cpu().numpy(), w.sum(1), rtol=1e-6)
    torch.cuda.synchronize()
    (_, p0), (_, p1) = pops
    # same arithmetic, but the two kernels may contract multiply-adds differently: a few ulp of the largest element per tensor
    close = lambda x, y, what: np.testing.assert_allclose(x, y, rtol=1e-5, atol=2e-6 * float(np.max(np.abs(y))), err_msg=what)
    for name in ("actor", "critic", "t_actor", "t_critic"):
        close(getattr(p0, name).flat.cpu().numpy(), getattr(p1, name).flat.cpu().numpy(), name)
    for bank0, bank1 in ((p0.actor, p1.actor), (p0.critic, p1.critic)):
        close(bank0.m.cpu().numpy(), bank1.m.cpu().numpy(), "m")
        close(bank0.v.cpu().numpy(), bank1.v.cpu().numpy(), "v")
        assert bank0.step.tolist() == bank1.step.tolist()
    want_steps = [0 if (directional and a_ < G) else 3 for a_ in range(A)]
    assert p0.actor.step.tolist() == want_steps and p0.critic.step.tolist() == want_steps


def test_weighted_fedavg_in_batched_trainer(mods):
    """Weighted FedAvg wired into the batched loop (trainer.py:331-333, 385-398, 358-359): from episode `weighted_window` on every
    member is weighted by |1 / mean(its last weighted_window episodic rewards)|, taken from the ring of finished-episode rewards the
    env kernel keeps; RewardLog receives fed_weights / fed_weight_sums like update_reward_list (521-528)."""
    from avddpg_b200 import results
    conf = mods["Config"](pl_size=2, batch_size=8, buffer_size=64, fed_method="interfrl", weighted_average_enabled=True, weighted_window=2,
                          episode_sim_time=1.25, can_terminate=False)
    assert conf.steps_per_episode == 12
    G, M = 3, 2
    tr = mods["trainer"].BatchedTrainer(conf, num_groups=G, envs_per_group=2, ring_capacity=64)
    log = results.RewardLog(conf, num_platoons=G, num_models=M)
    tr.run(5, log)
    torch.cuda.synchronize()
    assert len(log.all_ep_reward_lists[0][0]) == 5 and len(log.all_fed_weights[0][0]) == 3          # episodes 2, 3, 4 are weighted
    for k, ep in enumerate((2, 3, 4)):
        for m in range(M):
            ws = [abs(1.0 / np.mean(np.asarray(log.all_ep_reward_lists[g][m][ep - 2:ep], dtype=np.float32))) for g in range(G)]
            for g in range(G):
                np.testing.assert_allclose(log.all_fed_weights[g][m][k], ws[g], rtol=2e-6)
                np.testing.assert_allclose(log.all_fed_weight_sums[g][m][k], np.sum(ws), rtol=2e-6)
    assert len({float(log.all_fed_weights[g][0][0]) for g in range(G)}) == G                            # the platoons really differ
    for m in range(M):                                                                                  # replicas of a follower stay identical
        rows = tr.pop.actor.flat[m * G:(m + 1) * G]
        assert torch.equal(rows[0], rows[1]) and torch.equal(rows[0], rows[2])
    # the in-kernel history: last finished episode == what the log recorded for episode 3 after the reset that opened episode 4 ...
    # (run() resets at the START of an episode, so the ring holds episodes 2 and 3; episode 4 is still in ep_reward)
    hist = tr.env.ep_hist.reshape(2, M, G, 2).mean(dim=3).cpu().numpy()                                 # [slot][m][g]
    for g in range(G):
        for m in range(M):        # which slot holds which episode depends on how many resets preceded the run: compare as a set
            np.testing.assert_allclose(sorted(hist[:, m, g]), sorted(log.all_ep_reward_lists[g][m][2:4]), rtol=1e-6)


def test_frl_schedule_uses_the_step_inside_the_episode(mods):
    """is_valid_update_step is fed the in-episode step index `i` (trainer.py:251-266, 345), not a run-wide counter: with
    fed_update_delay_steps = 3 and 10-step episodes the rounds fall on i = 0, 3, 6, 9 of EVERY episode."""
    conf = mods["Config"](pl_size=2, batch_size=8, buffer_size=64, fed_method="interfrl", weighted_average_enabled=False,
                          fed_update_delay=0.35, episode_sim_time=1.05, can_terminate=False)
    assert conf.fed_update_delay_steps == 3 and conf.steps_per_episode == 10
    tr = mods["trainer"].BatchedTrainer(conf, num_groups=2, envs_per_group=1, ring_capacity=64)
    tr.run(2)
    # learn() starts once buffer_counter > 8, i.e. at i = 8 of episode 0: round at i = 9; episode 1: i = 0, 3, 6, 9
    assert tr.fed.rounds == 5
    # local updates happen on the other learn steps (i = 8 of episode 0; i = 1, 2, 4, 5, 7, 8 of episode 1): 7 local + 5 federated
    assert tr.pop.actor.step.tolist() == [12] * 4
    # free-running loop: the nominal episode clock wraps at steps_per_episode
    tr2 = mods["trainer"].BatchedTrainer(conf, num_groups=2, envs_per_group=1, ring_capacity=64)
    for _ in range(23):
        tr2.step()
    assert (tr2.episode, tr2.step_in_episode) == (2, 3) and tr2.fed.rounds == 1 + 4 + 1


def test_aggregator_weights_mode_and_quirk(mods):
    G, M, R = 2, 2, 64
    A = G * M
    conf, pop, nets, batches, _ = build_population(mods, A, R, list(range(40, 40 + A)), G=G, M=M)
    conf.fed_method = "interfrl"
    a0, c0 = pop.actor.flat.cpu().numpy().copy(), pop.critic.flat.cpu().numpy().copy()
    mods["fed"].FederatedAggregator(pop, conf, reference_weights_quirk=False).aggregate_weights()
    for m in range(M):
        rows = [m * G + g for g in range(G)]
        for row in rows:
            np.testing.assert_allclose(pop.actor.flat[row].cpu().numpy(), a0[rows].mean(0), rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(pop.t_critic.flat[row].cpu().numpy(), c0[rows].mean(0), rtol=1e-6, atol=1e-7)
    pop.actor.flat.copy_(torch.as_tensor(a0)); pop.critic.flat.copy_(torch.as_tensor(c0))
    mods["fed"].FederatedAggregator(pop, conf, reference_weights_quirk=True).aggregate_weights()   # trainer.py:442-456: `[0]`
    sys0 = [g for g in range(G)]
    for row in range(A):
        np.testing.assert_allclose(pop.actor.flat[row].cpu().numpy(), a0[sys0].mean(0), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(pop.t_actor.flat[row].cpu().numpy(), a0[sys0].mean(0), rtol=1e-6, atol=1e-7)
    # intrafrl + directional averaging leaves follower 0 untouched
    conf.fed_method = "intrafrl"; conf.intra_directional_averaging = True
    pop.actor.flat.copy_(torch.as_tensor(a0))
    mods["fed"].FederatedAggregator(pop, conf, reference_weights_quirk=False).aggregate_weights()
    for g in range(G):
        assert np.array_equal(pop.actor.flat[g].cpu().numpy(), a0[g])            # m = 0 rows
        np.testing.assert_allclose(pop.actor.flat[G + g].cpu().numpy(), a0[[g, G + g]].mean(0), rtol=1e-6, atol=1e-7)


def test_reference_shaped_learn_with_model_objects(mods, golden):
    """Trainer.learn(rbuffer, actor, critic, target_actor, target_critic) on drop-in objects, with the
    reference's own np.random.choice indices injected (identical inputs -> gradients vs oracle)."""
    conf = mods["Config"]()
    m = mods["model"]
    actor = m.get_actor(4, 1, conf.action_high, seed_int=1, layer1_size=256, layer2_size=128)
    critic = m.get_critic(4, 1, layer1_size=256, layer2_size=128, action_layer_size=48)
    t_actor = m.get_actor(4, 1, conf.action_high, seed_int=1, layer1_size=256, layer2_size=128)
    t_critic = m.get_critic(4, 1, layer1_size=256, layer2_size=128, action_layer_size=48)
    ac, cr, ta, tc = make_nets(77)
    for obj, ref, names in ((actor, ac, D.ACTOR_WEIGHTS), (t_actor, ta, D.ACTOR_WEIGHTS), (critic, cr, D.CRITIC_WEIGHTS), (t_critic, tc, D.CRITIC_WEIGHTS)):
        obj.set_weights([ref[k] for k in names])
        assert len(obj.weights) == len(names) and len(obj.trainable_variables) == len(names) - (4 if "W1" in ref else 6)
    g = golden("replay")
    rb = mods["rb"].ReplayBuffer(128, 16, 4, 1, 2)
    for i in range(128):
        rb.add((g["S"][i], g["A"][i], g["R"][i], g["S2"][i]))
    idx = g["idx"][1]   # drawn by the reference at i = 127 (range 128)
    cg, ag = mods["trainer"].Trainer(conf=conf).learn(rb, actor, critic, t_actor, t_critic, indices=idx)
    batch = (g["ring_s"][idx][:, :4], g["ring_a"][idx], g["ring_r"][idx], g["ring_s2"][idx])
    # ring arrays in the golden are the final (wrapped) state of a 300-add run; rebuild the 128-add state
    batch = (g["S"][:128][idx], g["A"][:128][idx], g["R"][:128][idx].reshape(-1, 1), g["S2"][:128][idx])
    ocg, oag, _ = D.learn(ac, cr, ta, tc, batch, gamma=conf.gamma, high=conf.action_high)
    assert len(cg) == 14 and len(ag) == 10
    for t, name in zip(cg, D.CRITIC_TRAINABLE):
        assert _nrm(t.cpu().numpy(), ocg[name].reshape(t.shape)) < 2e-4, name
    for t, name in zip(ag, D.ACTOR_TRAINABLE):
        assert _nrm(t.cpu().numpy(), oag[name].reshape(t.shape)) < 2e-4, name
    a_out = actor(torch.as_tensor(batch[0], dtype=torch.float32))
    assert a_out.shape == (16, 1) and _nrm(a_out.cpu().numpy(), D.actor_forward(ac, batch[0])[0]) < 1e-5
    q_out = critic([batch[0], batch[1]])
    assert _nrm(q_out.cpu().numpy(), D.critic_forward(cr, batch[0], batch[1])[0]) < 1e-5


def test_population_act_on_native_state(mods):
    from avddpg_b200.environment import BatchedPlatoons
    conf = mods["Config"](pl_size=3)
    G, M, E = 2, 3, 50
    pop = mods["trainer"].DDPGPopulation(G, M, conf)
    nets = [make_nets(60 + a) for a in range(G * M)]
    for a, (ac, _, _, _) in enumerate(nets):
        pop.actor.load_named(a, ac)
    env = BatchedPlatoons(G * E, M, conf)
    env.reset()
    pop.act(env.native_state, env.action_mu, envs_per_group=E)
    st = env.state.cpu().numpy()      # [P, M, 4]
    mu = env.action_mu.cpu().numpy()   # [M, P]
    for m in range(M):
        for g in range(G):
            ref, _ = D.actor_forward(nets[m * G + g][0], st[g * E:(g + 1) * E, m])
            assert _nrm(mu[m, g * E:(g + 1) * E], ref.ravel()) < 1e-5


def test_batched_trainer_end_to_end(mods):
    """act -> env -> replay add -> sample -> learn -> Adam -> Polyak for a small population; weights move, the
    replay rings fill, statistics stay finite; CUDA-graph replay continues the same trajectory."""
    conf = mods["Config"](pl_size=2, batch_size=8, buffer_size=64)
    tr = mods["trainer"].BatchedTrainer(conf, num_groups=2, envs_per_group=3, ring_capacity=64)
    w0 = tr.pop.actor.flat.clone()
    for _ in range(8):
        tr.step()
    assert torch.equal(w0, tr.pop.actor.flat)            # nothing learned until buffer_counter > batch_size (trainer.py:322)
    for _ in range(6):
        tr.step()
    torch.cuda.synchronize()
    assert not torch.equal(w0, tr.pop.actor.flat) and torch.isfinite(tr.pop.actor.flat).all() and torch.isfinite(tr.pop.critic.flat).all()
    assert tr.pop.actor.step.tolist() == [6] * 4
    clk = tr.rings.clock.read()
    assert clk["ring_count"] == 14 and clk["step_tick"] == 14 and clk["update_tick"] == 6
    tr.capture(warmup=1)
    before = tr.pop.actor.flat.clone()
    tr.replay(); tr.replay()
    torch.cuda.synchronize()
    assert not torch.equal(before, tr.pop.actor.flat) and torch.isfinite(tr.pop.actor.flat).all()
    assert tr.rings.clock.read()["ring_count"] == 14 + 1 + 4 == tr.buffer_counter   # capture itself executes nothing


def test_batched_trainer_interfrl_keeps_replicas_identical(mods):
    """interfrl + gradients with delay 1 == synchronous data-parallel SGD: all platoons' agents for a follower
    start identical (trainer.py:121-131) and must stay identical after federated rounds."""
    conf = mods["Config"](pl_size=2, batch_size=8, buffer_size=64, fed_method="interfrl", weighted_average_enabled=False)
    tr = mods["trainer"].BatchedTrainer(conf, num_groups=3, envs_per_group=1, ring_capacity=64)
    for _ in range(14):
        tr.step()
    torch.cuda.synchronize()
    G = 3
    for m in range(2):
        rows = tr.pop.actor.flat[m * G:(m + 1) * G]
        assert torch.equal(rows[0], rows[1]) and torch.equal(rows[0], rows[2])
    assert tr.fed.rounds == 6 and tr.pop.actor.step.tolist() == [6] * 6


# ------------------------------------------------------------------------------------------ bf16 tensor-core mode
@pytest.mark.parametrize("A,R", [(1, 64), (3, 64), (2, 1000), (1, 4096), (1, 250_000), (3, 100_003), (200, 64)])
def test_learn_gradients_tensor_core_mode(mods, A, R):
    """precision=1 (legacy bf16 operands; superseded by precision=2 as the bench default, kept for its unbounded operand range):
    every contraction of the learn step runs on tcgen05 (bf16 operands, fp32 TMEM accumulation; layer 1 with
    hi/lo-split operands), heads / losses / reductions in fp32.  Bar against the fp32 oracle, per gradient tensor: relative L2
    error < 6e-2 (max error < 1.2e-1) on single 64-row minibatches, where the critic gradient is driven by the TD error q - y, a
    small difference of two bf16-noisy values (measured 0.05 % - 4.5 %), and < 2e-2 (4e-2) from 1000 rows up (measured < 0.7 %,
    profiles/r01_tensor_core_accuracy.txt).  The actor gradient of a single 64-row minibatch is bounded at 3e-1: it is
    proportional to d q / d action, a 128-term bf16 sum with cancellation, whose per-row signs cancel again in the reductions over
    the batch (seed 111: sum dq = -2.2e-4 against sum |dq| = 6.8e-3), and reaches 23 % for the worst of 200 seeds (typical
    0.5 % - 2 %); with the bench's 262,144 rows per update it is 0.5 %.  Losses agree to 1e-3.  precision=0 is the parity mode (2e-4).
    The two large cases give every persistent CTA 6-14 row tiles (ragged last tile, several agents per launch), so the mbarrier
    phase arithmetic of the software pipelines wraps many times; (200, 64) has more agents than SMs (one CTA per agent, two waves)."""
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, A, R, list(range(50, 50 + A)))
    pop.precision = 1
    pop.learn(s, a, r, s2, apply_updates=False)
    torch.cuda.synchronize()
    bad = []
    for i in range(A):
        ocg, oag, info = D.learn(nets[i][0], nets[i][1], nets[i][2], nets[i][3], batches[i], gamma=conf.gamma, high=conf.action_high)
        for bank, ref in ((pop.critic, ocg), (pop.actor, oag)):
            if R >= 1000:
                l2_tol, mx_tol = 2e-2, 4e-2
            else:                # one 64-row minibatch: see the docstring
                l2_tol, mx_tol = (6e-2, 1.2e-1) if bank is pop.critic else (3e-1, 4e-1)
            # tensors whose true gradient nearly cancels are measured against a quarter of the norm of the net's whole gradient
            # instead of their own, tiny, norm (seed 111: b3 = sum of signed dq = -2.2e-4 with sum |dq| = 6.8e-3)
            floor = 0.25 * float(np.sqrt(sum(np.sum(np.square(ref[n].astype(np.float64))) for n in bank.trainable_names)))
            for name in bank.trainable_names:
                got = bank.view(name, i, bank.grad).cpu().numpy().astype(np.float64)
                want = ref[name].reshape(got.shape).astype(np.float64)
                e2 = float(np.linalg.norm(got - want) / max(np.linalg.norm(want), floor))
                em = float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), floor / np.sqrt(got.size)))
                if not (e2 < l2_tol and em < mx_tol):
                    bad.append((i, bank.kind, name, round(e2, 4), round(em, 4)))
        loss = pop.loss[i].cpu().numpy()
        assert abs(loss[0] - info["critic_loss"]) < 1e-3 * max(1, abs(info["critic_loss"]))
        assert abs(loss[1] - info["actor_loss"]) < 1e-3 * max(1, abs(info["actor_loss"]))
    assert not bad, f"(agent, net, tensor, rel-L2, rel-max) out of tolerance: {bad}"


def _fp16_errors(mods, A, R, seeds):
    """-> (rel-L2, condition-normalised) error of every (agent, net, tensor) of one precision = 2 learn step vs the fp32 oracle.
    condition-normalised = ||got - ref|| / || sum_n |g_n| ||: the error against the sum of the MAGNITUDES of the per-row
    contributions (oracle.ddpg_np.learn(with_abs=True)), i.e. the scale a rounding error of the sum lives on; it equals rel-L2
    for a tensor whose rows all pull the same way and stays meaningful when the signed sum cancels."""
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, A, R, seeds)
    pop.precision = 2
    pop.learn(s, a, r, s2, apply_updates=False)
    torch.cuda.synchronize()
    rel, cond, names = [], [], []
    for i in range(A):
        ocg, oag, info = D.learn(nets[i][0], nets[i][1], nets[i][2], nets[i][3], batches[i], gamma=conf.gamma, high=conf.action_high, with_abs=True)
        for bank, ref, ab in ((pop.critic, ocg, info["critic_abs"]), (pop.actor, oag, info["actor_abs"])):
            for name in bank.trainable_names:
                got = bank.view(name, i, bank.grad).cpu().numpy().astype(np.float64)
                want = ref[name].reshape(got.shape).astype(np.float64)
                err = float(np.linalg.norm(got - want))
                rel.append(err / max(float(np.linalg.norm(want)), 1e-300))
                cond.append(err / max(float(np.linalg.norm(ab[name].astype(np.float64))), 1e-300))
                names.append((i, bank.kind, name))
        loss = pop.loss[i].cpu().numpy()
        assert abs(loss[0] - info["critic_loss"]) < 2e-4 * max(1, abs(info["critic_loss"]))
        assert abs(loss[1] - info["actor_loss"]) < 2e-4 * max(1, abs(info["actor_loss"]))
    return np.array(rel), np.array(cond), names


@pytest.mark.parametrize("A,R", [(1, 64), (3, 64), (200, 64), (2, 1000), (8, 1000), (1, 4096), (1, 16384), (1, 250_000), (3, 100_003)])
def test_learn_gradients_fp16_mode(mods, A, R):
    """precision = 2, the bench default: the layer-2 products, the backward tile and the dgrad run on tcgen05 with fp16 operands
    (11-bit significand, power-of-two operand scales), layer 1 with hi/lo-split bf16, fp32 accumulation in TMEM, heads / losses /
    reductions in fp32.  Tolerance against the fp32 oracle, per gradient tensor:
      >= 16384 rows per update :  rel-L2 <= 2e-3 for EVERY tensor                   (measured <= 7.4e-4; bench: 262,144 rows, <= 4.4e-4)
      1000 .. 16383 rows       :  condition-normalised error <= 8e-3, rel-L2 <= 2e-2
      one 64-row minibatch     :  condition-normalised error <= 4e-2 for every tensor and rel-L2 <= 1e-2 for >= 90 % of them.
    At 64 rows a ReLU whose pre-activation flips sign under the operand rounding changes a row's contribution by a discrete amount
    (1/64 of the batch), and some tensors (b3, beta2 of the actor) are signed sums that nearly cancel, so rel-L2 <= 1e-2 cannot hold
    for every seed: the bit-level model of this path (tools/precision_model.py) gives, over 200 seeds, a median of 2.5e-4, 95 % of
    the tensors below 1e-2 and a worst condition-normalised error of 2.9e-2.  No tensor is excused and there is no norm floor.
    The large cases give every persistent CTA 6-14 row tiles (ragged last tile, several agents per launch); (200, 64) has more
    agents than SMs."""
    rel, cond, names = _fp16_errors(mods, A, R, list(range(50, 50 + A)))
    worst = sorted(zip(rel, cond, names), reverse=True)[:4]
    if R >= 16384:
        assert rel.max() <= 2e-3, worst
    elif R >= 1000:
        assert cond.max() <= 8e-3 and rel.max() <= 2e-2, worst
    else:
        assert cond.max() <= 4e-2, worst
        assert np.quantile(rel, 0.9) <= 1e-2, (float(np.quantile(rel, 0.9)), worst)


def test_fp16_mode_survives_out_of_range_operands(mods):
    """fp16 has 5 exponent bits: activations beyond 65504 saturate (F2FP.SATFINITE) instead of becoming inf, and the backward
    tile / W2'' / T packs are scaled by powers of two, so a batch with huge states or a huge TD error still yields finite gradients."""
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, 1, 256, [77])
    pop.precision = 2
    s, s2 = s * 3e4, s2 * 3e4              # layer-1 outputs ~1e5
    r = r * 1e6                            # TD error ~1e6
    pop.learn(s, a, r, s2, apply_updates=False)
    torch.cuda.synchronize()
    assert torch.isfinite(pop.critic.grad).all() and torch.isfinite(pop.actor.grad).all() and torch.isfinite(pop.loss).all()


def test_forward_tensor_core_mode(mods):
    conf, pop, nets, batches, (s, a, r, s2) = build_population(mods, 2, 300, [71, 72])
    lib, d = mods["lib"], pop.dims
    n = 2 * 300
    out = torch.empty(n, device="cuda"); q = torch.empty(n, device="cuda")
    ws = torch.empty(n * (d.l1 + d.la + d.l2) * 4 + 2 * (d.l1 + d.la) * d.l2 * 2 + 1024, dtype=torch.uint8, device="cuda")
    lib.check(lib.load().avd_actor_forward(d, 2, 300, lib.ptr(pop.actor.flat), lib.ptr(s), 4, 1, 2.5, lib.ptr(out), lib.ptr(ws), ws.numel(), 1, lib.current_stream()))
    lib.check(lib.load().avd_critic_forward(d, 2, 300, lib.ptr(pop.critic.flat), lib.ptr(s), lib.ptr(a), lib.ptr(q), lib.ptr(ws), ws.numel(), 1, lib.current_stream()))
    for i in range(2):
        ref_a, _ = D.actor_forward(nets[i][0], batches[i][0])
        ref_q, _ = D.critic_forward(nets[i][1], batches[i][0], batches[i][1])
        assert _nrm(out[i * 300:(i + 1) * 300].cpu().numpy(), ref_a.ravel()) < 1e-2
        assert _nrm(q[i * 300:(i + 1) * 300].cpu().numpy(), ref_q.ravel()) < 1e-2


def test_batched_trainer_tensor_core_mode(mods):
    conf = mods["Config"](pl_size=2, batch_size=8, buffer_size=64)
    tr = mods["trainer"].BatchedTrainer(conf, num_groups=2, envs_per_group=40, ring_capacity=64, precision=1)
    w0 = tr.pop.critic.flat.clone()
    for _ in range(14):
        tr.step()
    torch.cuda.synchronize()
    assert not torch.equal(w0, tr.pop.critic.flat) and torch.isfinite(tr.pop.actor.flat).all() and torch.isfinite(tr.pop.critic.flat).all()
    tr.capture(warmup=1)
    tr.replay()
    torch.cuda.synchronize()
    assert torch.isfinite(tr.pop.actor.flat).all()


def test_host_step_pipeline_matches_synchronous_stepping(mods):
    """HostStepPipeline (double-buffered pinned inputs / results, results read one step late) returns for every step exactly
    what a synchronous host loop reads (same seeds, same host inputs): pipelining the hand-off changes no arithmetic."""
    conf = mods["Config"](pl_size=2, batch_size=8, buffer_size=64)
    mk = lambda: mods["trainer"].BatchedTrainer(conf, num_groups=2, envs_per_group=3, ring_capacity=64)
    n = 14
    inputs = 0.1 * torch.randn(n, 6, generator=torch.Generator().manual_seed(5))
    # synchronous loop
    tr = mk()
    h_in = torch.zeros(6).pin_memory()
    want = []
    for k in range(n):
        h_in.copy_(inputs[k])
        tr.env.stats.zero_()
        tr.step(host_leader_exog=h_in)
        want.append(torch.cat([tr.env.stats.cpu(), tr.pop.loss.reshape(-1).cpu()]))
    w_sync = tr.pop.actor.flat.clone()
    # pipelined loop
    tr2 = mk()
    pipe = mods["trainer"].HostStepPipeline(tr2)
    assert pipe.h2d_bytes_per_step == 6 * 4 and pipe.d2h_bytes_per_step == want[0].numel() * 4
    got = []
    for k in range(n):
        pipe.input_buffer().copy_(inputs[k])
        prev = pipe.submit()
        assert (prev is None) == (k == 0)
        if prev is not None:
            got.append(prev.clone())
    got.append(pipe.drain().clone())
    assert len(got) == n
    for k in range(n):
        assert torch.equal(got[k], want[k]), k
    assert torch.equal(w_sync, tr2.pop.actor.flat)


def test_evaluator_rollout_vs_reference_golden(mods, golden):
    """N2: noise-free evaluation rollout (workers/evaluator.py:40-96,145) on the reference's own seed-6 leader inputs
    and fixed evaluator initial state; actors injected.  precision=0 so the 100-step closed loop stays on the
    reference trajectory; pl_rew is the value the reference reports (3 decimals)."""
    from avddpg_b200 import evaluator
    g = golden("evaluator")
    conf = mods["Config"](pl_size=3)
    pop = mods["trainer"].DDPGPopulation(1, 3, conf)
    for m in range(3):
        pop.actor.load_named(m, {k[len(f"actor{m}_"):]: v for k, v in g.items() if k.startswith(f"actor{m}_")})
    T = int(g["T"])
    pl_rew, tr = evaluator.run(conf, pop, num_platoons=2, leader_inputs=np.repeat(g["inputs"][:, None], 2, axis=1),
                               manual_timestep_override=T, precision=0, return_traces=True)
    st = tr["states"].cpu().numpy()
    assert np.array_equal(st[:, 0], st[:, 1])                 # both platoons saw the same inputs
    assert _nrm(st[:, 0], g["states"]) < 2e-5 and _nrm(tr["inputs"][:, 0].cpu().numpy(), g["actions"]) < 2e-5
    assert _nrm(tr["jerks"][:, 0].cpu().numpy(), g["jerks"]) < 5e-4
    np.testing.assert_allclose(tr["rewards"][0].cpu().numpy(), g["ep_reward"], rtol=2e-5)
    assert abs(pl_rew - float(g["pl_rew"])) <= 1e-3
    # device-drawn leader inputs: runs, finite, same initial state
    pl2 = evaluator.run(conf, pop, num_platoons=64, manual_timestep_override=20, precision=1)
    assert np.isfinite(pl2)


def test_batched_trainer_reference_episode_loop_and_csvs(mods, tmp_path):
    """Trainer.run-shaped episode loop (trainer.py:232-273) on the batched trainer + the reference's CSV / conf.json formats."""
    from avddpg_b200 import results
    conf = mods["Config"](pl_size=2, episode_sim_time=3, can_terminate=False, reward_averaging_window=2)
    assert conf.steps_per_episode == 30
    tr = mods["trainer"].BatchedTrainer(conf, num_groups=3, envs_per_group=2, ring_capacity=128, precision=0)
    log = results.RewardLog(conf, num_platoons=3, num_models=2)
    steps = tr.run(4, reward_log=log, learn=False)
    assert steps == 4 * 30 and tr.env.auto_reset is True          # restored afterwards
    ep = np.array(log.all_ep_reward_lists, dtype=np.float64)      # [3][2][4]
    assert ep.shape == (3, 2, 4) and np.isfinite(ep).all() and (ep < 0).all()
    # the logged value is the sum over the episode of the per-step rewards (trainer.py:321), averaged over the group's platoons
    tr.env.auto_reset = False
    tr.env.reset()
    acc = torch.zeros(2, 6, device="cuda")
    for _ in range(30):
        tr.step(learn=False)
        acc += tr.env._reward          # raw [M, P] layout
    tr.env.auto_reset = True
    want = acc.reshape(2, 3, 2).mean(dim=2).t().cpu().numpy()
    got = np.array(tr.episodic_rewards(), dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=1e-5)
    paths = log.generate_csvs(str(tmp_path))
    rows = open(paths[1]).read().strip().splitlines()
    assert rows[0] == ",Vehicle 1,Vehicle 2,seed,platoon" and len(rows) == 1 + 3 * 4
    results.config_writer(str(tmp_path / "conf.json"), conf)
    assert results.config_loader(str(tmp_path / "conf.json")).pl_size == 2


@pytest.mark.parametrize("la", [16, 32, 64])
def test_learn_tensor_core_mode_other_action_layer_sizes(mods, la):
    """The persistent kernels take the critic's action-branch width at run time (16 ... 64 columns: partial last k-block /
    dgrad chunk / weight-gradient slab); 48 is the reference default covered above."""
    A, R = 2, 700
    conf = mods["Config"](critic_act_layer_size=la)
    pop = mods["trainer"].DDPGPopulation(A, 1, conf, rows_per_agent=R, precision=1)
    assert pop.dims.la == la
    rng = np.random.default_rng(90 + la)
    nets = []
    for a in range(A):
        n4 = [D.init_actor(rng), D.init_critic(rng, la=la), D.init_actor(rng), D.init_critic(rng, la=la)]
        D.randomize_bn(n4[0], rng, [("g1", "be1", "mu1", "var1"), ("g2", "be2", "mu2", "var2")])
        D.randomize_bn(n4[1], rng, [("gs", "bes", "mus", "vars"), ("ga", "bea", "mua", "vara"), ("g2", "be2", "mu2", "var2")])
        n4[1]["W3"] *= 300
        n4[0]["W3"] *= 50
        for bank, net in zip((pop.actor, pop.critic, pop.t_actor, pop.t_critic), n4):
            bank.load_named(a, net)
        nets.append(n4)
    batches = [make_batch(300 + a, R) for a in range(A)]
    cat = lambda i, w: torch.as_tensor(np.concatenate([b[i].reshape(R, w) for b in batches]), device="cuda").contiguous()
    pop.learn(cat(0, 4), cat(1, 1).reshape(-1), cat(2, 1).reshape(-1), cat(3, 4), apply_updates=False)
    torch.cuda.synchronize()
    bad = []
    for i in range(A):
        ocg, oag, info = D.learn(nets[i][0], nets[i][1], nets[i][2], nets[i][3], batches[i], gamma=conf.gamma, high=conf.action_high)
        for bank, ref in ((pop.critic, ocg), (pop.actor, oag)):
            for name in bank.trainable_names:
                got = bank.view(name, i, bank.grad).cpu().numpy()
                e2 = _l2(got, ref[name].reshape(got.shape))
                if not e2 < 4e-2:
                    bad.append((i, bank.kind, name, round(e2, 4)))
    assert not bad, bad


def test_learn_edge_cases_and_errors(mods):
    """Empty population is a no-op; bad arguments come back as ValueError through the C ABI's error convention."""
    lib, conf = mods["lib"], mods["Config"]()
    pop = mods["trainer"].DDPGPopulation(1, 1, conf, rows_per_agent=64, precision=1)
    s = torch.zeros(64, 4, device="cuda"); a = torch.zeros(64, device="cuda")
    pop.learn(s, a, a, s, apply_updates=False)
    io = pop.io
    io.A = 0
    assert lib.load().avd_ddpg_learn(io, lib.current_stream()) == 0               # no agents: nothing to do
    io.A, io.precision = 1, 7
    with pytest.raises(ValueError, match="precision"):
        lib.check(lib.load().avd_ddpg_learn(io, lib.current_stream()))
    io.precision, io.rows_per_agent = 1, 0
    with pytest.raises(ValueError):
        lib.check(lib.load().avd_ddpg_learn(io, lib.current_stream()))
    io.rows_per_agent, io.workspace_bytes = 64, 16
    with pytest.raises(ValueError, match="workspace"):
        lib.check(lib.load().avd_ddpg_learn(io, lib.current_stream()))
    with pytest.raises(ValueError):
        pop.learn(s[:10], a, a, s)                                                # shape mismatch caught on the host side
    torch.cuda.synchronize()


def test_learn_full_size_batch_additivity(mods):
    """Size-independent property at the BASELINE C2 size (262,144 rows per agent-update): every loss is a MEAN over the rows, so
    the gradient of the full batch equals the average of the gradients of its two halves, and the losses average the same way.
    The three learn calls tile the rows differently across the persistent CTAs (55 vs. 28 tiles per CTA)."""
    conf = mods["Config"]()
    R = 262_144
    pop = mods["trainer"].DDPGPopulation(1, 1, conf, rows_per_agent=R, precision=1)
    pop.t_actor.flat.copy_(pop.actor.flat * 1.01)
    pop.t_critic.flat.copy_(pop.critic.flat * 0.99)
    g = torch.Generator(device="cuda").manual_seed(3)
    s = torch.randn(R, 4, device="cuda", generator=g) * 2
    a = torch.rand(R, device="cuda", generator=g) * 5 - 2.5
    r = -torch.rand(R, device="cuda", generator=g) * 0.5
    s2 = s + 0.1 * torch.randn(R, 4, device="cuda", generator=g)

    def run(lo, hi):
        pop.learn(s[lo:hi].contiguous(), a[lo:hi].contiguous(), r[lo:hi].contiguous(), s2[lo:hi].contiguous(), apply_updates=False,
                  rows_per_agent=hi - lo)
        return pop.critic.grad.clone(), pop.actor.grad.clone(), pop.loss.clone()

    cf, af, lf = run(0, R)
    c1, a1, l1 = run(0, R // 2)
    c2, a2, l2 = run(R // 2, R)
    for full, h1, h2 in ((cf, c1, c2), (af, a1, a2), (lf, l1, l2)):
        want = 0.5 * (h1 + h2)
        err = float((full - want).norm() / want.norm())
        assert err < 2e-3, err          # fp32 accumulation order only: both sides carry the same bf16 operand rounding
