"""Pin the CPU restatement (oracle/platoon_np.py) against fixtures produced by the reference's own
code (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import platoon_np as onp


def _prm(**over):
    return onp.EnvParams(**over)


def test_survey_known_answers(golden):
    """The literal values quoted in SURVEY.md §8c."""
    g = golden("default_trace")
    np.testing.assert_allclose(g["reset_obs"], [[2.19316191, -3.09021106, -0.01612086, 0.0],
                                                [1.70065416, -1.6498369, -0.00862141, -0.01612086]], atol=5e-9)
    np.testing.assert_allclose(g["ou"][0, 0], -0.00175572, atol=5e-9)  # first OU sample after reset
    np.random.seed(1)
    assert list(np.random.choice(100, 8)) == [37, 12, 72, 9, 75, 5, 79, 64]
    assert list(golden("replay")["choice_kat"]) == [37, 12, 72, 9, 75, 5, 79, 64]


def test_serial_port_bit_exact_default_trace(golden):
    """Same np.random seed, same draw order => identical float64 bits as the reference."""
    g = golden("default_trace")
    prm = _prm()
    np.random.seed(1)
    pl = onp.SerialPlatoon(2, prm)
    assert np.array_equal(np.stack([f.x for f in pl.followers]), g["ctor_x"])
    assert np.array_equal(np.stack(pl.reset()), g["reset_obs"])
    ous = [onp.SerialOUNoise(prm) for _ in range(2)]
    for k in range(g["mu"].shape[0]):
        acts = []
        for m in range(2):
            nz = ous[m]()
            assert nz[0] == g["ou"][k, m]
            a = onp.clip_action(np.float32([[g["mu"][k, m]]]).squeeze(), nz, prm.action_low, prm.action_high)
            acts.append(np.squeeze(a))
        assert np.array_equal(np.array(acts), g["actions"][k])
        ex = np.random.normal(0, prm.reset_max_u)
        assert ex == g["exog"][k]
        st, rw, dn = pl.step(acts, ex)
        assert np.array_equal(np.stack(st), g["obs"][k])
        assert np.array_equal(np.array(rw), g["reward"][k])
        assert dn == g["done"][k]


def test_terminal_known_answer(golden):
    g = golden("terminal")
    prm = _prm()
    np.random.seed(1)
    pl = onp.SerialPlatoon(2, prm)
    pl.reset()
    for m in range(2):
        pl.followers[m].x = g["x_before"][m].copy()
        pl.followers[m].prev_x = pl.followers[m].x
    st, rw, dn = pl.step([0.0, 0.0], 0.0)
    assert dn and rw[0] == -0.5 == g["reward"][0]
    assert np.array_equal(np.stack(st), g["obs"])
    assert st[0][0] == 25.0  # state still advanced: ep' = 25 + T*0 - hT*0


ROLLOUTS = [
    ("rollout_euler_M4", dict(can_terminate=False)),
    ("rollout_exact_M4", dict(can_terminate=False, method="exact")),
    ("rollout_exact_hetero_M3", dict(can_terminate=False, method="exact", pl_leader_tau=0.25, timegap=1.3, dyn_coeff=0.15)),
    ("rollout_terminating_M4", dict(max_ep=4.0, max_ev=4.0)),
    ("rollout_modelA_M3", dict(model="ModelA", can_terminate=False)),
    ("rollout_M8", dict(can_terminate=False)),
    ("rollout_central_M3", dict(framework="centralized", can_terminate=False)),
    ("rollout_leader_none_M2", dict(can_terminate=False)),
]


@pytest.mark.parametrize("name,over", ROLLOUTS)
def test_matrices_match_reference(golden, name, over):
    g = golden(name)
    mats = onp.follower_matrices(_prm(**over), g["A"].shape[0])
    for m, (A, B, C) in enumerate(mats):
        assert np.array_equal(A, g["A"][m]) and np.array_equal(B, g["B"][m]) and np.array_equal(C, g["C"][m])


@pytest.mark.parametrize("name,over", ROLLOUTS)
def test_serial_port_bit_exact_rollouts(golden, name, over):
    g = golden(name)
    prm = _prm(**over)
    M = g["actions"].shape[1]
    np.random.seed(1)
    pl = onp.SerialPlatoon(M, prm)
    s0 = pl.reset()
    assert np.array_equal(np.reshape(s0, g["reset_obs"].shape), g["reset_obs"])
    assert pl.front_u == g["front_u"]
    leader_none = name == "rollout_leader_none_M2"
    for k in range(g["actions"].shape[0]):
        st, rw, dn = pl.step(g["actions"][k], None if leader_none else g["exog"][k])
        assert np.array_equal(np.reshape(st, g["obs"][k].shape), g["obs"][k]), k
        assert np.array_equal(np.asarray(rw, dtype=np.float64).reshape(-1), g["reward"][k]), k
        assert dn == g["done"][k]
        assert np.array_equal(np.reshape(pl.jerks(), M), g["jerk"][k])
        assert np.array_equal([f.velocity for f in pl.followers], g["velocity"][k])
        assert np.array_equal([f.headway for f in pl.followers], g["headway"][k])
    if name == "rollout_terminating_M4":
        assert g["done"].any() and not g["done"].all()


@pytest.mark.parametrize("name,over", ROLLOUTS)
def test_batched_statement_matches_rollouts(golden, name, over):
    g = golden(name)
    prm = _prm(**over)
    M = g["actions"].shape[1]
    env = onp.BatchedPlatoons(1, M, prm)
    env.set_state(g["x0"][None], front_accel=[g["front_accel"]], front_u=[g["front_u"]])
    leader_none = name == "rollout_leader_none_M2"
    for k in range(g["actions"].shape[0]):
        obs, rew, done = env.step(g["actions"][k][None], None if leader_none else g["exog"][k:k + 1])
        np.testing.assert_allclose(obs[0], g["obs"][k], rtol=0, atol=1e-11)
        np.testing.assert_allclose(rew[0], g["reward"][k], rtol=0, atol=1e-13)
        assert done[0] == g["done"][k]
        np.testing.assert_allclose(env.st.jerk[0], g["jerk"][k], atol=1e-11)
        np.testing.assert_allclose(env.st.headway[0], g["headway"][k], atol=1e-10)


def test_batched_statement_multi_platoon(golden):
    g = golden("multi_platoon")
    P, M = g["x0"].shape[:2]
    env = onp.BatchedPlatoons(P, M, _prm())
    env.set_state(g["x0"], front_accel=g["front_accel"])
    for k in range(g["actions"].shape[0]):
        obs, rew, done = env.step(g["actions"][k], g["exog"][k])
        np.testing.assert_allclose(obs, g["obs"][k], rtol=0, atol=1e-10)
        np.testing.assert_allclose(rew, g["reward"][k], rtol=0, atol=1e-13)
        assert np.array_equal(done, g["done"][k])
    assert g["done"].any()


@pytest.mark.parametrize("tag,over,kw", [("normal", {}, {}), ("uniform", {"rand_gen": "uniform"}, {}),
                                         ("fixed", {}, {"rand_states": False}), ("eval", {}, {"eval_states": True})])
def test_reset_variants(golden, tag, over, kw):
    g = golden("resets")
    np.random.seed(11)
    pl = onp.SerialPlatoon(3, _prm(**over), **kw)
    assert np.array_equal(np.stack(pl.reset()), g[f"{tag}_first"])
    assert np.array_equal(np.stack(pl.reset()), g[f"{tag}_second"])
    assert pl.front_u == g[f"{tag}_front_u"] and pl.front_accel == g[f"{tag}_front_accel"]


def test_ou_noise(golden):
    g = golden("ou")
    prm = _prm()
    np.random.seed(5)
    ou = onp.SerialOUNoise(prm)
    xs = np.array([ou()[0] for _ in range(1000)])
    assert np.array_equal(xs, g["samples"])
    x = 0.0
    for k in range(1000):  # injected-draw form used to check the CUDA kernel
        x = onp.ou_step(x, g["z"][k], prm)
        assert abs(x - g["samples"][k]) < 1e-15


def test_replay_ring(golden):
    g = golden("replay")
    rb = onp.SerialReplay(128, 16, 4, 1)
    j = 0
    for i in range(300):
        rb.add(g["S"][i], g["A"][i], g["R"][i], g["S2"][i])
        if i in g["at"]:
            np.random.seed(100 + i)
            idx = rb.sample_indices()
            assert idx.dtype == np.int64 and np.array_equal(idx, g["idx"][j])
            s, a, r, s2 = rb.gather(idx)
            assert r.dtype == np.float32
            got = np.concatenate([s, a, r.astype(np.float64), s2], axis=1)
            assert np.array_equal(got, g["batch"][j])
            j += 1
    assert rb.count == g["counter"] == 300
    assert np.array_equal(rb.s, g["ring_s"]) and np.array_equal(rb.s2, g["ring_s2"])
    assert np.array_equal(rb.a, g["ring_a"]) and np.array_equal(rb.r, g["ring_r"])
