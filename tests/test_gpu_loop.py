"""Closed-loop parity of the whole training step: K steps of the reference's hot loop (workers/trainer.py:251-271 with
train_all_models 304-359 and the federated gradients round 400-431), restated on the CPU in oracle/loop_np.py from the pinned
pieces (platoon_np, ddpg_np, philox), beside `BatchedTrainer.step()` on the GPU -- same initial weights, same counter-based
draws (reset states, OU noise, leader inputs, replay indices).  Every step feeds the next one: actor -> action -> environment ->
replay ring -> sampled minibatch -> gradients -> Adam -> Polyak -> actor ..., so a mistake anywhere in the wiring (ring layout,
row order of the gather, which weights a pass reads, Adam step counters, FRL membership) shows up as a diverging trajectory.

Tolerances:
  precision 0 (fp32 SIMT)    states 2e-5 normwise, losses 1e-4, weight UPDATES (theta_K - theta_0) 5e-3 rel-L2 per tensor
                             (measured 3.7e-6 / 2.8e-5 / 2.8e-3)
  precision 2 (fp16 tcgen05) states 5e-4, losses 1e-2, weight updates 5e-2 (64 x E = 256 rows per update: the small-batch regime;
                             measured 1.4e-4 / 1.5e-3 / 1.9e-2)
"""
import numpy as np
import pytest
import torch

from oracle import ddpg_np as D
from oracle.loop_np import TrainLoopOracle

pytestmark = pytest.mark.gpu

K_STEPS = 50


def _nets(rng, n_agents):
    out = []
    for _ in range(n_agents):
        ac, cr = D.init_actor(rng), D.init_critic(rng)
        D.randomize_bn(ac, rng, [("g1", "be1", "mu1", "var1"), ("g2", "be2", "mu2", "var2")])
        D.randomize_bn(cr, rng, [("gs", "bes", "mus", "vars"), ("ga", "bea", "mua", "vara"), ("g2", "be2", "mu2", "var2")])
        ac["W3"] *= 50
        cr["W3"] *= 100
        out.append([ac, cr, {k: v.copy() for k, v in ac.items()}, {k: v.copy() for k, v in cr.items()}])
    return out


@pytest.mark.parametrize("precision,fed", [(0, None), (0, "interfrl"), (2, None), (2, "interfrl"), (2, "intrafrl")])
def test_closed_loop_parity(precision, fed):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from avddpg_b200.config import Config
    from avddpg_b200.trainer import BatchedTrainer
    G, E, M = 2, 4, 2
    conf = Config(pl_size=M, batch_size=16, buffer_size=64, can_terminate=False, fed_method=fed or "normal", weighted_average_enabled=False)
    tr = BatchedTrainer(conf, num_groups=G, envs_per_group=E, ring_capacity=64, precision=precision)
    nets = _nets(np.random.default_rng(7), M * G)
    for a, four in enumerate(nets):
        for bank, net in zip((tr.pop.actor, tr.pop.critic, tr.pop.t_actor, tr.pop.t_critic), four):
            bank.load_named(a, net)
    ora = TrainLoopOracle(conf, G, E, M, nets, seed=tr.env.seed, ring_capacity=64, fed=fed)
    st_tol, loss_tol, upd_tol = (2e-5, 1e-4, 5e-3) if precision == 0 else (5e-4, 1e-2, 5e-2)
    worst_state = worst_loss = 0.0
    n_learn = 0
    for k in range(K_STEPS):
        tr.step()
        obs, rew, _ = ora.step()
        got = tr.env.obs.cpu().numpy().astype(np.float64)
        worst_state = max(worst_state, float(np.max(np.abs(got - obs)) / np.max(np.abs(obs))))
        assert np.max(np.abs(tr.env.reward.cpu().numpy() - rew)) < 5e-5 + 10 * st_tol, k
        if ora.count > ora.batch:
            n_learn += 1
            loss = tr.pop.loss.cpu().numpy().astype(np.float64)          # [A][2]
            want = np.asarray(ora.losses[-1], dtype=np.float64)
            worst_loss = max(worst_loss, float(np.max(np.abs(loss - want) / np.maximum(np.abs(want), 1e-3))))
    assert n_learn == K_STEPS - 16 and tr.pop.actor.step.tolist() == [n_learn] * (M * G)
    assert worst_state < st_tol, worst_state
    assert worst_loss < loss_tol, worst_loss
    # the replay rings hold the same transitions
    ring = tr.rings.data.cpu().numpy()
    np.testing.assert_allclose(ring[:K_STEPS, ..., 4], ora.ring[:K_STEPS, ..., 4], atol=1e-4 + 50 * st_tol)          # actions (slots K.. were never written)
    # weights: error of the accumulated UPDATE, per tensor
    worst_upd, where = 0.0, None
    for a in range(M * G):
        for bank, ref, init, names in ((tr.pop.actor, ora.nets[a][0], nets[a][0], D.ACTOR_TRAINABLE), (tr.pop.critic, ora.nets[a][1], nets[a][1], D.CRITIC_TRAINABLE),
                                       (tr.pop.t_actor, ora.nets[a][2], nets[a][2], D.ACTOR_WEIGHTS), (tr.pop.t_critic, ora.nets[a][3], nets[a][3], D.CRITIC_WEIGHTS)):
            for name in names:
                got = bank.view(name, a).cpu().numpy().astype(np.float64).ravel()
                want, w0 = ref[name].astype(np.float64).ravel(), init[name].astype(np.float64).ravel()
                upd = np.linalg.norm(want - w0)
                if upd < 1e-5 * max(np.linalg.norm(w0), 1e-30):
                    # not trained: BatchNormalization moving statistics -- online nets never touch them, the targets' Polyak update
                    # tau * x + (1 - tau) * x only adds fp32 rounding noise; compare against the tensor itself
                    assert np.linalg.norm(got - want) <= 5e-6 * np.linalg.norm(w0), (a, bank.kind, name)
                    continue
                e = float(np.linalg.norm(got - want) / upd)
                if e > worst_upd:
                    worst_upd, where = e, (a, bank.kind, name)
    assert worst_upd < upd_tol, (worst_upd, where)
    if fed == "interfrl":                  # replicas of a follower that started identical would stay identical; here: same Adam step count only
        assert tr.fed.rounds == n_learn
    print(f"closed loop precision={precision} fed={fed}: states {worst_state:.1e} losses {worst_loss:.1e} weight updates {worst_upd:.1e} at {where}")
