"""Drop-in demonstration: the reference's training loop, written call for call against the PER-OBJECT drop-ins
(avddpg_b200.reference_loop.PerObjectTrainer: one Platoon / OUActionNoise / ReplayBuffer / actor / critic / Adam object per platoon and
follower, exactly the objects workers/trainer.py:61-179 builds and 223-456 drives), lands where the batched GPU loop lands.

Both sides see the same counter-based draws -- reset states (platoon id, episode), OU noise (vehicle id, step), replay indices (ring id,
update) by construction, the leaders' inputs by injection -- and start from the same weights, so after 40 steps (24 learn steps) the
networks must agree to fp32 round-off accumulated through Adam: the per-object path runs one tiny launch per object, the batched one
whole-population kernels, but the arithmetic is the same.  (The reference's own workers/trainer.py cannot run here: it needs
/root/reference, which the GPU box does not have, and this package has no CPU path -- tests/test_reference_surface.py checks the call
surface statically in the build container instead.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fed", ["normal", "interfrl", "intrafrl"])
def test_per_object_reference_loop_matches_batched_trainer(fed):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from avddpg_b200.config import Config
    from avddpg_b200.reference_loop import PerObjectTrainer
    from avddpg_b200.trainer import BatchedTrainer
    from oracle import philox_env_np as penv
    from oracle import platoon_np as onp
    P, M, STEPS = 2, 2, 40
    mk = lambda: Config(num_platoons=P, pl_size=M, batch_size=16, buffer_size=64, can_terminate=False, fed_method=fed,
                        weighted_average_enabled=False, episode_sim_time=STEPS * 0.1 + 0.05)
    conf = mk()
    assert conf.steps_per_episode == STEPS
    prm = onp.EnvParams.from_config(conf)
    exog = lambda p, k: float(penv.leader_exog(prm, P, conf.random_seed, 0, tick=k)[p])
    per = PerObjectTrainer(conf, leader_exog_fn=exog).initialize()

    tr = BatchedTrainer(mk(), num_groups=P, envs_per_group=1, ring_capacity=64, precision=0)
    for p in range(P):
        for m in range(M):
            a = m * P + p
            tr.pop.actor.set_weights(a, per.actors[p][m].get_weights())
            tr.pop.critic.set_weights(a, per.critics[p][m].get_weights())
            tr.pop.t_actor.set_weights(a, per.t_actors[p][m].get_weights())
            tr.pop.t_critic.set_weights(a, per.t_critics[p][m].get_weights())
    tr.env.auto_reset = False
    tr.env.reset()                         # the second reset of every platoon, like Platoon() + env.reset() at the episode start
    per.run(1)
    for _ in range(STEPS):
        tr.step()
    torch.cuda.synchronize()
    # same environment trajectory ...
    for p in range(P):
        got = np.stack([per.envs[p].followers[m].x for m in range(M)])
        np.testing.assert_allclose(got, tr.env.state[p].double().cpu().numpy(), rtol=1e-5, atol=1e-6)
    # ... same number of updates, same networks
    n_learn = STEPS - 16
    assert tr.pop.actor.step.tolist() == [n_learn] * (P * M)
    worst = 0.0
    for p in range(P):
        for m in range(M):
            a = m * P + p
            assert per.actor_opt[p][m].iterations == n_learn and per.critic_opt[p][m].iterations == n_learn
            for bank, obj in ((tr.pop.actor, per.actors[p][m]), (tr.pop.critic, per.critics[p][m]), (tr.pop.t_actor, per.t_actors[p][m]),
                              (tr.pop.t_critic, per.t_critics[p][m])):
                for name, w in zip(bank.weight_names, obj.get_weights()):
                    got = bank.view(name, a).cpu().numpy()
                    np.testing.assert_allclose(got, w.reshape(got.shape), rtol=2e-4, atol=2e-6, err_msg=f"{fed} ({p},{m}) {bank.kind}.{name}")
                    worst = max(worst, float(np.max(np.abs(got - w.reshape(got.shape)))))
    if fed == "interfrl":                  # platoons of a follower started identical and received identical averaged gradients
        for m in range(M):
            assert all(np.array_equal(x, y) for x, y in zip(per.actors[0][m].get_weights(), per.actors[1][m].get_weights()))
    print(f"per-object loop vs batched trainer, fed={fed}: max |delta weight| {worst:.2e} after {n_learn} updates")
