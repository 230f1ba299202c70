"""Generate tests/golden/evaluator.npz (container only; needs /root/reference): the reference's evaluator loop
(workers/evaluator.py:40-96,145) -- seed 6, Platoon(evaluator_states_enabled=True), pre-drawn N(0, reset_max_u) leader
inputs, ddpgagent.policy without noise, float32 episodic reward counters, round(mean, 3) -- with the oracle's actor
forward standing in for the Keras actors (TensorFlow is not installable here).
    python tools/make_evaluator_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, ddpg_np as D

ref = ref_import.load()
import tensorflow as tf   # the NumPy-backed shim

conf = ref.config.Config(); conf.pl_size = 3
np.random.seed(conf.evaluation_seed)
pl = ref_import.make_platoon(ref, 3, conf, 1, evaluator_states_enabled=True)
T = 100
inputs = [ref.util.get_random_val(conf.rand_gen, conf.reset_max_u, std_dev=conf.reset_max_u, config=conf) for _ in range(T)]
rng = np.random.default_rng(99)
actors = []
for m in range(3):
    a = D.init_actor(rng); D.randomize_bn(a, rng, [("g1", "be1", "mu1", "var1"), ("g2", "be2", "mu2", "var2")]); a["W3"] *= 100
    actors.append(a)
states = pl.reset()
ep = np.array([0] * 3, dtype=np.float32)
S = np.zeros((T, 3, 4)); U = np.zeros((T, 3)); J = np.zeros((T, 3))
acts = np.zeros((3, 1))
for i in range(T):
    for m in range(3):
        out, _ = D.actor_forward(actors[m], np.asarray(states[m], np.float32)[None])
        acts[m] = ref.ddpgagent.policy(tf.convert_to_tensor(out), lbound=conf.action_low, hbound=conf.action_high)[0]
    states, rewards, term = pl.step(acts.flatten(), inputs[i], False)
    J[i] = np.reshape(pl.get_jerk(), 3); S[i] = np.stack(states); U[i] = acts.flatten()
    for m in range(3):
        ep[m] += rewards[m]
d = {"inputs": np.array(inputs), "states": S, "actions": U, "jerks": J, "ep_reward": ep, "pl_rew": np.float64(round(np.average(ep), 3)), "T": np.int64(T)}
for m in range(3):
    for k, v in actors[m].items():
        d[f"actor{m}_{k}"] = v
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "evaluator.npz"), **d)
print("pl_rew", d["pl_rew"], ep)
