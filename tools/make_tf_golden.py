#!/usr/bin/env python
"""Pin the learn step / Adam against REAL TensorFlow: writes tests/golden/tf_learn.npz.

TensorFlow 2.4.1 (the reference's requirements.txt:2) cannot be installed in the build container (no network, not vendored), so the
learn / Adam oracle (oracle/ddpg_np.py) is "parity unpinned" there.  This script is what a maintainer runs ONCE wherever the
reference's environment exists:

    cd /path/to/avddpg                      # the reference checkout (agent/, workers/, src/ importable)
    pip install tensorflow==2.4.1 numpy==1.19.5
    python /path/to/this/repo/tools/make_tf_golden.py --reference . --out /path/to/this/repo/tests/golden/tf_learn.npz

It builds the reference's own Keras models (agent/model.py get_actor / get_critic), overwrites their weights with seeded NumPy
values (so nothing depends on TF's initialiser bit stream), runs the reference's own Trainer.learn arithmetic
(workers/trainer.py:489-506, copied call for call below because Trainer.__init__ needs a whole experiment directory), applies
tf.keras.optimizers.Adam exactly as workers/trainer.py:138-139, 348-349 do, for three consecutive steps, and
ddpgagent.update_target (agent/ddpgagent.py:31-55).  Everything a consumer needs is stored: initial weights (Keras `.weights`
order), the batch, per-step gradients (Keras `trainable_variables` order), losses, and the weights after each step.

tests/test_tf_golden.py consumes the file when it exists: the NumPy oracle must reproduce it (CPU), and the CUDA path must (GPU).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of cboin1996/avddpg")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tf_learn.npz"))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    sys.path.insert(0, os.path.abspath(args.reference))
    import tensorflow as tf
    from agent import ddpgagent, model
    from src import config

    conf = config.Config()
    rng = np.random.default_rng(args.seed)
    ns, na, high = 4, 1, conf.action_high
    mk_actor = lambda: model.get_actor(ns, na, high, seed_int=1, hidd_mult=1, layer1_size=conf.actor_layer1_size, layer2_size=conf.actor_layer2_size)
    mk_critic = lambda: model.get_critic(ns, na, hidd_mult=1, layer1_size=conf.critic_layer1_size, action_layer_size=conf.critic_act_layer_size,
                                         layer2_size=conf.critic_layer2_size)
    actor, critic, t_actor, t_critic = mk_actor(), mk_critic(), mk_actor(), mk_critic()

    def randomise(net, head_scale):
        """Seeded weights with non-trivial BatchNormalization parameters and statistics and non-zero biases."""
        new = []
        for w in net.weights:
            shape, name = tuple(w.shape), w.name
            if "moving_variance" in name or "gamma" in name:
                v = rng.uniform(0.5, 1.5, shape)
            elif "moving_mean" in name:
                v = rng.uniform(-0.1, 0.1, shape)
            elif "beta" in name:
                v = rng.uniform(-0.2, 0.2, shape)
            elif len(shape) == 2:
                bound = 1.0 / np.sqrt(shape[1]) if shape[1] > 1 else head_scale
                v = rng.uniform(-bound, bound, shape)
            else:
                v = rng.normal(0, 0.05, shape)
            new.append(v.astype(np.float32))
        net.set_weights(new)

    randomise(actor, 0.15)
    randomise(critic, 0.09)
    t_actor.set_weights([w + rng.normal(0, 0.01, w.shape).astype(np.float32) for w in actor.get_weights()])
    t_critic.set_weights([w + rng.normal(0, 0.01, w.shape).astype(np.float32) for w in critic.get_weights()])

    B = args.batch
    s = rng.normal(0, 2, (B, ns)).astype(np.float32)
    a = rng.uniform(-high, high, (B, na)).astype(np.float32)
    r = -rng.uniform(0, 0.5, (B, 1)).astype(np.float32)
    s2 = (s + rng.normal(0, 0.2, (B, ns))).astype(np.float32)
    out = {"batch_s": s, "batch_a": a, "batch_r": r, "batch_s2": s2, "gamma": np.float64(conf.gamma), "tau": np.float64(conf.tau),
           "actor_lr": np.float64(conf.actor_lr), "critic_lr": np.float64(conf.critic_lr), "high": np.float64(high),
           "tf_version": np.array(tf.__version__), "steps": np.int64(args.steps)}
    for tag, net in (("actor", actor), ("critic", critic), ("t_actor", t_actor), ("t_critic", t_critic)):
        out[f"{tag}_weight_names"] = np.array([w.name for w in net.weights])
        out[f"{tag}_trainable_names"] = np.array([w.name for w in net.trainable_variables])
        for i, w in enumerate(net.get_weights()):
            out[f"init_{tag}_{i:02d}"] = w

    critic_opt = tf.keras.optimizers.Adam(conf.critic_lr)          # workers/trainer.py:138-139
    actor_opt = tf.keras.optimizers.Adam(conf.actor_lr)
    ts, ta, tr_, ts2 = (tf.convert_to_tensor(x) for x in (s, a, r, s2))
    for k in range(args.steps):
        with tf.GradientTape() as tape:                              # workers/trainer.py:491-496
            target_actions = t_actor(ts2)
            y = tr_ + conf.gamma * t_critic([ts2, target_actions])
            critic_value = critic([ts, ta])
            critic_loss = tf.math.reduce_mean(tf.math.square(y - critic_value))
        critic_grad = tape.gradient(critic_loss, critic.trainable_variables)      # :498
        with tf.GradientTape() as tape:                              # :501-504
            actions = actor(ts)
            critic_value = critic([ts, actions])
            actor_loss = -tf.math.reduce_mean(critic_value)
        actor_grad = tape.gradient(actor_loss, actor.trainable_variables)         # :506
        out[f"step{k}_critic_loss"], out[f"step{k}_actor_loss"] = np.float64(critic_loss.numpy()), np.float64(actor_loss.numpy())
        for i, g in enumerate(critic_grad):
            out[f"step{k}_critic_grad_{i:02d}"] = g.numpy()
        for i, g in enumerate(actor_grad):
            out[f"step{k}_actor_grad_{i:02d}"] = g.numpy()
        critic_opt.apply_gradients(zip(critic_grad, critic.trainable_variables))   # :348-349
        actor_opt.apply_gradients(zip(actor_grad, actor.trainable_variables))
        tc_new, ta_new = ddpgagent.update_target(conf.tau, t_critic.weights, critic.weights, t_actor.weights, actor.weights)      # :352-354
        t_actor.set_weights(ta_new)                                  # :355-356
        t_critic.set_weights(tc_new)
        for tag, net in (("actor", actor), ("critic", critic), ("t_actor", t_actor), ("t_critic", t_critic)):
            for i, w in enumerate(net.get_weights()):
                out[f"step{k}_{tag}_{i:02d}"] = w
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: TF {tf.__version__}, batch {B}, {args.steps} learn + Adam + Polyak steps")


if __name__ == "__main__":
    main()
