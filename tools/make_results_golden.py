"""Golden files for avddpg_b200/results.py, produced by the REFERENCE's own code (container only: needs /root/reference).

Runs workers/trainer.py:Trainer.generate_reward_data / generate_frl_weight_data / update_reward_list arithmetic and
src/util.py:config_writer on synthetic reward lists and writes tests/golden/results_*.{csv,json}.
    python tools/make_results_golden.py
"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
from oracle import ref_import

ref = ref_import.load()
# extra import-time stubs for workers/trainer.py (matplotlib, keras backend); none of them is executed
for name in ("matplotlib", "matplotlib.pyplot", "tensorflow.python", "tensorflow.python.keras", "tensorflow.python.keras.backend"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["tensorflow.python.keras.backend"].dtype = None
sys.modules["tensorflow.python.keras.backend"].gradients = None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
import tensorflow as tf
tf.python = sys.modules["tensorflow.python"]
from workers import trainer as ref_trainer  # type: ignore

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
P, M, EPISODES = 2, 3, 57
conf = ref.config.Config()
conf.random_seed = 7
conf.reward_averaging_window = 5
conf.weighted_window = 4
rng = np.random.default_rng(3)
ep_rewards = -rng.uniform(0, 40, (EPISODES, P, M)).astype(np.float32)
fed_w = rng.uniform(0.01, 2.0, (EPISODES, P, M))
fed_ws = fed_w.sum(axis=1, keepdims=True).repeat(P, axis=1)      # interfrl: weights of a follower index summed over platoons

me = types.SimpleNamespace(conf=conf, num_platoons=P, num_models=M,
                           all_ep_reward_lists=[[[] for _ in range(M)] for _ in range(P)],
                           all_avg_reward_lists=[[[] for _ in range(M)] for _ in range(P)],
                           all_fed_weights=[[[] for _ in range(M)] for _ in range(P)],
                           all_fed_weight_sums=[[[] for _ in range(M)] for _ in range(P)])
for ep in range(EPISODES):          # the list arithmetic of Trainer.update_reward_list (trainer.py:513-516)
    for p in range(P):
        for m in range(M):
            me.all_ep_reward_lists[p][m].append(ep_rewards[ep, p, m])
            me.all_avg_reward_lists[p][m].append(np.mean(me.all_ep_reward_lists[p][m][-conf.reward_averaging_window:]))
            if ep >= conf.weighted_window:
                me.all_fed_weights[p][m].append(fed_w[ep, p, m])
                me.all_fed_weight_sums[p][m].append(fed_ws[ep, p, m])
T = ref_trainer.Trainer
avg_frames, ep_frames, w_frames = [], [], []
for p in range(P):
    a, e = T.generate_reward_data(me, p, me.all_avg_reward_lists[p], me.all_ep_reward_lists[p])
    avg_frames.append(a); ep_frames.append(e)
    w_frames.append(T.generate_frl_weight_data(me, p))
pd.concat(avg_frames).to_csv(os.path.join(OUT, "results_avg_ep_reward.csv"))      # DataFrame.append of trainer.py:566-567 == concat
pd.concat(ep_frames).to_csv(os.path.join(OUT, "results_ep_reward.csv"))
pd.concat(w_frames).to_csv(os.path.join(OUT, "results_frl_weightings.csv"))
np.savez(os.path.join(OUT, "results_inputs.npz"), ep_rewards=ep_rewards, fed_w=fed_w, fed_ws=fed_ws,
         meta=np.array([P, M, EPISODES, conf.random_seed, conf.reward_averaging_window, conf.weighted_window]))
ref.util.config_writer(os.path.join(OUT, "results_conf.json"), conf)
print("wrote", [f for f in os.listdir(OUT) if f.startswith("results_")])
