"""On-disk formats of a training run, compatible with the reference's own tooling.

* ``config_writer`` / ``config_loader``            -- ``conf.json`` (src/util.py:28-35)
* ``RewardLog``                                    -- episodic / window-averaged reward lists (workers/trainer.py:510-517) and the
                                                      ``ep_reward__seed%s.csv`` / ``avg_ep_reward__seed%s.csv`` /
                                                      ``frl_weightings__seed%s.csv`` files (trainer.py:552-629; columns src/env/env.py:5-11),
                                                      so that ``accumr`` / ``lmany`` of the reference can read a B200 run
* ``save_weights`` / ``load_weights``              -- one ``.npz`` per network with the tensors in Keras ``.weights`` order (what
                                                      ``model.save`` writes as ``.h5`` at trainer.py:581-594; h5py / Keras are not
                                                      available here -- INTEGRATION.md shows the three-line conversion on the reference side)

Pure host code: nothing here touches the device.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import List, Sequence

import numpy as np

# column labels of the reference's data frames (src/env/env.py:5-11)
PLATOON_COL = "platoon"
SEED_COL = "seed"
EPISODIC_REWARD_AVGWINDOW_COL = "avg window"
VEHICLE_COL = "Vehicle %s"
TRAINING_EPISODE_COLNAME = "Episode"
FED_WEIGHT_SUM_COL = "Vehicle %s fws"
FED_WEIGHT_PCT_COL = "Vehicle %s pct"


def config_writer(fpath: str, obj) -> None:
    """json.dump(conf.__dict__) -- src/util.py:28-31."""
    with open(fpath, "w") as f:
        json.dump(obj.__dict__, f)


def config_loader(fpath: str):
    """json -> SimpleNamespace (nested dicts too) -- src/util.py:33-35."""
    with open(fpath, "r") as f:
        return json.load(f, object_hook=lambda d: SimpleNamespace(**d))


class RewardLog:
    """all_ep_reward_lists / all_avg_reward_lists of the reference trainer, indexed [platoon][model]."""

    def __init__(self, conf, num_platoons: int, num_models: int):
        self.conf, self.P, self.M = conf, int(num_platoons), int(num_models)
        self.all_ep_reward_lists: List[List[list]] = [[[] for _ in range(self.M)] for _ in range(self.P)]
        self.all_avg_reward_lists: List[List[list]] = [[[] for _ in range(self.M)] for _ in range(self.P)]
        self.all_fed_weights: List[List[list]] = [[[] for _ in range(self.M)] for _ in range(self.P)]
        self.all_fed_weight_sums: List[List[list]] = [[[] for _ in range(self.M)] for _ in range(self.P)]

    def update_reward_list(self, episodic_rewards, fed_weights=None, fed_weight_sums=None) -> None:
        """End of an episode (trainer.py:510-531): episodic_rewards[p][m] is the cumulative reward counter of agent (p, m)
        (np.float32 in the reference, trainer.py:249); fed_weights / fed_weight_sums [p][m] are recorded when given."""
        w = int(self.conf.reward_averaging_window)
        for p in range(self.P):
            for m in range(self.M):
                self.all_ep_reward_lists[p][m].append(episodic_rewards[p][m])
                self.all_avg_reward_lists[p][m].append(np.mean(self.all_ep_reward_lists[p][m][-w:]))
                if fed_weights is not None:
                    self.all_fed_weights[p][m].append(fed_weights[p][m])
                    self.all_fed_weight_sums[p][m].append(fed_weight_sums[p][m])

    # ---- data frames (trainer.py:596-629)
    def generate_reward_data(self, pl_idx: int):
        import pandas as pd
        tag = pl_idx + 1
        cols = [VEHICLE_COL % (m + 1) for m in range(self.M)]
        avg = np.stack(np.array(self.all_avg_reward_lists[pl_idx], dtype=object), axis=1)
        ep = np.stack(np.array(self.all_ep_reward_lists[pl_idx], dtype=object), axis=1)
        avg_df = pd.DataFrame(data=avg, columns=cols)
        avg_df[SEED_COL] = self.conf.random_seed
        avg_df[PLATOON_COL] = tag
        avg_df[EPISODIC_REWARD_AVGWINDOW_COL] = self.conf.reward_averaging_window
        ep_df = pd.DataFrame(data=ep, columns=cols)
        ep_df[SEED_COL] = self.conf.random_seed
        ep_df[PLATOON_COL] = tag
        return avg_df, ep_df

    def generate_frl_weight_data(self, idx: int):
        import pandas as pd
        tag = idx + 1
        vehicle_cols = [VEHICLE_COL % (m + 1) for m in range(self.M)]
        df = pd.DataFrame(data=np.stack(np.array(self.all_fed_weights[idx], dtype=object), axis=1), columns=vehicle_cols)
        df[SEED_COL] = self.conf.random_seed
        df[PLATOON_COL] = tag
        sum_cols = [FED_WEIGHT_SUM_COL % (m + 1) for m in range(self.M)]
        df2 = pd.DataFrame(data=np.stack(np.array(self.all_fed_weight_sums[idx], dtype=object), axis=1), columns=sum_cols)
        df = df.join(df2[sum_cols])
        for i in range(self.M):
            df[FED_WEIGHT_PCT_COL % (i + 1)] = df[vehicle_cols[i]] / df[sum_cols[i]]
        df.index += self.conf.weighted_window          # episodes spent waiting for the weighting window
        return df

    def generate_csvs(self, base_dir: str) -> List[str]:
        """The three CSV files of trainer.py:552-579 (per-platoon frames appended in platoon order, pandas index column kept)."""
        import pandas as pd
        avg_frames, ep_frames = zip(*(self.generate_reward_data(p) for p in range(self.P)))
        seed = self.conf.random_seed
        paths = [os.path.join(base_dir, _path(self.conf, "avg_ep_reward_path", "avg_ep_reward__seed%s.csv") % seed),
                 os.path.join(base_dir, _path(self.conf, "ep_reward_path", "ep_reward__seed%s.csv") % seed)]
        pd.concat(avg_frames).to_csv(paths[0])
        pd.concat(ep_frames).to_csv(paths[1])
        if getattr(self.conf, "weighted_average_enabled", False) and self.all_fed_weights[0][0]:
            paths.append(os.path.join(base_dir, _path(self.conf, "frl_weighted_avg_parameters_path", "frl_weightings__seed%s.csv") % seed))
            pd.concat([self.generate_frl_weight_data(p) for p in range(self.P)]).to_csv(paths[2])
        return paths


def _path(conf, attr: str, default: str) -> str:
    return getattr(conf, attr, default)


def save_weights(path: str, weights: Sequence, names: Sequence[str] = ()) -> None:
    """One .npz with the tensors in Keras `.weights` order; keys "00_<name>", "01_<name>", ... keep that order on load."""
    arrays = [np.asarray(w.detach().cpu().numpy() if hasattr(w, "detach") else w, dtype=np.float32) for w in weights]
    keys = [f"{i:02d}_{names[i] if i < len(names) else 'w'}" for i in range(len(arrays))]
    np.savez(path, **dict(zip(keys, arrays)))


def load_weights(path: str) -> List[np.ndarray]:
    with np.load(path) as z:
        return [z[k] for k in sorted(z.files)]
