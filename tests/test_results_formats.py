"""On-disk formats: avddpg_b200/results.py against files written by the reference's own code
(tools/make_results_golden.py -> tests/golden/results_*): reward / FRL-weight CSVs byte for byte, conf.json round trip."""
import json
import os
from types import SimpleNamespace

import numpy as np

from avddpg_b200 import results
from avddpg_b200.config import Config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _log(tmp_path):
    z = np.load(os.path.join(GOLD, "results_inputs.npz"))
    P, M, EP, seed, window, wwin = (int(v) for v in z["meta"])
    conf = Config(random_seed=seed, reward_averaging_window=window, weighted_window=wwin, weighted_average_enabled=True)
    log = results.RewardLog(conf, P, M)
    for ep in range(EP):
        if ep >= wwin:
            log.update_reward_list(z["ep_rewards"][ep], z["fed_w"][ep], z["fed_ws"][ep])
        else:
            log.update_reward_list(z["ep_rewards"][ep])
    return conf, log


def test_reward_and_weight_csvs_match_reference_bytes(tmp_path):
    conf, log = _log(tmp_path)
    paths = log.generate_csvs(str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["avg_ep_reward__seed7.csv", "ep_reward__seed7.csv", "frl_weightings__seed7.csv"]
    for got, ref in zip(paths, ("results_avg_ep_reward.csv", "results_ep_reward.csv", "results_frl_weightings.csv")):
        assert open(got).read() == open(os.path.join(GOLD, ref)).read(), ref


def test_windowed_average_semantics():
    conf = Config(reward_averaging_window=3)
    log = results.RewardLog(conf, 1, 1)
    for v in (1.0, 2.0, 6.0, 10.0):
        log.update_reward_list([[np.float32(v)]])
    assert [float(x) for x in log.all_avg_reward_lists[0][0]] == [1.0, 1.5, 3.0, 6.0]      # trainer.py:514-515


def test_conf_json_round_trip_and_reference_file(tmp_path):
    conf = Config(pl_size=4, fed_method="interfrl")
    p = str(tmp_path / "conf.json")
    results.config_writer(p, conf)
    back = results.config_loader(p)
    assert isinstance(back, SimpleNamespace) and back.pl_size == 4 and back.fed_method == "interfrl"
    assert json.load(open(p)) == {k: v for k, v in conf.__dict__.items()}
    ref = results.config_loader(os.path.join(GOLD, "results_conf.json"))      # written by the reference's util.config_writer
    mine = Config(random_seed=7, weighted_window=4, reward_averaging_window=5)     # the three fields the golden script changed
    missing = [k for k in ("pl_size", "batch_size", "buffer_size", "gamma", "tau", "actor_lr", "critic_lr", "timegap", "dyn_coeff",
                           "reward_ep_coeff", "max_ep", "max_ev", "action_high", "std_dev", "theta", "ou_dt", "fed_method",
                           "aggregation_method", "weighted_window", "random_seed") if getattr(ref, k) != getattr(mine, k)]
    assert not missing, f"defaults differ from the reference's conf.json: {missing}"


def test_weights_npz_keeps_keras_order(tmp_path):
    ws = [np.full((2, 3), i, np.float32) for i in range(12)]
    p = str(tmp_path / "actor.npz")
    results.save_weights(p, ws, names=["dense/kernel", "dense/bias"])
    back = results.load_weights(p)
    assert len(back) == 12 and all(np.array_equal(a, b) for a, b in zip(ws, back))
