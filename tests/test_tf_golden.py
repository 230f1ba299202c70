"""Consumers of tests/golden/tf_learn.npz -- the fixture tools/make_tf_golden.py writes from REAL TensorFlow 2.4.1 (the
reference's learn step, tf.keras Adam and update_target on the reference's own Keras models).  TensorFlow is not installable in
the build container, so the file may be absent; then the TF-pinned tests skip and say so, and the learn / Adam oracle stays
"parity unpinned" (DESIGN.md section 5).  The schema and the consumers themselves are always exercised through a synthetic file
written from the oracle (which pins nothing).

When the file is present:
  CPU  oracle/ddpg_np.py must reproduce TensorFlow's gradients (1e-4 rel-L2: fp32 summation order), losses (1e-5) and the weights
       after three learn + Adam + Polyak steps (1e-5)
  GPU  the CUDA path (precision 0) must, at the tolerances of tests/test_gpu_ddpg.py (gradients 2e-4 normwise)
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ddpg_np as D
from oracle import tf_golden

TF_FILE = os.path.join(GOLDEN, "tf_learn.npz")
needs_tf_file = pytest.mark.skipif(not os.path.exists(TF_FILE), reason="tests/golden/tf_learn.npz absent: run tools/make_tf_golden.py where "
                                   "TensorFlow 2.4.1 is installed (learn / Adam oracle stays parity-unpinned until then)")


def test_consumer_on_synthetic_file(tmp_path):
    """The reader maps the Keras-ordered lists onto the oracle's named tensors, checks every shape and replays the steps."""
    g = tf_golden.load(tf_golden.write_synthetic(str(tmp_path / "synthetic.npz")))
    assert g["tf_version"] == "oracle-synthetic" and len(g["steps"]) == 3
    assert set(g["init"]) == {"actor", "critic", "t_actor", "t_critic"}
    assert list(g["init"]["critic"]) == D.CRITIC_WEIGHTS and list(g["steps"][0]["actor_grad"]) == D.ACTOR_TRAINABLE
    worst = tf_golden.replay_with_oracle(g)
    assert worst["grad"] < 1e-6 and worst["loss"] < 1e-6 and worst["weights"] < 1e-6, worst


def test_consumer_rejects_a_different_keras_order(tmp_path):
    path = tf_golden.write_synthetic(str(tmp_path / "synthetic.npz"))
    with np.load(path) as f:
        z = {k: f[k] for k in f.files}
    z["init_critic_02"], z["init_critic_04"] = z["init_critic_04"], z["init_critic_02"]      # Wa <-> gs: shapes (1, 48) vs (256,)
    np.savez(str(tmp_path / "bad.npz"), **z)
    with pytest.raises(ValueError, match="Keras order"):
        tf_golden.load(str(tmp_path / "bad.npz"))


@needs_tf_file
def test_oracle_vs_tensorflow():
    g = tf_golden.load(TF_FILE)
    assert g["tf_version"] != "oracle-synthetic", "tests/golden/tf_learn.npz must come from tools/make_tf_golden.py (real TensorFlow)"
    worst = tf_golden.replay_with_oracle(g)
    assert worst["grad"] < 1e-4 and worst["loss"] < 1e-5 and worst["weights"] < 1e-5, worst


def _cuda_vs_file(path, grad_tol, w_tol):
    import torch
    from avddpg_b200.config import Config
    from avddpg_b200.trainer import DDPGPopulation
    g = tf_golden.load(path)
    h = g["hyper"]
    conf = Config(gamma=h["gamma"], tau=h["tau"], actor_lr=h["actor_lr"], critic_lr=h["critic_lr"], action_high=h["high"], action_low=-h["high"])
    s, a, r, s2 = g["batch"]
    pop = DDPGPopulation(1, 1, conf, rows_per_agent=len(s), precision=0)
    for tag, bank in (("actor", pop.actor), ("critic", pop.critic), ("t_actor", pop.t_actor), ("t_critic", pop.t_critic)):
        bank.load_named(0, g["init"][tag])
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    ts, ta, tr, ts2 = dev(s), dev(a.reshape(-1)), dev(r.reshape(-1)), dev(s2)
    for st in g["steps"]:
        pop.learn(ts, ta, tr, ts2, apply_updates=True)
        for bank, ref in ((pop.critic, st["critic_grad"]), (pop.actor, st["actor_grad"])):
            for name in bank.trainable_names:
                got = bank.view(name, 0, bank.grad).cpu().numpy()
                want = ref[name].reshape(got.shape)
                assert np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-12) < grad_tol, (bank.kind, name)
        for tag, bank in (("actor", pop.actor), ("critic", pop.critic), ("t_actor", pop.t_actor), ("t_critic", pop.t_critic)):
            for name in bank.weight_names:
                got = bank.view(name, 0).cpu().numpy()
                np.testing.assert_allclose(got, st["weights"][tag][name].reshape(got.shape), rtol=w_tol, atol=w_tol * 1e-2, err_msg=f"{tag}.{name}")


@pytest.mark.gpu
def test_cuda_vs_synthetic_file(tmp_path):
    """The GPU consumer runs end to end on the synthetic (oracle-written) file: same bars as tests/test_gpu_ddpg.py."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    _cuda_vs_file(tf_golden.write_synthetic(str(tmp_path / "synthetic.npz")), 2e-4, 2e-4)


@pytest.mark.gpu
@needs_tf_file
def test_cuda_vs_tensorflow():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    _cuda_vs_file(TF_FILE, 2e-4, 2e-4)
