"""Import the UNMODIFIED reference modules from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  /root/reference does not exist on
the GPU box, so only ``oracle/make_golden.py`` and the container-only consistency tests
use this module; everything that runs on the GPU box reads ``tests/golden/`` instead.

The reference needs ``tensorflow`` and ``h5py`` at import time (src/util.py:1,3,
src/replaybuffer.py:4, agent/ddpgagent.py:1, src/server/federated.py:4); neither is
installed.  We register a *NumPy-backed shim* that provides exactly the handful of
``tf.*`` calls the hot path makes, so the reference's own arithmetic runs unmodified:

  tf.convert_to_tensor / tf.cast / tf.squeeze / tf.expand_dims   (replaybuffer.py:57-61,
                                                                  ddpgagent.py:18)
  tf.stack / tf.reduce_mean / tf.reduce_sum / tf.math.scalar_mul (federated.py:56-62,
                                                                  109-110; trainer.py:359)

Every shim function is the obvious NumPy equivalent; no reference logic is restated here.
"""
from __future__ import annotations

import contextlib
import io
import logging
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("AVDDPG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "environment.py"))


class _Tensor(np.ndarray):
    """ndarray with the ``.numpy()`` accessor TF eager tensors have."""

    def numpy(self):
        return np.asarray(self)


def _as_tensor(x, dtype=None):
    arr = np.asarray(x, dtype=dtype)
    return arr.view(_Tensor)


def _install_shims() -> None:
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "_avd_shim", False):
        return
    tf = types.ModuleType("tensorflow")
    tf._avd_shim = True
    tf.float32 = np.float32
    tf.float64 = np.float64
    tf.convert_to_tensor = lambda x, dtype=None: _as_tensor(x, dtype)
    tf.cast = lambda x, dtype=None: _as_tensor(np.asarray(x).astype(dtype))
    tf.squeeze = lambda x, axis=None: _as_tensor(np.squeeze(np.asarray(x), axis=axis))
    tf.expand_dims = lambda x, axis: _as_tensor(np.expand_dims(np.asarray(x), axis))
    tf.stack = lambda xs, axis=0: _as_tensor(np.stack([np.asarray(v) for v in xs], axis=axis))
    tf.reduce_mean = lambda x, axis=None: _as_tensor(np.mean(np.asarray(x), axis=axis))
    tf.reduce_sum = lambda x, axis=None: _as_tensor(np.sum(np.asarray(x), axis=axis))
    tf_math = types.ModuleType("tensorflow.math")
    tf_math.scalar_mul = lambda s, x: _as_tensor(np.asarray(s) * np.asarray(x))
    tf_math.reduce_mean = tf.reduce_mean
    tf_math.square = lambda x: _as_tensor(np.square(np.asarray(x)))
    tf.math = tf_math
    # modules imported for side effects only by files we never execute
    keras = types.ModuleType("tensorflow.keras")
    keras.layers = types.ModuleType("tensorflow.keras.layers")
    tf.keras = keras
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.math"] = tf_math
    sys.modules["tensorflow.keras"] = keras
    sys.modules["tensorflow.keras.layers"] = keras.layers
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))


_CACHE = {}


def load():
    """Return a namespace with the reference modules (environment, noise, config, util,
    replaybuffer, ddpgagent, federated)."""
    if _CACHE:
        return _CACHE["ns"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    server_dir = os.path.join(REFERENCE_ROOT, "src", "server")
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src import config, environment, noise, replaybuffer, util  # type: ignore
        from agent import ddpgagent  # type: ignore
        from src.server import federated  # type: ignore
    for name in ("src.environment", "src.util", "src.server.federated"):
        logging.getLogger(name).setLevel(logging.CRITICAL)
    ns = types.SimpleNamespace(config=config, environment=environment, noise=noise,
                               replaybuffer=replaybuffer, util=util, ddpgagent=ddpgagent,
                               federated=federated, server_dir=server_dir)
    _CACHE["ns"] = ns
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints from Vehicle.set_system_matrices (environment.py:394)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def make_platoon(ref, length, conf, pl_idx=0, **kw):
    """Construct a reference Platoon, bypassing the 6-colour rendering guard
    (environment.py:84-85) for length > 6: the guard only protects the GUI, so for
    longer platoons we build with 6 followers' worth of colours by temporarily
    constructing at the requested length and swallowing that one ValueError after the
    followers list (the only state the hot path uses) has been fully built."""
    with quiet():
        if length <= 6:
            return ref.environment.Platoon(length, conf, pl_idx, **kw)
        pl = ref.environment.Platoon.__new__(ref.environment.Platoon)
        try:
            ref.environment.Platoon.__init__(pl, length, conf, pl_idx, **kw)
        except ValueError:
            pass  # raised by the colour guard *after* followers/attributes are set
        assert len(pl.followers) == length
        return pl
