"""Drop-in for the reference's ``src/noise.py`` (OUActionNoise, noise.py:3-29) on the GPU.

In the batched training path the OU update is fused into the environment-step kernel
(csrc/avd_env.cu); this class is the stand-alone, per-object form with the reference's signature.  Draws
come from the Philox stream (AVD_RNG_OU, stream_id, call counter) instead of NumPy's global MT19937.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .config import env_params_from_config


class OUActionNoise:
    def __init__(self, mean, x_init=None, config=None, *, stream_id: int = 0, seed=None):
        self.config = config
        self.theta, self.dt = config.theta, config.ou_dt
        self.mean = np.asarray(mean, dtype=np.float64)
        self.std_dev = float(config.std_dev) * np.ones(1)
        self.x_init = x_init
        self._lib = _lib.load()
        _lib.require_device()
        self._prm = env_params_from_config(config, 1)
        self._prm.ou_mean = float(self.mean.reshape(-1)[0]) if self.mean.size else 0.0
        self._seed = int(getattr(config, "random_seed", 1) if seed is None else seed)
        self._id, self._tick = int(stream_id), 0
        self._n = max(1, self.mean.size)
        self._state = torch.zeros(self._n, dtype=torch.float32, device="cuda")
        self.reset()

    def reset(self):
        if self.x_init is not None:
            self._state.copy_(torch.as_tensor(np.asarray(self.x_init, dtype=np.float32).reshape(-1)))
        else:
            self._state.zero_()

    @property
    def x_prev(self):
        return self._state.double().cpu().numpy().reshape(self.mean.shape)

    def __call__(self):
        _lib.check(self._lib.avd_ou_sample(C.byref(self._prm), _lib.ptr(self._state), None, self._n,
                                           self._id * self._n, self._seed, self._tick, _lib.current_stream()))
        self._tick += 1
        return self.x_prev
