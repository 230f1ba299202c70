"""Drop-in for the reference's ``agent/ddpgagent.py``: ``policy`` and ``update_target``.

In the batched training loop both are fused into kernels (OU noise + clip inside the environment step,
Polyak inside ``avd_ddpg_learn``); these functions keep the per-object call signatures of
agent/ddpgagent.py:6-29 and 31-55 for code written against the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def policy(actor_state, noise_object=None, lbound=None, hbound=None):
    """clip(squeeze(actor output) [+ noise()], lbound, hbound) -> [scalar]  (ddpgagent.py:18-29)."""
    sampled = actor_state.detach().reshape(-1).double().cpu().numpy() if torch.is_tensor(actor_state) \
        else np.asarray(actor_state, dtype=np.float64).reshape(-1)
    sampled = np.squeeze(sampled)
    if noise_object is not None:
        sampled = sampled + noise_object()
    return [np.squeeze(np.clip(sampled, lbound, hbound))]


def _polyak_list(tau, targets, onlines):
    lib = _lib.load()
    out = []
    for t, o in zip(targets, onlines):
        if not (torch.is_tensor(t) and t.is_cuda):
            t = torch.as_tensor(np.asarray(t, dtype=np.float32), device="cuda")
        if not (torch.is_tensor(o) and o.is_cuda):
            o = torch.as_tensor(np.asarray(o, dtype=np.float32), device="cuda")
        new = t.detach().clone().contiguous()
        oc = o.detach().contiguous()
        _lib.check(lib.avd_polyak_update(_lib.ptr(new), _lib.ptr(oc), None, 1, new.numel(), float(tau), _lib.current_stream()))
        out.append(new)
    return out


def update_target(tau, t_critic_weights, critic_weights, t_actor_weights, actor_weights):
    """-> (tc_new_weights, ta_new_weights): theta' = tau*theta + (1-tau)*theta' per tensor, over ALL weights incl.
    BatchNorm moving statistics (ddpgagent.py:44-55).  Pure function: inputs are not modified."""
    _lib.require_device()
    return _polyak_list(tau, t_critic_weights, critic_weights), _polyak_list(tau, t_actor_weights, actor_weights)
