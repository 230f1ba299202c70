"""Federated averaging -- the reference's ``src/server/federated.py`` + the FRL glue of
``workers/trainer.py:361-456`` -- on the GPU, with NCCL for the cross-GPU exchange.

``Server`` is the drop-in: ``get_avg_params`` / ``get_weighted_avg_params`` over nested lists
``system_params[system][member][layer]`` (federated.py:18-122), each layer reduced by one kernel launch.

``FederatedAggregator`` is what the batched trainer uses: the actor/critic gradients (or weights) of a
``DDPGPopulation`` are reduced over the members of each system on this GPU (``avd_fed_reduce``), exchanged
with ONE ``all_reduce(sum)`` over NVLink on a flat fp32 buffer ``[systems x (actor ‖ critic ‖ sum_w)]`` when a
process group spans several GPUs, scaled by 1/P or 1/sum(w), and written back to every member
(``avd_fed_broadcast``), after which each member applies Adam / Polyak (gradients mode, trainer.py:400-431) or
takes the averaged weights for online AND target nets (weights mode, trainer.py:433-456).

Systems: interfrl = follower index m averaged over platoons (trainer.py:338-339, 419-421) -- the one
cross-GPU exchange, since platoons are sharded over ranks; intrafrl = platoon averaged over its followers
(trainer.py:341-342, 423-425) -- purely local.  Agent index a = m*G + g.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from .. import _lib


def _as_cuda(x):
    if torch.is_tensor(x):
        return x.detach().to("cuda", torch.float32)
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device="cuda")


class Server:
    def __init__(self, name, debug_enabled=False):
        self.name, self.debug = name, debug_enabled
        self.lib = _lib.load()
        _lib.require_device()

    def _reduce(self, members, scale):
        """members: list of X same-shaped layer tensors -> scale * sum (one avd_fed_reduce launch)."""
        shape = tuple(np.shape(members[0]))
        stack = torch.stack([_as_cuda(m).reshape(-1) for m in members]).contiguous()
        x, n = stack.shape
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        sc = torch.tensor([scale], dtype=torch.float32, device="cuda")
        _lib.check(self.lib.avd_fed_reduce(_lib.ptr(out), n, _lib.ptr(stack), n, 1, x, 0, 1, None, _lib.ptr(sc), n,
                                           _lib.current_stream()))
        return out.reshape(shape)

    def get_avg_params(self, system_params: list):
        """[[mean over members of layer l for l in layers] for each system]  (federated.py:47-63)."""
        out = []
        for members in system_params:
            n_layers = len(members[0])
            out.append([self._reduce([mem[l] for mem in members], 1.0 / len(members)) for l in range(n_layers)])
        return out

    def get_weighted_avg_params(self, system_params: list, weight_sums):
        """Members arrive pre-multiplied by their weights; result = float32(1/sum_w) * sum  (federated.py:99-118)."""
        out = []
        for members, wsum in zip(system_params, weight_sums):
            n_layers = len(members[0])
            scale = float(np.float32(1 / float(wsum)))
            out.append([self._reduce([mem[l] for mem in members], scale) for l in range(n_layers)])
        return out


def exchange_and_scale(buf: torch.Tensor, process_group=None) -> torch.Tensor:
    """buf[systems, n+1]: local (weighted) SUMS with the local member count / weight sum in the last column.
    One all_reduce(sum) over the group (NCCL on GPUs, gloo in the CPU tests), then the division that turns the
    sums into the mean (federated.py:62) or the weighted mean (federated.py:110).  In place."""
    if process_group is not None:
        import torch.distributed as dist
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=process_group)
    if buf.is_cuda:
        _lib.check(_lib.load().avd_fed_finalize(_lib.ptr(buf), buf.shape[1], buf.shape[0], buf.shape[1] - 1, _lib.current_stream()))
    else:   # host tensors: only the gloo unit test of the exchange logic
        buf[:, :-1] *= (1.0 / buf[:, -1]).unsqueeze(1)
    return buf


def shard_platoons(num_platoons: int, rank: int, world: int):
    """Contiguous platoon range of a rank: [lo, hi)."""
    base, rem = divmod(num_platoons, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FederatedAggregator:
    def __init__(self, population, conf, process_group=None, world_size: Optional[int] = None,
                 reference_weights_quirk: bool = True):
        """population: DDPGPopulation (A = G*M agents of THIS rank).  process_group: a torch.distributed group
        (NCCL on GPUs) or None for single-process.  reference_weights_quirk: weights mode applies system 0's
        average to every agent, as workers/trainer.py:442-456 does (`[...][0]`)."""
        self.pop, self.conf, self.pg = population, conf, process_group
        self.lib = _lib.load()
        if conf.fed_method not in ("interfrl", "intrafrl"):
            raise ValueError(f"fed_method {conf.fed_method!r} is not a federated method")
        self.inter = conf.fed_method == "interfrl"
        self.quirk = reference_weights_quirk
        self.world = 1
        if process_group is not None or (world_size or 1) > 1:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        G, M = population.G, population.M
        # member row of (system s, member x) = s*stride_s + x*stride_x in the [A, n] banks (a = m*G + g)
        if self.inter:
            self.n_systems, self.n_members, self.stride_s, self.stride_x = M, G, G, 1
        else:
            self.n_systems, self.n_members, self.stride_s, self.stride_x = G, M, 1, G
        dev = population.device
        self._bufs = {}
        self.device = dev
        self._ones = torch.ones(self.n_systems, dtype=torch.float32, device=dev)
        self._count = torch.full((self.n_systems,), float(self.n_members), dtype=torch.float32, device=dev)
        self.apply_mask = None
        if (not self.inter) and getattr(conf, "intra_directional_averaging", False):
            mask = torch.ones(G * M, dtype=torch.uint8, device=dev)
            mask[:G] = 0            # follower m = 0 ("leader is king", trainer.py:417-418, 450-451)
            self.apply_mask = mask
        self.rounds = 0

    def _buffer(self, na, nc):
        key = (na, nc)
        if key not in self._bufs:
            self._bufs[key] = torch.zeros(self.n_systems, na + nc + 1, dtype=torch.float32, device=self.device)
        return self._bufs[key]

    def _reduce_exchange(self, a_src, na, c_src, nc, a_pitch, c_pitch, weights):
        """-> buffer [systems, na+nc+1] holding the (weighted) means of actor / critic vectors."""
        S, X = self.n_systems, self.n_members
        buf = self._buffer(na, nc)
        pitch = buf.shape[1]
        st = _lib.current_stream()
        ones = self._ones
        w = None
        if weights is not None:
            w = torch.as_tensor(weights, dtype=torch.float32, device=self.device).reshape(S, X).contiguous()
        for src, n, src_pitch, off in ((a_src, na, a_pitch, 0), (c_src, nc, c_pitch, na)):
            out = buf[:, off:]
            _lib.check(self.lib.avd_fed_reduce(_lib.ptr(out), pitch, _lib.ptr(src), src_pitch, S, X, self.stride_s, self.stride_x,
                                               _lib.ptr(w), _lib.ptr(ones), n, st))
        buf[:, na + nc].copy_(w.sum(dim=1) if w is not None else self._count)
        exchange_and_scale(buf, self.pg if (self.inter and self.world > 1) else None)
        self.rounds += 1
        return buf

    def aggregate_gradients(self, weights=None, apply: bool = True):
        """train_all_models_federated_gradients (trainer.py:400-431): average actor/critic gradients per system,
        then every member applies them with its own Adam and soft-updates its targets."""
        pop = self.pop
        na, nc = pop.actor.n_train, pop.critic.n_train
        buf = self._reduce_exchange(pop.actor.grad, na, pop.critic.grad, nc, na, nc, weights)
        pitch = buf.shape[1]
        st = _lib.current_stream()
        for bank, n, off in ((pop.actor, na, 0), (pop.critic, nc, na)):
            src = buf[:, off:]
            _lib.check(self.lib.avd_fed_broadcast(_lib.ptr(bank.grad), n, _lib.ptr(src), pitch, self.n_systems, self.n_members,
                                                  self.stride_s, self.stride_x, _lib.ptr(self.apply_mask), n, st))
        if apply:
            pop.apply_gradients(self.apply_mask)
            pop.soft_update(self.apply_mask)
        return buf

    def aggregate_weights(self, weights=None):
        """train_all_models_federated_weights (trainer.py:433-456): average `.weights` (incl. BN statistics) and
        set them on the online AND target nets of every member."""
        pop = self.pop
        na, nc = pop.actor.total, pop.critic.total
        buf = self._reduce_exchange(pop.actor.flat, na, pop.critic.flat, nc, na, nc, weights)
        pitch = buf.shape[1]
        st = _lib.current_stream()
        for banks, n, off in (((pop.actor, pop.t_actor), na, 0), ((pop.critic, pop.t_critic), nc, na)):
            src = buf[:, off:]
            for bank in banks:
                if self.quirk:   # every agent receives system 0's average
                    _lib.check(self.lib.avd_fed_broadcast(_lib.ptr(bank.flat), n, _lib.ptr(src), pitch, 1, pop.A, 0, 1,
                                                          _lib.ptr(self.apply_mask), n, st))
                else:
                    _lib.check(self.lib.avd_fed_broadcast(_lib.ptr(bank.flat), n, _lib.ptr(src), pitch, self.n_systems,
                                                          self.n_members, self.stride_s, self.stride_x,
                                                          _lib.ptr(self.apply_mask), n, st))
        return buf
