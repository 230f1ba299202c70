"""world_size-2 `gloo` test of the FRL exchange step (host logic only; no GPU): platoons sharded over ranks,
local pre-reduction, ONE all_reduce(sum), scaling -- compared against the oracle FedAvg over the global
member set (oracle/ddpg_np.py, itself pinned to the reference's Server via tests/golden/fedavg.npz)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avddpg_b200.server.federated import exchange_and_scale, shard_platoons
from oracle import ddpg_np as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, P, M, n, weighted, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    grads = rs.normal(size=(P, M, n)).astype(np.float32)          # every rank knows the global truth for checking
    w = rs.uniform(0.1, 3.0, size=(P, M)).astype(np.float32) if weighted else np.ones((P, M), np.float32)
    lo, hi = shard_platoons(P, rank, world)
    buf = torch.zeros(M, n + 1)
    for m in range(M):                                             # interfrl: system = follower m, members = platoons
        buf[m, :n] = torch.as_tensor((w[lo:hi, m, None] * grads[lo:hi, m]).sum(0))
        buf[m, n] = float(w[lo:hi, m].sum()) if weighted else float(hi - lo)
    exchange_and_scale(buf, dist.group.WORLD)
    out_q.put((rank, buf.numpy().copy(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def _run(P, M, n, weighted):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, M, n, weighted, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def _truth(P, M, n, weighted):
    rs = np.random.RandomState(0)
    grads = rs.normal(size=(P, M, n)).astype(np.float32)
    w = rs.uniform(0.1, 3.0, size=(P, M)).astype(np.float32) if weighted else None
    if weighted:
        sys_params = [[[w[p, m] * grads[p, m]] for p in range(P)] for m in range(M)]
        return D.fed_weighted_average(sys_params, w.sum(0))
    return D.fed_average([[[grads[p, m]] for p in range(P)] for m in range(M)])


def test_shard_platoons_partition():
    for P in (1, 7, 8, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_platoons(P, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_interfrl_exchange_unweighted_world2():
    P, M, n = 7, 3, 257          # ragged shard: 4 + 3 platoons
    res = _run(P, M, n, False)
    truth = _truth(P, M, n, False)
    spans = sorted(r[2] for r in res)
    assert spans == [(0, 4), (4, 7)]
    for _, buf, _ in res:
        for m in range(M):
            np.testing.assert_allclose(buf[m, :n], truth[m][0], rtol=1e-5, atol=1e-6)
            assert buf[m, n] == P
    np.testing.assert_array_equal(res[0][1], res[1][1])   # both ranks end with identical averages


def test_interfrl_exchange_weighted_world2():
    P, M, n = 8, 4, 100
    res = _run(P, M, n, True)
    truth = _truth(P, M, n, True)
    for _, buf, _ in res:
        for m in range(M):
            np.testing.assert_allclose(buf[m, :n], truth[m][0], rtol=2e-5, atol=1e-6)
