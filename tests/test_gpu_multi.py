"""Multi-GPU tests (run with at least 2 visible GPUs; skipped otherwise): interfrl aggregation -- over NCCL and over the
NVLink-native peer exchange -- equals the single-process aggregation over the union of the platoons, and sharded platoons
reproduce the single-GPU streams."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from avddpg_b200.config import Config
    from avddpg_b200.environment import BatchedPlatoons
    from avddpg_b200.server.federated import FederatedAggregator
    from avddpg_b200.trainer import DDPGPopulation
    conf = Config(pl_size=2, fed_method="interfrl", weighted_average_enabled=False)
    G_total, M = 2 * world, 2
    G = G_total // world
    pop = DDPGPopulation(G, M, conf)
    gen = torch.Generator(device="cuda").manual_seed(123)
    full_a = torch.randn(M, G_total, pop.actor.n_train, device="cuda", generator=gen)      # same on every rank
    full_c = torch.randn(M, G_total, pop.critic.n_train, device="cuda", generator=gen)
    lo = rank * G
    pop.actor.grad.copy_(full_a[:, lo:lo + G].reshape(M * G, -1))
    pop.critic.grad.copy_(full_c[:, lo:lo + G].reshape(M * G, -1))
    ok_a = ok_c = True
    transports = []
    for transport in ("nccl", "p2p", "peer"):      # NCCL all_reduce + finalize vs. the one-kernel NVLink exchange (peer loads / NVLS)
        agg = FederatedAggregator(pop, conf, process_group=dist.group.WORLD, transport=transport)
        transports.append(agg.transport)
        for rnd in range(5):                # several rounds: epochs advance and the symmetric buffer alternates halves
            scale = 1.0 + rnd
            pop.actor.grad.copy_(scale * full_a[:, lo:lo + G].reshape(M * G, -1))
            pop.critic.grad.copy_(scale * full_c[:, lo:lo + G].reshape(M * G, -1))
            w = None
            want_a, want_c = full_a.mean(1, keepdim=True), full_c.mean(1, keepdim=True)
            if rnd % 2 == 1:                # weighted round: members arrive pre-multiplied by w, result = sum / sum_w (federated.py:99-118)
                wf = 0.5 + torch.arange(M * G_total, device="cuda", dtype=torch.float32).reshape(M, G_total) / 7.0
                w = wf[:, lo:lo + G].contiguous()
                want_a = (wf.unsqueeze(-1) * full_a).sum(1, keepdim=True) / wf.sum(1).reshape(M, 1, 1)
                want_c = (wf.unsqueeze(-1) * full_c).sum(1, keepdim=True) / wf.sum(1).reshape(M, 1, 1)
            agg.aggregate_gradients(weights=w, apply=False)
            ok_a &= torch.allclose(pop.actor.grad.reshape(M, G, -1), (scale * want_a).expand(M, G, -1), rtol=2e-5, atol=1e-6)
            ok_c &= torch.allclose(pop.critic.grad.reshape(M, G, -1), (scale * want_c).expand(M, G, -1), rtol=2e-5, atol=1e-6)
    assert transports[0] == "nccl" and transports[1] == "peer/p2p" and transports[2].startswith("peer/"), transports
    # full rounds with the update applied: the fused consumer of the peer transports (barrier -> in-switch reduction -> Adam ->
    # Polyak in one kernel) ends where all_reduce + finalize + broadcast + Adam/Polyak ends; three rounds, weighted and not
    snap = {k: getattr(pop, k).flat.clone() for k in ("actor", "critic", "t_actor", "t_critic")}
    finals = []
    for transport in ("nccl", "p2p", "peer"):
        for k, v in snap.items():
            getattr(pop, k).flat.copy_(v)
        for bank in (pop.actor, pop.critic):
            bank.m.zero_(); bank.v.zero_(); bank.step.zero_()
        agg = FederatedAggregator(pop, conf, process_group=dist.group.WORLD, transport=transport)
        for rnd in range(3):
            pop.actor.grad.copy_((1.0 + rnd) * full_a[:, lo:lo + G].reshape(M * G, -1))
            pop.critic.grad.copy_((1.0 + rnd) * full_c[:, lo:lo + G].reshape(M * G, -1))
            wf = 0.5 + torch.arange(M * G_total, device="cuda", dtype=torch.float32).reshape(M, G_total) / 7.0
            w = wf[:, lo:lo + G].contiguous() if rnd == 1 else None
            out = agg.aggregate_gradients(weights=w, apply=True, write_back=False)
            assert (out is None) == (transport != "nccl")          # peer transports: fused, nothing materialised
            if w is not None:
                ok_a &= torch.allclose(agg.last_weight_sums, wf.sum(1), rtol=1e-6)
        ok_a &= pop.actor.step.tolist() == [3] * (M * G)
        finals.append({k: getattr(pop, k).flat.clone() for k in snap})
    for other in finals[1:]:
        for k in snap:
            ok_c &= torch.allclose(other[k], finals[0][k], rtol=2e-5, atol=1e-7)
    # sharded env == slice of the global env (RNG streams keyed by global platoon id)
    P_total = 64
    Pl = P_total // world
    env = BatchedPlatoons(Pl, M, conf, platoon_id_base=rank * Pl, seed=9)
    env.reset()
    for _ in range(5):
        env.action_mu.zero_()
        env.step_native(explore=True, gen_exog=True)
    q.put((rank, bool(ok_a), bool(ok_c), env.state.cpu().numpy(), transports[2]))
    dist.barrier()
    dist.destroy_process_group()


def _world_sizes():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n] or [2]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", _world_sizes())
def test_interfrl_transports_and_sharded_env(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    print("peer transport:", res[0][4])
    from avddpg_b200.config import Config
    from avddpg_b200.environment import BatchedPlatoons
    conf = Config(pl_size=2)
    env = BatchedPlatoons(64, 2, conf, seed=9)
    env.reset()
    for _ in range(5):
        env.action_mu.zero_()
        env.step_native(explore=True, gen_exog=True)
    full = env.state.cpu().numpy()
    assert np.array_equal(np.concatenate([r[3] for r in res], axis=0), full)


def _timeout_worker(rank, world, port, q):
    """Rank 0 issues one round more than rank 1: its barrier must give up after AVD_PEER_TIMEOUT_MS and report the missing peer."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), AVD_PEER_TIMEOUT_MS="300")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from avddpg_b200.config import Config
    from avddpg_b200.server.federated import FederatedAggregator
    from avddpg_b200.trainer import DDPGPopulation
    conf = Config(pl_size=2, fed_method="interfrl", weighted_average_enabled=False)
    pop = DDPGPopulation(1, 2, conf)
    pop.actor.grad.fill_(1.0 + rank)
    pop.critic.grad.fill_(1.0 + rank)
    agg = FederatedAggregator(pop, conf, process_group=dist.group.WORLD, transport="peer")
    agg.aggregate_gradients(apply=False)                 # a healthy round: both ranks take part
    torch.cuda.synchronize()
    healthy = True
    try:
        agg.check_health()
    except RuntimeError:
        healthy = False
    mean_ok = bool(torch.allclose(pop.actor.grad, torch.full_like(pop.actor.grad, 1.5)))
    dist.barrier()
    raised = None
    if rank == 0:
        agg.aggregate_gradients(apply=False)             # nobody answers this one
        torch.cuda.synchronize()                         # returns after ~0.3 s instead of never
        try:
            agg.check_health()
            raised = ""
        except RuntimeError as e:
            raised = str(e)
    q.put((rank, healthy, mean_ok, raised))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_barrier_gives_up_and_reports_the_missing_rank():
    """a rank whose peers never issue the round must not spin forever (csrc/avd_peer.cu:peer_barrier, ctrl[2])."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_timeout_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res          # the healthy round: no flag, correct mean on both ranks
    assert res[0][3] and "rank 1 did not reach the barrier" in res[0][3], res
    assert res[1][3] is None
