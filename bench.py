#!/usr/bin/env python
"""bench.py -- avddpg hot path on B200: platoon env-steps/s (+ DDPG updates/s) vs the reference CPU path.

    python bench.py --gpus N --steps K --warmup W            # native arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU port of the reference loop

One "step" = one pass of the hot path over the BASELINE.json configs[1] population on each GPU:
4096 platoons x 4 followers (16,384 vehicles): act (OU noise + clip) -> Platoon.step -> ReplayBuffer.add
for every agent -> replay sample (64 per ring) [-> DDPG learn + Adam + Polyak when the learn kernels are
enabled].  Weak scaling: every rank owns its own 4096 platoons (global platoon ids are offset by rank, so
RNG streams do not depend on the GPU count); there is no data-path collective in the env/replay path.

The JSON line also carries
  roofline     : the env-step kernel on a population larger than L2 (HBM-bound), CUDA-event timed
  e2e          : the same step driven through host buffers (pinned H2D of the leader inputs, D2H of the
                 per-step reward/done statistics) every step
  cpu_baseline : oracle port of the reference loop on this box's host cores (rank 0, N=1 only)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "platoon-vehicle env-steps/s"
WORKLOAD = "C2: 4096 platoons x 4 followers per GPU, decentralized Model B euler, OU noise, replay cap 100000, batch 64"
P_C2, M_C2, RING_CAP, BATCH = 4096, 4, 100_000, 64


def _ncu_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` captures (profiles/r01_traffic.json, written by
    tools/ncu_traffic.py): {"env": bytes per env_step_kernel launch at the roofline population,
    "learn": bytes summed over the launches of one learn step at C2}.  None when the file is missing."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path):
        return {"env": None, "learn": None}
    d = json.load(open(path))
    env = d.get("r01_env_plain.ncu-rep") or []
    learn = d.get("r01_learn_step.ncu-rep") or []
    tot = lambda rows: float(sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in rows)) if rows else None
    return {"env": tot(env[:1]), "learn": tot(learn)}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = int(float(self.rows[0][1])) if self.rows else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _dist_setup(n_gpus):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def _max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


# --------------------------------------------------------------------------------------------- native arm
LEARN_MACS_PER_SAMPLE = 340_464          # reference-equivalent MACs per sampled transition
ENV_BYTES_PER_VEHICLE_STEP = 48


def build_trainer(rank, world, args, pg=None):
    import torch
    from avddpg_b200.config import Config
    from avddpg_b200.trainer import BatchedTrainer
    conf = Config(pl_size=M_C2, num_platoons=P_C2)
    free, _ = torch.cuda.mem_get_info()
    ring_cap = RING_CAP
    if ring_cap * M_C2 * P_C2 * 40 > 0.6 * free:      # never drive the box out of memory: shrink the ring and say so
        ring_cap = int(0.4 * free / (M_C2 * P_C2 * 40))
    tr = BatchedTrainer(conf, num_groups=1, envs_per_group=P_C2, ring_capacity=ring_cap, rank=rank, world=world,
                        process_group=pg, precision=args.precision)
    tr.rings.fill_synthetic()                # steady state: sampling range == capacity from the first timed step
    tr.buffer_counter = ring_cap
    return conf, tr


def time_env_roofline(P_big, M, steps=20, warmup=5):
    """Env-step kernel alone on a population whose working set exceeds L2 (126 MB): CUDA events around each
    launch on the launching stream, average duration -> achieved algorithmic GB/s (48 B per vehicle-step:
    x[4], prev_a, u in; x'[4], prev_a', reward out; +5 B per platoon for the leader input and the done flag)."""
    import torch
    from avddpg_b200.config import Config
    from avddpg_b200.environment import BatchedPlatoons
    conf = Config(pl_size=M, can_terminate=False)
    env = BatchedPlatoons(P_big, M, conf, track_kinematics=False, track_episodes=False, store_actions=False)
    env.reset()
    env.action_mu.normal_(0, 0.5)
    env.leader_exog.normal_(0, 0.1)
    for _ in range(warmup):
        env.step_native(leader_exog=True, advance_clock=False)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        env.step_native(leader_exog=True, advance_clock=False)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    avg = sum(ms) / len(ms)
    alg = P_big * (M * ENV_BYTES_PER_VEHICLE_STEP + 5.0)
    del env
    torch.cuda.empty_cache()
    return dict(avg_ms=avg, min_ms=ms[0], alg_bytes=alg, working_set_mb=P_big * M * 4 * 11 / 1e6)


def _timed(fn, n, world):
    import torch
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    _barrier(world)
    return _max_over_ranks(e0.elapsed_time(e1), world) / n


def run_native(args):
    import torch
    rank, world, local = _dist_setup(args.gpus)
    from avddpg_b200 import _lib
    _lib.require_device()
    hbm_peak, tf_peak, peak_src = _peaks()
    conf, tr = build_trainer(rank, world, args)
    env, rings, pop = tr.env, tr.rings, tr.pop
    P, M = env.P, env.M
    warm = max(3, args.warmup)
    for _ in range(warm):
        tr.step()
    if args.graph:
        tr.capture(warmup=1)
    step_fn = (lambda: tr.replay()) if args.graph else (lambda: tr.step())
    steps_per_call = 2 if args.graph else 1
    calls = max(1, args.steps // steps_per_call)

    # ---- device-resident timing of the whole training step (inputs already in HBM)
    from avddpg_b200 import _lib as _avd
    l0 = _avd.load().avd_kernel_launches()      # counted inside the library, one per kernel launch site executed
    clk = ClockSampler(local)        # nvidia-smi takes ~0.1 s per query: keep sampling through all timed legs (device-timed steps,
    clk.__enter__()                  # attribution, end-to-end) so that the clocks line rests on more than one sample under load
    ms_call = _timed(step_fn, calls, world)
    ms_step = ms_call / steps_per_call
    launches = _avd.load().avd_kernel_launches() - l0
    if args.graph:      # replays do not pass through the launch sites: kernels per captured step x replayed steps
        launches = calls * steps_per_call * tr.kernels_per_step
    value = world * P * M / (ms_step * 1e-3)

    # ---- attribution: env part (act + env step + replay add) and learn part (sample + learn + Adam + Polyak) alone
    def env_part():
        pop.act(env.native_state, env.action_mu, tr.E)
        env.step_native(explore=True, gen_exog=True, advance_clock=False)
        rings.clock.advance(step=1, ring=1)

    def learn_part():
        s, a, r, s2 = rings.sample(advance_clock=True)
        pop.learn(s, a, r, s2, apply_updates=True)

    n_attr = max(5, min(30, args.steps // 4))
    ms_env = _timed(env_part, n_attr, world)
    ms_learn = _timed(learn_part, n_attr, world)

    # ---- end to end through host buffers: pinned leader inputs in, reward/done statistics + losses out, every step.
    # HostStepPipeline double-buffers both directions, so the host reads the results of step k - 1 while step k runs.
    from avddpg_b200.trainer import HostStepPipeline
    pipe = HostStepPipeline(tr)
    gen = torch.Generator().manual_seed(1 + rank)
    acc = 0.0

    def e2e_step():
        pipe.input_buffer().normal_(0, 0.1, generator=gen)
        prev = pipe.submit()
        return float(prev[0]) if prev is not None else 0.0

    for _ in range(3):
        e2e_step()
    pipe.drain()
    _barrier(world)
    n_e2e = max(5, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        acc += e2e_step()
    acc += float(pipe.drain()[0])                    # the last step's results: the timed region ends with the GPU drained
    _barrier(world)
    e2e_ms = _max_over_ranks((time.perf_counter() - t0) * 1e3, world) / n_e2e
    clk.__exit__(None, None, None)
    e2e = {"value": world * P * M / (e2e_ms * 1e-3), "unit": METRIC, "h2d_bytes_per_step": pipe.h2d_bytes_per_step,
           "d2h_bytes_per_step": pipe.d2h_bytes_per_step, "ms_per_step": e2e_ms, "steps": n_e2e,
           "api": "HostStepPipeline(BatchedTrainer).submit(): pinned H2D of the leader inputs + eager step + D2H of reward/done "
                  "statistics and losses every step; results are read on the host one step behind (double-buffered pinned buffers)"}

    # ---- FRL round (interfrl, gradients): local reduce -> ONE all_reduce over NVLink -> scale -> broadcast -> Adam x2 -> Polyak x2
    from avddpg_b200.config import Config as _Config
    from avddpg_b200.server.federated import FederatedAggregator
    pg = None
    if world > 1:
        import torch.distributed as dist
        pg = dist.group.WORLD
    agg = FederatedAggregator(pop, _Config(pl_size=M, fed_method="interfrl", weighted_average_enabled=False), process_group=pg)
    for _ in range(5):
        agg.aggregate_gradients()
    ms_frl = _timed(lambda: agg.aggregate_gradients(), 50, world)
    ms_frl_exchange = _timed(lambda: agg.aggregate_gradients(apply=False), 50, world)
    frl = {"mode": "interfrl / gradients / unweighted", "round_us": ms_frl * 1e3, "reduce_exchange_broadcast_us": ms_frl_exchange * 1e3,
           "payload_bytes": int(M * (pop.actor.n_train + pop.critic.n_train + 1) * 4), "ranks": world, "transport": agg.transport,
           "includes": "avd_fed_reduce + exchange [N>1: one NVLink kernel, in-switch NVLS reduction when available] + scale + "
                       "avd_fed_broadcast + Adam x2 + Polyak x2"}
    if world > 1:       # the NCCL transport (all_reduce + finalize kernel) for comparison
        agg_nccl = FederatedAggregator(pop, _Config(pl_size=M, fed_method="interfrl", weighted_average_enabled=False), process_group=pg,
                                       transport="nccl")
        for _ in range(5):
            agg_nccl.aggregate_gradients(apply=False)
        frl["reduce_exchange_broadcast_us_nccl"] = _timed(lambda: agg_nccl.aggregate_gradients(apply=False), 50, world) * 1e3

    out = None
    if rank == 0:
        rows = pop.A * pop.R
        learn_flops = 2.0 * LEARN_MACS_PER_SAMPLE * rows
        learn_tf = learn_flops / (ms_learn * 1e-3) / 1e12
        rl = time_env_roofline(args.roofline_platoons, M)
        traffic = _ncu_traffic()
        achieved = rl["alg_bytes"] / (rl["avg_ms"] * 1e-3) / 1e9
        roofline_env = {"bound": "hbm", "kernel": "env_step_kernel<4>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": traffic["env"] if args.roofline_platoons == 4 * 1024 * 1024 else None,
                        "traffic_unit": "DRAM bytes per launch (ncu, profiles/r01_env_step_plain_ncu.txt)", "alg_bytes_per_launch": rl["alg_bytes"],
                        "peak_source": peak_src,
                        "population": f"{args.roofline_platoons} platoons x {M} (state working set {rl['working_set_mb']:.0f} MB > 126 MB L2)",
                        "alg_bytes_per_vehicle_step": ENV_BYTES_PER_VEHICLE_STEP, "avg_launch_ms": rl["avg_ms"],
                        "vehicle_steps_per_s": args.roofline_platoons * M / (rl["avg_ms"] * 1e-3)}
        roofline_learn = {"bound": "tensor", "kernel": ("learn step: 6 fused pass launches (fused3_kernel) + 2 wgrad3 + 2 dgrad3, bf16 tcgen05"
                                                       if args.precision else "learn step, fp32 SIMT parity mode"),
                          "achieved": learn_tf, "peak": tf_peak / 1e0, "unit": "TFLOP/s", "frac": learn_tf / tf_peak,
                          "traffic": traffic["learn"] if args.precision else None,
                          "traffic_unit": "DRAM bytes per learn step, summed over its launches (ncu, profiles/r01_learn_kernels_ncu.txt)",
                          "peak_source": peak_src + " bf16 sustained", "alg_flops_per_step": learn_flops,
                          "alg_macs_per_sample": LEARN_MACS_PER_SAMPLE, "rows_per_step": rows, "learn_ms": ms_learn}
        dominant = roofline_learn if ms_learn > ms_env else roofline_env
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline
            cpu = cpu_baseline.time_env_steps(M=M, target_seconds=args.cpu_seconds, with_learn=True)
            cpu_env = cpu_baseline.time_env_steps(M=M, target_seconds=max(2.0, args.cpu_seconds / 3), with_learn=False)
            cpu["env_only_loop"] = {"value": cpu_env["value"], "sample": cpu_env["sample"]}
        out = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": calls * steps_per_call, "warmup": warm,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16" if args.precision else "f32", "data": "synthetic", "impl": "native",
               "config": {"workload": WORKLOAD, "platoons_per_gpu": P, "followers": M, "agents_per_gpu": pop.A,
                          "rows_per_agent_update": pop.R, "ring_capacity": rings.capacity, "cuda_graph": bool(args.graph),
                          "l2": "C2 state is 0.8 MB/step (L2-resident by nature); replay gathers hit a pre-filled "
                                f"{rings.capacity * M * P * 40 / 1e9:.1f} GB ring and the learn workspace is "
                                f"{pop._ws.numel() / 1e9:.1f} GB (both >> 126 MB L2); env roofline measured on a >L2 population",
                          "step": "act(actor fwd) + OU/clip + Platoon.step + ReplayBuffer.add + sample(64/ring) + learn + Adam x2 + Polyak"},
               "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches,
               "roofline": dominant, "roofline_env": roofline_env, "roofline_learn": roofline_learn, "cpu_baseline": cpu,
               "platoon_steps_per_s": value / M,
               "ddpg_minibatch_updates_per_s": world * P * M / (ms_step * 1e-3),
               "ddpg_weight_updates_per_s": world * pop.A / (ms_step * 1e-3),
               "ms_env_part": ms_env, "ms_learn_part": ms_learn, "frl": frl}
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return out


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """Reference arm: the CPU port of the reference's own loop (oracle/), all host cores, rank 0 only.
    Each bench "step" is a bounded sample: every core steps one 4-follower platoon for ~budget seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    total = max(1, args.steps)
    budget = max(0.25, min(args.cpu_seconds, 150.0 / (total + args.warmup)))
    pool = cpu_baseline.EnvLoopPool(M=M_C2, with_learn=True)
    vals, steps, slowest = [], 0, 0.0
    try:
        for i in range(args.warmup + total):
            v, steps, slowest = pool.run(budget)
            if i >= args.warmup:
                vals.append(v)
    finally:
        pool.close()
    v = sum(vals) / len(vals)
    line = {"metric": METRIC, "value": v, "unit": METRIC, "n_gpus": args.gpus, "steps": total, "warmup": args.warmup,
            "ms_per_step": slowest * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": v, "unit": METRIC, "cores": pool.cores, "kind": "port", "sample": pool.describe(steps, slowest)},
            "e2e": {"value": v, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--roofline-platoons", type=int, default=4 * 1024 * 1024)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", type=int, default=1, help="0: fp32 SIMT learn kernels (parity mode); 1: bf16 tcgen05 GEMMs")
    ap.add_argument("--graph", action="store_true", help="replay a captured CUDA graph of the training step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
