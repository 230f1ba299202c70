"""BASELINE.json configs[4]: env-step throughput sweep, P = 2^10 .. 2^24 platoons, M in {4, 8}, on this GPU.
Two variants per point: `plain` (actions and leader inputs supplied: 48 B/vehicle-step) and `train` (OU noise +
clip + leader draw + replay-ring write + episodic bookkeeping fused: the launch the training loop makes).
Prints one JSON object; run under torchrun for N GPUs (each rank sweeps its own shard, value = sum).
    python tools/env_sweep.py > profiles/r01_env_sweep.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avddpg_b200.config import Config
from avddpg_b200.environment import BatchedPlatoons
from avddpg_b200.replaybuffer import ReplayRings

HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def time_env(P, M, mode, iters=30):
    conf = Config(pl_size=M, can_terminate=(mode == "train"))
    if mode == "train":
        rings = ReplayRings(4, M, P, 64)
        env = BatchedPlatoons(P, M, conf, ring=rings, clock=rings.clock, auto_reset=True, track_kinematics=False)
        step = lambda: env.step_native(explore=True, gen_exog=True, advance_clock=True)
        bpv = 48 + 16 + 40 + 8 + 8.0 / M   # +OU state r/w & applied action, +ring record, +episodic reward, +episode counters
    else:
        env = BatchedPlatoons(P, M, conf, track_kinematics=False, track_episodes=False, store_actions=False)
        step = lambda: env.step_native(leader_exog=True, advance_clock=False)
        bpv = 48 + 5.0 / M
    env.reset()
    env.action_mu.normal_(0, 0.5); env.leader_exog.normal_(0, 0.1)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[iters // 2]
    return {"P": P, "M": M, "mode": mode, "us": ms * 1e3, "vehicle_steps_per_s": P * M / (ms * 1e-3), "alg_bytes_per_vehicle_step": bpv,
            "alg_GBps": P * M * bpv / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": P * M * bpv / (ms * 1e-3) / 1e9 / HBM}


if __name__ == "__main__":
    out = []
    for M in (4, 8):
        for e in range(10, 25):
            P = 1 << e
            if P * M > (1 << 26):
                continue
            for mode in ("plain", "train"):
                if mode == "train" and P * M > (1 << 25):      # 4-slot ring = 160 B per vehicle: keep the sweep far from the HBM capacity
                    continue
                out.append(time_env(P, M, mode))
                torch.cuda.empty_cache()
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "hbm_peak_GBps": HBM, "points": out}, indent=1))
