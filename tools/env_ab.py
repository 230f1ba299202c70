"""Median launch time of the env-step kernel, plain and training launch, for a same-box A/B of kernel variants.
    python tools/env_ab.py [tag]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from avddpg_b200.config import Config
from avddpg_b200.environment import BatchedPlatoons
from avddpg_b200.replaybuffer import ReplayRings


def t(P, M, train, iters=40):
    conf = Config(pl_size=M, can_terminate=train)
    if train:
        rings = ReplayRings(4, M, P, 64)
        env = BatchedPlatoons(P, M, conf, ring=rings, clock=rings.clock, auto_reset=True, track_kinematics=False, store_actions=train != 2)
        step = lambda: env.step_native(explore=True, gen_exog=True, advance_clock=False)
    else:
        env = BatchedPlatoons(P, M, conf, track_kinematics=False, track_episodes=False, store_actions=False)
        step = lambda: env.step_native(leader_exog=True, advance_clock=False)
    env.reset(); env.action_mu.normal_(0, 0.5); env.leader_exog.normal_(0, 0.1)
    for _ in range(5): step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[iters // 2] * 1e3


tag = sys.argv[1] if len(sys.argv) > 1 else ""
print(tag, " ".join(f"P=4Mi M={M} {('plain', 'train', 'train-noact')[tr]} {t(1 << 22, M, tr):.1f}us" for M in (4, 8) for tr in (0, 1, 2)), flush=True)
