"""Run the env-step kernel a few times on a >L2 population (for ncu / quick timing).
    python tools/profile_env.py [P] [M] [mode]    mode: plain | train (the training loop's launch: OU + clip + leader draw + replay
                                                        ring + episodic bookkeeping + auto-reset) | full (train + kinematics)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avddpg_b200.config import Config
from avddpg_b200.environment import BatchedPlatoons
from avddpg_b200.replaybuffer import ReplayRings

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
conf = Config(pl_size=M, can_terminate=False)
if mode == "train":
    conf = Config(pl_size=M)
    rings = ReplayRings(4, M, P, 64)
    env = BatchedPlatoons(P, M, conf, ring=rings, clock=rings.clock, auto_reset=True, track_kinematics=False)
elif mode == "full":
    rings = ReplayRings(8, M, P, 64)
    env = BatchedPlatoons(P, M, conf, ring=rings, clock=rings.clock)
else:
    env = BatchedPlatoons(P, M, conf, track_kinematics=False, track_episodes=False, store_actions=False)
env.reset()
env.action_mu.normal_(0, 0.5)
env.leader_exog.normal_(0, 0.1)
def step():
    if mode in ("full", "train"):
        env.step_native(explore=True, gen_exog=True, advance_clock=True)
    else:
        env.step_native(leader_exog=True, advance_clock=False)
for _ in range(5):
    step()
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in evs:
    a.record(); step(); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)
bpv = 48 + (16 + 8 + 12 + 40 + 8 if mode == "full" else 16 + 40 + 8 if mode == "train" else 0)   # +OU(16: ou r/w, action_out, mu counted in 48) +cum(8) +jerk/vel/headway(12) +ring(40) +ep_reward(8)
print(f"P={P} M={M} mode={mode}: median {ms[len(ms)//2]*1e3:.1f} us, min {ms[0]*1e3:.1f} us -> "
      f"{P*M/(ms[len(ms)//2]*1e-3):.3e} vehicle-steps/s, {P*M*bpv/(ms[len(ms)//2]*1e-3)/1e9:.0f} GB/s algorithmic ({bpv} B/vehicle-step)")
