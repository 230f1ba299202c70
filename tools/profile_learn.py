"""Run the C2 learn step a few times (for ncu launch lists / captures).
    python tools/profile_learn.py [precision] [envs_per_group] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avddpg_b200.config import Config
from avddpg_b200.trainer import DDPGPopulation

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
conf = Config(pl_size=4)
pop = DDPGPopulation(1, 4, conf, rows_per_agent=E * 64, precision=prec)
n = pop.A * pop.R
g = torch.Generator(device="cuda").manual_seed(0)
s = torch.randn(n, 4, device="cuda", generator=g) * 2
a = (torch.rand(n, device="cuda", generator=g) * 5 - 2.5)
r = -torch.rand(n, device="cuda", generator=g) * 0.5
s2 = s + 0.1 * torch.randn(n, 4, device="cuda", generator=g)
for _ in range(2):
    pop.learn(s, a, r, s2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    pop.learn(s, a, r, s2)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"precision={prec} rows={n}: learn {ms:.3f} ms  -> {2*340464*n/ms/1e9:.1f} TFLOP/s algorithmic, {n/64/ms*1e3:.3e} minibatch-updates/s")
