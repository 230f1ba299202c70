"""Step latency of the small BASELINE configs (launch-bound): C1 default run (1 platoon x 2 followers) and C3 (8 platoons x 4
followers, interfrl FedAvg every step), eager launches vs. the captured CUDA graph.
    python tools/small_configs.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avddpg_b200.config import Config
from avddpg_b200.trainer import BatchedTrainer


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


for name, kw, G in (("C1 1 platoon x 2 followers", dict(pl_size=2), 1),
                    ("C3 8 platoons x 4 followers, interfrl gradients", dict(pl_size=4, fed_method="interfrl", weighted_average_enabled=False), 8)):
    for prec in (0, 2):
        conf = Config(**kw)
        tr = BatchedTrainer(conf, num_groups=G, envs_per_group=1, ring_capacity=4096, precision=prec)
        for _ in range(conf.batch_size + 8):
            tr.step()
        dev, wall = timed(tr.step, 200)
        tr.capture()
        gdev, gwall = timed(tr.replay, 100)
        M = conf.pl_size
        print(f"{name:50s} precision={prec}: eager {dev*1e3:7.1f} us/step (host {wall*1e3:7.1f}) | graph {gdev/2*1e3:7.1f} us/step (host {gwall/2*1e3:7.1f})"
              f" -> {G*M/(gdev/2*1e-3):.3e} vehicle-steps/s")
