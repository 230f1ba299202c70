"""Probe: torch symmetric memory (peer pointers over NVLink) on this box.  torchrun --nproc-per-node 2 tools/probe_symm.py"""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "symm ok:", type(hdl).__name__, "world", hdl.world_size, "rank", hdl.rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4],
          "signal pads", [hex(p) for p in hdl.signal_pad_ptrs][:4], "pad bytes", getattr(hdl, "signal_pad_size", None))
    t.fill_(rank + 1.0)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer read:", peer[:4].tolist())
    hdl.barrier()
    print(rank, "has multicast:", getattr(hdl, "multicast_ptr", None))
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symm FAILED:", repr(e))
print(rank, "can_access_peer", torch.cuda.can_device_access_peer(local, (local + 1) % world))
dist.barrier(); dist.destroy_process_group()
