"""Times the pieces of the peer-memory exchange (run under torchrun, >= 2 GPUs): copy-in, barrier, pull kernel, NCCL."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pynqs_b200 import peer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1_000_000 // world
x = torch.randint(0, 255, (n * 8,), dtype=torch.uint8, device=dev)
area = peer.PeerExchange.get(n * 16, dev)
out = torch.empty(world * n * 8, dtype=torch.uint8, device=dev)


def t(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


res = {}
res["copy_in_us"] = t(lambda: area.local(0, n * 8).copy_(x))
res["barrier_us"] = t(lambda: area.barrier())
res["pull_us"] = t(lambda: area.gather(0, n * 8, out))
res["full_us"] = t(lambda: (area.local(0, n * 8).copy_(x), area.barrier(), area.gather(0, n * 8, out), area.barrier()))
res["nccl_all_gather_us"] = t(lambda: dist.all_gather_into_tensor(out, x))
small = torch.zeros(7, dtype=torch.float64, device=dev)
outs = torch.zeros(7 * world, dtype=torch.float64, device=dev)
res["nccl_small_all_gather_us"] = t(lambda: dist.all_gather_into_tensor(outs, small))
if rank == 0:
    print(world, {k: round(v, 1) for k, v in res.items()}, flush=True)
torch.cuda.synchronize()
dist.barrier()
import os, sys  # noqa: E401,E402

sys.stdout.flush()
os._exit(0)  # (no destroy_process_group: it can block after the symmetric-memory rendezvous)
