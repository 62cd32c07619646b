"""Host<->device copy bandwidth of one box: every rank alone, then all ranks at once (torchrun).  Explains what bounds
the end-to-end leg of bench.py at N = 8 (pinned host buffers, one process per GPU).  Diagnostic."""
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
MB = 256
host = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
dev = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d: bool, d2h: bool, reps: int = 8) -> float:
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    up.wait_event(e0)
    down.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(up):
                dev.copy_(host, non_blocking=True)
        if d2h:
            with torch.cuda.stream(down):
                host.copy_(dev, non_blocking=True) if not h2d else host2.copy_(dev2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(up)
    torch.cuda.current_stream().wait_stream(down)
    e1.record()
    torch.cuda.synchronize()
    return reps * MB / 1024 / (e0.elapsed_time(e1) * 1e-3)      # GB/s per direction


host2 = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
dev2 = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, (a, b) in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
    run(a, b, 2)
    v = torch.tensor([run(a, b)], device="cuda")
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        res[name] = [round(float(x), 1) for x in allv]
    else:
        res[name] = [round(float(v), 1)]
if rank == 0:
    for k, v in res.items():
        print(f"{world} ranks at once, {k}: per rank GB/s (each direction) {v}  sum {sum(v):.0f}")
if world > 1:
    dist.destroy_process_group()
