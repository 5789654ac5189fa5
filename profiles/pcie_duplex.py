"""Copy-only limits of the host: pinned-memory H2D and D2H per GPU, alone and together, on 1..N GPUs at once (no kernels).
Run under torchrun with N ranks; rank 0 prints one JSON line per group size (1, 2, 4, ... N ranks copying concurrently)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 2 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, active, reps=3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps if active else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt)


run(True, True, True, 1)
g = 1
while g <= world:
    active = rank < g
    a, b, c = run(True, False, active), run(False, True, active), run(True, True, active)
    if rank == 0:
        print(json.dumps({"gpus_copying": g, "h2d_alone_gbs_per_gpu": round(n / a / 1e9, 1), "d2h_alone_gbs_per_gpu": round(n / b / 1e9, 1),
                          "both_gbs_per_gpu_each_way": round(n / c / 1e9, 1), "aggregate_both_gbs": round(2 * g * n / c / 1e9, 1)}), flush=True)
    g *= 2
if world > 1:
    dist.destroy_process_group()
