"""Host <-> device staging ceiling of the box (development aid, not a test): pure cudaMemcpyAsync from / to page-locked
memory, both directions at once, on every rank at the same time.

    python tests/gpu_pcie_probe.py                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_pcie_probe.py                                      N GPUs of one box, concurrently

Rank 0 prints one JSON line: per-direction GB/s alone and with both directions busy, per rank (min / max) and summed over
the ranks.  bench.py's end-to-end number is reported as a fraction of the "both" figure (its copies run both ways at once).
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 29   # 512 MiB per buffer
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t) / 1e9


run(True, True, 1)
vals = [run(True, False), run(False, True), run(True, True)]
t = torch.tensor(vals, dtype=torch.float64, device="cuda")
if world > 1:
    lo, hi, sm = t.clone(), t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
else:
    lo = hi = sm = t
if rank == 0:
    names = ["h2d_only", "d2h_only", "both_each_direction"]
    print(json.dumps(dict(n_gpus=world, cpus=os.cpu_count(), unit="GB/s",
                          per_rank_min={k: round(float(v), 1) for k, v in zip(names, lo)},
                          per_rank_max={k: round(float(v), 1) for k, v in zip(names, hi)},
                          all_ranks_sum={k: round(float(v), 1) for k, v in zip(names, sm)})), flush=True)
if world > 1:
    dist.destroy_process_group()
