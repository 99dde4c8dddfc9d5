"""CPU, world_size 2, gloo: the multi-GPU plumbing of bench.py (frame sharding by rank, max-over-ranks timing,
whole-job throughput).  Images are independent, so there is no data-path collective to test."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    ms, e2e, launches = bench.reduce_over_ranks(100.0 + 50.0 * rank, 200.0 - 10.0 * rank, 1000 + rank, torch.device("cpu"))
    tmax = bench.reduce_max([1.0 + rank, 5.0 - rank], torch.device("cpu"))
    q.put((rank, ms, e2e, launches, bench.frame_seed(rank), bench.dist_env()[:2], tmax, bench.batch_seeds(rank, 4)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_and_sharding():
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, e2e, launches, seed, env, tmax, bseeds in res:
        assert ms == 150.0 and e2e == 200.0 and launches == 2001   # max, max, sum
        assert seed == rank + 1 and env == (rank, world)           # one frame stream per rank
        assert tmax == [2.0, 5.0]                                  # batch config: slowest rank per direction
        assert bseeds == [rank * 512 + i for i in range(4)]        # BASELINE config 4: rank r owns the seeds r*512 ..
    import bench
    # weak scaling: each rank did `steps` frames in the slowest rank's time
    assert abs(bench.job_mpixels_per_s(7680 * 4320, 5, 2, 150.0) - 2 * 5 * 33.1776 / 0.150) < 1e-6
