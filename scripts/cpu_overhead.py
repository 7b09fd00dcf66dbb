"""Host-side (Python + launch) time per step of the sharded path: tiny shapes so the GPU is never the
bottleneck.  torchrun --nproc-per-node N scripts/cpu_overhead.py"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from anomaly_clustering_b200 import distributed, pipeline, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_img = 40
layers = [(96, 12, 12, True), (96, 12, 12, True)]
lo, hi = distributed.shard_bounds(n_img, world)[rank]
feats, _ = synth.planted_features_device(range(lo, hi), layers, device="cuda")


def step():
    if world == 1:
        return pipeline.run_path(feats, 3, 1, 256, 512, "unsupervised", [1.0])
    return distributed.run_path_sharded(feats, n_img, 3, 1, 256, 512, [1.0])


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
if rank == 0:
    print("world=%d: host enqueue %.3f ms/step, wall incl. drain %.3f ms/step" % (world, (t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))
if world == 1:   # cProfile only single-process: every rank must execute the same collectives
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
