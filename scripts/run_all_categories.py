"""BASELINE config 4, per-category form (the reference's own semantics: one make_category_data per category,
examples/main.py:353): the 10 MVTec-object test-set sizes, each category an independent problem with its own
bank.  Categories are assigned to ranks by greedy LPT on n_c*(n_c-1); no collective
on the data path.  [torchrun --nproc-per-node N] python scripts/run_all_categories.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from anomaly_clustering_b200 import distributed, pipeline, synth  # noqa: E402

COUNTS = [83, 150, 132, 110, 115, 167, 160, 42, 100, 151]     # bottle ... zipper (info_<cat>.pickle sizes)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mine = distributed.lpt_assign([n * (n - 1) for n in COUNTS], world)[rank]
layers = [(768, 28, 28, True), (768, 28, 28, True)]
feats = {c: synth.planted_features_device(range(1000 * c, 1000 * c + COUNTS[c]), layers, device="cuda")[0] for c in mine}


def step():
    return [pipeline.run_path(feats[c], 3, 1, 2048, 4096, "unsupervised", [1.0]).Dmat for c in mine]


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    step()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / K], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    flops = sum(2.0 * n * (n - 1) * 784 * 784 * 4096 for n in COUNTS)
    print(json.dumps({"workload": "config4 per-category banks (10 categories, 1210 images)", "n_gpus": world, "ms_per_pass": ms,
                      "images_per_s": sum(COUNTS) / ms * 1e3, "algorithmic_tflops": flops / ms / 1e9,
                      "lpt_assignment_rank0": mine}))
if world > 1:
    dist.destroy_process_group()
