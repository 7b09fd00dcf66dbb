"""torchrun --nproc-per-node N scripts/run_configs_sharded.py : BASELINE configs 3 (supervised, 100 queries vs a
200-image bank) and 5 (64 images x 3136 patches, 6 taus from one distance pass) on N GPUs, device-timed, max over
ranks.  Companion of scripts/run_configs.py (the 1-GPU numbers in DESIGN.md section 5b)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from anomaly_clustering_b200 import distributed, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29561")
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {}
vitb = [(768, 28, 28, True), (768, 28, 28, True)]
(lo, hi), (blo, bhi) = distributed.shard_bounds(100, world)[rank], distributed.shard_bounds(200, world)[rank]
q, _ = synth.planted_features_device(range(lo, hi), vitb, device="cuda")
bank, _ = synth.planted_features_device(range(1000 + blo, 1000 + bhi), vitb, n_classes=1, device="cuda")
for overlap in (True, False):
    ms = timed(lambda: distributed.run_path_sharded_supervised(q, 100, bank, 200, 3, 1, 2048, 4096, [1.0], overlap=overlap), 5)
    out["config3_supervised_overlap_%s" % overlap] = {"ms": ms, "images_per_s": 100 / ms * 1e3,
                                                        "algorithmic_tflops": 2 * 78400 * 156800 * 4096 / ms / 1e9}
del q, bank
vits = [(384, 56, 56, True), (384, 56, 56, True)]
lo, hi = distributed.shard_bounds(64, world)[rank]
q5, _ = synth.planted_features_device(range(lo, hi), vits, device="cuda")
taus = [0.1, 0.5, 1, 2, 5, 10]
for prec in ("f16", "auto"):      # auto = f16x3 here (min tau < 0.5)
    ms = timed(lambda: distributed.run_path_sharded(q5, 64, 3, 1, 2048, 4096, taus, precision=prec), 3)
    out["config5_%s" % prec] = {"ms": ms, "images_per_s": 64 / ms * 1e3}
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}))
dist.barrier()
dist.destroy_process_group()
