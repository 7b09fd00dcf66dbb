"""[torchrun --nproc-per-node N] scripts/check_supervised_sharded.py : the sharded supervised path (queries and
the normal-image bank both sharded, bank operands all-gathered, local bank shard multiplied during the gather)
against the single-GPU supervised path on the same synthetic images.  Works with 1 rank too."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from anomaly_clustering_b200 import distributed, pipeline, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
ok = True
for n_img, n_bank, layers, Dp, D in ((13, 9, [(96, 12, 12, True), (96, 12, 12, True)], 256, 512),
                                     (10, 21, [(768, 28, 28, True), (768, 28, 28, True)], 2048, 4096)):
    (lo, hi), (blo, bhi) = distributed.shard_bounds(n_img, world)[rank], distributed.shard_bounds(n_bank, world)[rank]
    feats, _ = synth.planted_features_device(range(lo, hi), layers, device="cuda")
    bank, _ = synth.planted_features_device(range(5000 + blo, 5000 + bhi), layers, device="cuda")
    for overlap in (True, False):
        a64, X, Dm, w = distributed.run_path_sharded_supervised(feats, n_img, bank, n_bank, 3, 1, Dp, D, [1.0, 2.0], overlap=overlap)
        if rank == 0:
            allf, _ = synth.planted_features_device(range(n_img), layers, device="cuda")
            allb, _ = synth.planted_features_device(range(5000, 5000 + n_bank), layers, device="cuda")
            ref = pipeline.run_path(allf, 3, 1, Dp, D, "supervised", [1.0, 2.0], bank_features=allb)
            e_w = ((w - ref.w[lo:hi]).abs() / ref.w[lo:hi]).max().item()
            e_a = (a64 - ref.alpha64[:, lo:hi]).abs().max().item()
            e_x = ((X - ref.X).norm() / ref.X.norm()).item()
            e_d = ((Dm - ref.Dmat).norm() / ref.Dmat.norm()).item()
            good = e_w < 2e-4 and e_a < 1e-3 and e_x < 1e-4 and e_d < 1e-4
            ok &= good
            print("supervised: %d query / %d bank images on %d ranks  overlap=%s: w rel %.1e  alpha abs %.1e  X relL2 %.1e  "
                  "Dmat relL2 %.1e  %s" % (n_img, n_bank, world, overlap, e_w, e_a, e_x, e_d, "OK" if good else "MISMATCH"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
