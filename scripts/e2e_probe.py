"""Where does the end-to-end step go?  Times compute-only, copy-only and overlapped loops (config 2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import pipeline, synth  # noqa: E402

layers = [(768, 28, 28, True), (768, 28, 28, True)]
feats, _ = synth.planted_features_device(range(100), layers, device="cuda")
host = [f.cpu().pin_memory() for f in feats]
bufs = [[torch.empty_like(f) for f in feats] for _ in range(2)]
copy_stream = torch.cuda.Stream()
main = torch.cuda.current_stream()
K = 20


def step(f):
    r = pipeline.run_path(f, 3, 1, 2048, 4096, "unsupervised", [1.0])
    return r.alpha32, r.X, r.Dmat


def timeit(fn):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(K)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def compute_only(n):
    for i in range(n):
        step(bufs[i & 1])


def copy_only(n):
    for i in range(n):
        with torch.cuda.stream(copy_stream):
            for d_, h_ in zip(bufs[i & 1], host):
                d_.copy_(h_, non_blocking=True)
    main.wait_stream(copy_stream)


def serial(n):
    for i in range(n):
        for d_, h_ in zip(bufs[0], host):
            d_.copy_(h_, non_blocking=True)
        step(bufs[0])


def overlapped(n):
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    for e in freed:
        e.record(main)

    def issue(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            for d_, h_ in zip(bufs[slot], host):
                d_.copy_(h_, non_blocking=True)
            ready[slot].record(copy_stream)

    issue(0)
    for i in range(n):
        s = i & 1
        if i + 1 < n:
            issue(s ^ 1)
        main.wait_event(ready[s])
        step(bufs[s])
        freed[s].record(main)


for name, fn in (("compute only", compute_only), ("copy only", copy_only), ("serial copy+compute", serial), ("overlapped", overlapped)):
    print("%-22s %.2f ms/step" % (name, timeit(fn)), flush=True)


# ---- timeline of the overlapped loop (event timestamps relative to the first one)
def timeline(n=6):
    ready = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    cstart = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    kstart = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    kend = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record(main)
    copy_stream.wait_event(t0)

    def issue(i):
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(kend[i - 2])
            cstart[i].record(copy_stream)
            for d_, h_ in zip(bufs[i & 1], host):
                d_.copy_(h_, non_blocking=True)
            ready[i].record(copy_stream)

    issue(0)
    for i in range(n):
        if i + 1 < n:
            issue(i + 1)
        main.wait_event(ready[i])
        kstart[i].record(main)
        step(bufs[i & 1])
        kend[i].record(main)
    torch.cuda.synchronize()
    for i in range(n):
        print("step %d: copy %.1f..%.1f   compute %.1f..%.1f" % (i, t0.elapsed_time(cstart[i]), t0.elapsed_time(ready[i]),
                                                                t0.elapsed_time(kstart[i]), t0.elapsed_time(kend[i])))


timeline()
