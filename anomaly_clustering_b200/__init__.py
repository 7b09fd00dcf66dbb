"""anomaly_clustering_b200 -- B200-native (sm_100a) embedding-to-distance path of KevinWangHP/Anomaly-Clustering.

    build          python -m anomaly_clustering_b200.build   (nvcc -> libac_b200.so, the C ABI of include/ac_b200.h)
    _lib, ops      ctypes binding + torch-tensor wrappers (no CPU fallback: raises without the library / a B200)
    pipeline       features -> Z -> w -> alpha -> X -> Dmat on one GPU          (examples/main.py:266-296)
    distributed    the same, query-sharded over the GPUs of one box (NCCL)
    patchcore.*    mirror of the reference's patchcore.{patchcore,common,utils} call surface
    driver         make_category_data(...)                                      (examples/main.py:183-311)
    cluster, io    Ward + NMI/ARI/F1 consumer, reference pickle format          (examples/test.py)
    backbones      offline random-init WideResNet50 / ViT (the backbone forward stays in torch)
    synth          synthetic feature generators for tests and benchmarks

See DESIGN.md and INTEGRATION.md."""

__version__ = "0.1.0"
