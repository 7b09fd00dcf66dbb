"""Drop-in mirror of the reference's `patchcore` package surface for the clustering hot path
(Anomaly-Clustering/models/patchcore/{patchcore,common,utils}.py), routed to libac_b200.so."""
from . import common, patchcore, utils  # noqa: F401
