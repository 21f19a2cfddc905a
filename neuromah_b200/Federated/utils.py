"""federated_average (reference: Federated/utils.py:20-23): sum_i u_i * (n_i / N).

Host API mirror (lists in, list out) plus the device path used by FLTrainer /
config 5: scale the weight slab by n_i/N in place and SUM-all-reduce it over NCCL.
"""
import numpy as np

from .._lib import lib
from ..device import stream


def federated_average(updates, num_samples):
    total_samples = sum(num_samples)
    weighted = [np.asarray(u, dtype=np.float64) * (s / total_samples) for u, s in zip(updates, num_samples)]
    return np.sum(weighted, axis=0).tolist()


def federated_average_device(arena, n_local, n_total, comm):
    """In-place FedAvg of ``arena.params`` across the ranks of ``comm`` (one client per GPU)."""
    lib.nm_scale_f32(arena.params.p, arena.params.p, arena.total, float(n_local) / float(n_total), 0, stream())
    if comm is not None and comm.world > 1:
        comm.allreduce(arena.params, arena.total)
    return arena.params
