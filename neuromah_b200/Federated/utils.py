"""federated_average (reference: Federated/utils.py:20-23): sum_i u_i * (n_i / N).

Host API mirror (lists in, list out) plus two device paths over the flat parameter slab (neuromah_b200/arena.py):
  * ``federated_average_device``  one client per GPU (BASELINE.json config 5): scale the slab by n_i / N in place, then
    SUM-all-reduce it over NCCL -- every client ends the round holding the new global model;
  * ``federated_average_slabs``   several clients' slabs resident on ONE GPU (the server side of the reference's
    Federated/Server.py aggregation): the same weighted sum accumulated in client order.
Optimizer state is not aggregated (the reference never ships it).
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


def federated_average_slabs(out_arena, client_arenas, num_samples):
    """``out_arena.params`` <- sum_i client_arenas[i].params * (n_i / N), all slabs on this GPU (same layer layout).
    ``out_arena`` may be one of the clients (it is read before it is overwritten)."""
    if len(client_arenas) != len(num_samples) or not client_arenas:
        raise ValueError("federated_average_slabs: one sample count per client")
    total = float(sum(num_samples))
    if any(a.total != out_arena.total for a in client_arenas):
        raise ValueError("federated_average_slabs: client models differ in layout")
    order = sorted(range(len(client_arenas)), key=lambda i: client_arenas[i] is not out_arena)    # the aliased one first
    for k, i in enumerate(order):
        lib.nm_scale_f32(out_arena.params.p, client_arenas[i].params.p, out_arena.total, float(num_samples[i]) / total,
                         0 if k == 0 else 1, stream())
    return out_arena.params
