"""Data-parallel plumbing: one process per GPU; the gradient slab is summed over ranks either by the fused peer-memory
kernel (nm_dp_*: push over NVLink + rank-ordered sum + Adam in one launch) or by an NCCL all-reduce.

Sharding: rank r owns rows [r*b, (r+1)*b) of every global batch.  The
softmax-CE gradient and the Dense parameter gradients are divided by the GLOBAL batch
(``set_global_batch``), conv gradients are unnormalised sums of already-scaled dvalues, so
a plain SUM all-reduce of the gradient slab reproduces the single-GPU batch-B step exactly.

Bootstrap uses torch.distributed (whatever backend the launcher initialised; gloo works)
only to ship the 128-byte NCCL unique id and the 64-byte cudaIpc handles of the peer
mailboxes; the data path is our own kernels / our own communicator on our own stream.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

from ._lib import lib
from .device import new_event, new_stream, stream


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n rows for ``rank``; n must divide evenly."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    b = n // world
    return rank * b, (rank + 1) * b


def set_global_batch(model, global_batch: int | None):
    """Make every batch-normalised gradient divide by the global batch (None = local)."""
    from .src.layers.Dense import Layer_Dense
    for layer in model.layers:
        if isinstance(layer, Layer_Dense):
            layer.grad_batch = global_batch
    if model.softmax_classifier_output is not None:
        model.softmax_classifier_output.grad_batch = global_batch


def find_libnccl() -> str:
    """Prefer the NCCL wheel next to torch (2.28.x), fall back to the system library."""
    try:
        import nvidia.nccl as _n      # type: ignore
        for root in list(getattr(_n, "__path__", [])):
            hits = glob.glob(os.path.join(root, "lib", "libnccl.so*"))
            if hits:
                return sorted(hits)[0]
    except Exception:
        pass
    return "libnccl.so.2"


class Communicator:
    """NCCL communicator owned by libneuromah_b200 (one per process)."""

    def __init__(self, rank: int, world: int, unique_id: bytes | None = None):
        self.rank, self.world = rank, world
        self.handle = None
        self.peer = None                 # nm_dp context: fused all-reduce + Adam over NVLink peer memory
        self.comm_stream = None          # side stream for all-reduces overlapped with the rest of backward
        if world == 1:
            return
        lib.nm_comm_load(find_libnccl().encode())
        if unique_id is None:
            unique_id = self._exchange_id()
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        lib.nm_comm_init(C.byref(h), buf, rank, world)
        self.handle = h

    def _exchange_id(self) -> bytes:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised to bootstrap the NCCL id")
        t = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            lib.nm_comm_unique_id(buf)
            t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())

    def _all_gather_bytes(self, blob: bytes) -> bytes:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised to exchange the peer-memory handles")
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)

    def enable_peer(self, max_floats: int, required: bool = False) -> bool:
        """Set up the peer-memory mailboxes for gradient slabs of up to ``max_floats`` floats.  Returns False (and
        leaves the NCCL path in charge) when some pair of GPUs cannot map each other's memory; every rank takes the
        same decision."""
        if self.world == 1 or self.peer is not None:
            return self.peer is not None
        import torch
        import torch.distributed as dist
        h = C.c_void_p()
        ok, err = 1, ""
        try:
            lib.nm_dp_create(C.byref(h), self.rank, self.world, int(max_floats))
            mine = C.create_string_buffer(64)
            lib.nm_dp_export(h, mine)
        except Exception as e:          # noqa: BLE001 -- agreement with the peers comes first
            ok, err, mine = 0, str(e), C.create_string_buffer(64)
        handles = self._all_gather_bytes(mine.raw)
        if ok:
            try:
                lib.nm_dp_connect(h, C.create_string_buffer(handles, 64 * self.world))
            except Exception as e:      # noqa: BLE001
                ok, err = 0, str(e)
        flags = self._all_gather_bytes(bytes([ok]))
        if all(flags):
            self.peer = h
            return True
        if h:
            lib.nm_dp_destroy(h)
        if required:
            raise RuntimeError(f"peer-memory data parallelism unavailable (rank {self.rank}: {err or 'a peer failed'})")
        return False

    def check(self):
        """Raise if a peer-memory wait timed out (a rank died or never launched its step); blocking."""
        if self.peer is None:
            return
        bad, done = C.c_int(0), C.c_uint(0)
        lib.nm_dp_status(self.peer, C.byref(bad), C.byref(done))
        if bad.value:
            raise RuntimeError(f"rank {self.rank}: a peer did not deliver its gradients in time (step {done.value})")

    def allreduce(self, arr, count=None):
        if self.world == 1:
            return
        lib.nm_comm_allreduce_sum_f32(self.handle, arr.p, arr.size if count is None else count, stream())

    def allreduce_grads(self, arena):
        self.allreduce(arena.grads, arena.total)

    def allreduce_async(self, arr, offset: int, count: int):
        """SUM all-reduce of ``arr[offset : offset + count]`` (floats) on the communication stream, ordered after
        everything enqueued on the compute stream so far; the compute stream keeps going.  ``wait()`` joins."""
        if self.world == 1 or count <= 0:
            return
        if self.comm_stream is None:
            self.comm_stream = new_stream()
            self._ready, self._done = new_event(), new_event()
        lib.nm_event_record(self._ready, stream())
        lib.nm_stream_wait_event(self.comm_stream, self._ready)
        lib.nm_comm_allreduce_sum_f32(self.handle, C.c_void_p(arr.ptr + 4 * offset), count, self.comm_stream)

    def wait(self):
        """Make the compute stream wait for every all-reduce issued with ``allreduce_async``."""
        if self.world == 1 or self.comm_stream is None:
            return
        lib.nm_event_record(self._done, self.comm_stream)
        lib.nm_stream_wait_event(stream(), self._done)

    def broadcast(self, arr, root=0, count=None):
        if self.world == 1:
            return
        lib.nm_comm_broadcast_f32(self.handle, arr.p, arr.size if count is None else count, root, stream())

    def close(self):
        if self.peer is not None:
            lib.nm_dp_destroy(self.peer)
            self.peer = None
        if self.handle is not None:
            lib.nm_comm_destroy(self.handle)
            self.handle = None


def attach(model, comm: Communicator, global_batch: int, peer: bool | None = None):
    """Turn ``model`` into one data-parallel replica: global-batch normalisation + gradient sum over ranks,
    and identical initial weights on every rank (broadcast from rank 0).  ``peer``: True = require the fused
    peer-memory exchange, False = NCCL all-reduce, None = peer memory when the GPUs can map each other."""
    model.comm = comm if comm.world > 1 else None
    set_global_batch(model, global_batch)
    if comm.world > 1:
        comm.broadcast(model.arena.params, 0, model.arena.total)
        if peer is None:
            peer = os.environ.get("NM_DP_PEER", "1") != "0"
        if peer:
            comm.enable_peer(model.arena.total, required=os.environ.get("NM_DP_PEER") == "require")
        from .device import synchronize
        synchronize()
    return model
