"""Latency and NVLink bytes of the data-parallel exchange kernels in isolation (torchrun, N ranks; CUDA events on the launch
stream, back-to-back launches, max over ranks).  Slabs: the MNIST CNN's 50 186-float gradient slab (config 2/3) and the
VGG-style stack's 1 186 378 floats (config 4).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 profiles/dp_exchange.py

NVLink bytes are algorithmic: the protocol ships 8-byte {value, step tag} words, no acknowledgements, no reads over the
link.  Small slabs (all-to-all): one word per gradient to each of the world - 1 peers = 8 n (world - 1) bytes per rank and
step.  Large slabs on >= 4 GPUs (two-phase: reduce-scatter to the chunk owner + all-gather of the owner's sums):
16 n (world - 1) / world bytes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)
from neuromah_b200 import device, parallel
from neuromah_b200._lib import lib

device.init(rank)
comm = parallel.Communicator(rank, world)
s = device.stream()
e0, e1 = device.new_event(), device.new_event()
ms = C.c_float()


def run(fn, reps=200):
    for _ in range(20):
        fn()
    device.synchronize(); dist.barrier()
    lib.nm_event_record(e0, s)
    for _ in range(reps):
        fn()
    lib.nm_event_record(e1, s)
    device.synchronize()
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
    t = torch.tensor([ms.value / reps * 1e3], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


chunk = lib.nm_dp_chunk_floats()
for name, n, first in (("MNIST CNN slab", 50186, 320), ("VGG-style slab", 1186378, 1792)):
    comm.peer = None
    assert comm.enable_peer(n, required=True)
    p, g, m, v = (device.DeviceArray.zeros((n,)) for _ in range(4))
    st = device.DeviceArray.zeros((128,), np.uint8)
    lib.nm_adam_state_init(st.p, 1e-3, 0.0, 0.9, 0.999, 1e-7, C.c_longlong(0), s)
    nchunks = (n + chunk - 1) // chunk
    lo = (first + chunk - 1) // chunk
    t_fused = run(lambda: lib.nm_dp_allreduce_adam_dev_f32(comm.peer, p.p, g.p, m.p, v.p, n, st.p, 2, 1, s))

    def split():
        lib.nm_dp_push_f32(comm.peer, g.p, n, lo, nchunks, s)
        lib.nm_dp_collect_adam_dev_f32(comm.peer, p.p, g.p, m.p, v.p, n, st.p, 2, 1, lo, s)
    t_split = run(split)
    t_push = run(lambda: lib.nm_dp_push_f32(comm.peer, g.p, n, lo, nchunks, s)) if False else None   # a push alone would desynchronise the step tags
    t_adam = run(lambda: lib.nm_adam_step_advance_dev_f32(p.p, g.p, m.p, v.p, n, st.p, 2, s))
    t_nccl = run(lambda: lib.nm_comm_allreduce_sum_f32(comm.handle, g.p, n, s))
    comm.check()
    if rank == 0:
        two_phase = world >= 4 and n >= (1 << 18) if os.environ.get("NM_DP_TWO_PHASE") is None else os.environ["NM_DP_TWO_PHASE"] != "0"
        pushed = 16 * n * (world - 1) // world if two_phase else 8 * n * (world - 1)
        print(f"world {world} | {'two-phase (reduce-scatter + all-gather)' if two_phase else 'all-to-all'} | {name}: {n} floats ({4 * n / 1e6:.2f} MB) | fused push+collect+Adam {t_fused:.1f} us | "
              f"early push kernel + collect/Adam kernel (as in the training step, here back to back) {t_split:.1f} us | "
              f"Adam alone {t_adam:.1f} us | NCCL all-reduce alone {t_nccl:.1f} us | NVLink bytes pushed per rank and step "
              f"{pushed / 1e6:.2f} MB ({pushed / 1e3 / max(t_fused, 1e-9):.1f} GB/s per rank while the fused kernel runs)", flush=True)
    lib.nm_dp_destroy(comm.peer)
    comm.peer = None
dist.barrier()
comm.close()
dist.destroy_process_group()
