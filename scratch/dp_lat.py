"""Latency of the fused exchange + Adam kernel in isolation (torchrun, N ranks)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)
from neuromah_b200 import device, parallel
from neuromah_b200._lib import lib
device.init(rank)
comm = parallel.Communicator(rank, world)
n = 50186
assert comm.enable_peer(n, required=True)
p, g, m, v = (device.DeviceArray.zeros((n,)) for _ in range(4))
st = device.DeviceArray.zeros((128,), np.uint8)
lib.nm_adam_state_init(st.p, 1e-3, 0.0, 0.9, 0.999, 1e-7, C.c_longlong(0), device.stream())
s = device.stream()
e0, e1 = device.new_event(), device.new_event()
ms = C.c_float()
def run(fn, reps=300):
    for _ in range(20): fn()
    device.synchronize(); dist.barrier()
    lib.nm_event_record(e0, s)
    for _ in range(reps): fn()
    lib.nm_event_record(e1, s)
    device.synchronize()
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
    return ms.value / reps * 1e3
t_dp = run(lambda: lib.nm_dp_allreduce_adam_dev_f32(comm.peer, p.p, g.p, m.p, v.p, n, st.p, 2, 1, s))
t_ar = run(lambda: lib.nm_dp_allreduce_sum_f32(comm.peer, g.p, n, s))
t_adam = run(lambda: lib.nm_adam_step_advance_dev_f32(p.p, g.p, m.p, v.p, n, st.p, 2, s))
t_nccl = run(lambda: lib.nm_comm_allreduce_sum_f32(comm.handle, g.p, n, s))
comm.check()
if rank == 0:
    print(f"world {world}: fused exchange+Adam {t_dp:.2f} us | exchange only {t_ar:.2f} us | Adam only {t_adam:.2f} us | NCCL all-reduce {t_nccl:.2f} us  (back-to-back launches, 50186 floats)")
dist.barrier(); comm.close(); dist.destroy_process_group()
