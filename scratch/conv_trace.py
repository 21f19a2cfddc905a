import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuromah_b200 import device
from neuromah_b200._lib import lib, _Lib
from neuromah_b200.device import DeviceArray, stream
device.init(0)
dll = lib.load()
dll.nm_tc_debug_set_trace.argtypes = [C.c_void_p]; dll.nm_tc_debug_set_trace_pool.argtypes = [C.c_void_p]
n = 4096
rng = np.random.default_rng(0)
which = sys.argv[1]
trace = DeviceArray.zeros((64, 8), np.int64)
if which == "dgrad":
    a = DeviceArray.from_numpy(rng.integers(0, 2**16, (n, 16, 16, 64)).astype(np.uint16) & 0x3FFF)
    w = DeviceArray.from_numpy(rng.integers(0, 2**14, (32, 9, 64)).astype(np.uint16))
    out = DeviceArray.zeros((n, 16, 16, 32), np.uint16)
    call = lambda: lib.nm_tc_conv3x3_dgrad_bf16(a.p, w.p, out.p, n, 14, 14, 32, 64, stream())
else:
    a = DeviceArray.from_numpy(rng.integers(0, 2**14, (n, 16, 16, 32)).astype(np.uint16))
    w = DeviceArray.from_numpy(rng.integers(0, 2**14, (64, 9, 32)).astype(np.uint16))
    out = DeviceArray.zeros((n, 7, 7, 64), np.uint16)
    idx = DeviceArray.zeros((n, 7, 7, 32), np.uint8)
    call = lambda: lib.nm_tc_conv3x3_fprop_pool_bf16(a.p, w.p, out.p, idx.p, n, 14, 14, 32, 64, 0, stream())
for _ in range(20):
    call()
device.synchronize()
(dll.nm_tc_debug_set_trace if which == 'dgrad' else dll.nm_tc_debug_set_trace_pool)(C.c_void_p(trace.ptr))
call()
device.synchronize()
(dll.nm_tc_debug_set_trace if which == 'dgrad' else dll.nm_tc_debug_set_trace_pool)(None)
t = trace.numpy()
t0 = t[0, 0]
names = ["prod_issue", "mma_tempty", "mma_full", "mma_issued", "epi_tfull", "epi_done"]
print(which, "cycles relative to first producer issue; rows = tile iteration of CTA 0")
print("it   " + " ".join(f"{x:>11s}" for x in names) + "   d(mma_issued)")
prev = None
for i in range(56):
    if t[i, 3] == 0: break
    r = t[i, :6] - t0
    print(f"{i:3d}  " + " ".join(f"{int(x):11d}" for x in r) + (f"   {int(r[3] - prev):8d}" if prev is not None else ""))
    prev = r[3]
