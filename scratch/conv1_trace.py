import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NM_LIB_PATH"] = os.path.join(ROOT, "neuromah_b200", "libneuromah_b200_abl0.so")
sys.path.insert(0, ROOT)
import numpy as np
from neuromah_b200 import device
from neuromah_b200._lib import lib
from neuromah_b200.device import DeviceArray, stream
device.init(0)
dll = lib.load()
dll.nm_tc_debug_set_trace_conv1.argtypes = [C.c_void_p]
n = 4096
rng = np.random.default_rng(0)
x = DeviceArray.from_numpy(rng.random((n, 1, 28, 28), dtype=np.float32))
we = DeviceArray.from_numpy(rng.integers(0, 2**14, (128, 16)).astype(np.uint16))
y = DeviceArray.zeros((n, 16, 16, 32), np.uint16)
idx = DeviceArray.zeros((n, 14, 14, 16), np.uint8)
trace = DeviceArray.zeros((64, 8), np.int64)
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dw, db = DeviceArray.zeros((32, 9)), DeviceArray.zeros((32,))
ws = DeviceArray.zeros((lib.nm_tc_conv1_bwd_workspace(32),), np.uint8)
if which == "fwd":
    call = lambda: lib.nm_tc_conv1_fwd_pool_bf16(x.p, we.p, y.p, idx.p, n, 28, 28, 32, stream())
else:
    call = lambda: lib.nm_tc_conv1_bwd_bf16(x.p, y.p, idx.p, dw.p, db.p, ws.p, ws.nbytes, n, 28, 28, 32, stream())
for _ in range(20): call()
device.synchronize()
dll.nm_tc_debug_set_trace_conv1(C.c_void_p(trace.ptr))
call(); device.synchronize()
dll.nm_tc_debug_set_trace_conv1(None)
t = trace.numpy(); t0 = t[0, 0]
names = ["bld_start", "bld_xfull", "bld_aempty", "bld_arrived", "mma_afull", "epi_tfull", "epi_done"] if which == "fwd" else \
    ["prod_issue", "bld_start", "bld_lfull", "bld_aempty", "bld_arrived", "mma_afull", "mma_issued"]
print("it   " + " ".join(f"{x_:>11s}" for x_ in names))
for i in range(44):
    if t[i, 4] == 0: break
    print(f"{i:3d}  " + " ".join(f"{int(v - t0):11d}" for v in t[i, :7]))
