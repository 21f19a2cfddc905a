"""Role timeline (clock64, CTA 0) of the general conv kernel: python scratch/gen_trace.py [cin cout h]   (needs the abl0 build)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NM_LIB_PATH"] = os.path.join(ROOT, "neuromah_b200", "libneuromah_b200_abl0.so")
sys.path.insert(0, ROOT)
import numpy as np
from neuromah_b200 import device
from neuromah_b200._lib import lib
device.init(0)
dll = lib.load()
dll.nm_tc_debug_set_trace_gen.argtypes = [C.c_void_p]
cin, cout, h = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 64, 32)
n = 1024
s = device.stream()
x = device.DeviceArray.zeros((n, h + 2, h + 2, cin), np.uint16)
w = device.DeviceArray.zeros((cout, 9, cin), np.uint16)
y = device.DeviceArray.zeros((n, h + 2, h + 2, cout), np.uint16)
trace = device.DeviceArray.zeros((48, 16), np.int64)
call = lambda: lib.nm_tc_conv3x3_gen_bf16(x.p, w.p, None, None, y.p, n, h, h, cin, cout, 1, s)
for _ in range(5): call()
device.synchronize()
dll.nm_tc_debug_set_trace_gen(C.c_void_p(trace.ptr))
call(); device.synchronize()
dll.nm_tc_debug_set_trace_gen(None)
t = trace.numpy()
t0 = t[0, 2]
names = ["prod_A", "-", "mma_start", "mma_tempty", "mma_afull", "mma_commit", "epi_wait", "epi_tfull", "epi_done"]
print(f"conv {cin}->{cout} @{h}: it  " + " ".join(f"{x_:>10s}" for x_ in names))
for i in range(20):
    print(f"{i:3d}  " + " ".join(f"{int(v - t0) if v else 0:10d}" for v in t[i, :9]))
