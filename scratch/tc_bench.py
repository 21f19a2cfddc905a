import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
from neuromah_b200._lib import lib
from neuromah_b200.device import DeviceArray, stream, new_event, synchronize
N = 4096
rng = np.random.default_rng(0)
xp = DeviceArray.from_numpy(rng.integers(0, 2**15, (N, 16, 16, 32)).astype(np.uint16) & 0x3FFF | 0x3C00)
dyp = DeviceArray.from_numpy(rng.integers(0, 2**15, (N, 16, 16, 64)).astype(np.uint16) & 0x3FFF | 0x3C00)
w = DeviceArray.from_numpy((rng.standard_normal((64, 32, 3, 3)) * 0.1).astype(np.float32))
wf = DeviceArray((64, 9, 32), np.uint16); wg = DeviceArray((32, 9, 64), np.uint16)
lib.nm_tc_pack_conv_weights_bf16(w.p, wf.p, wg.p, 64, 32, stream())
y = DeviceArray.zeros((N, 7, 7, 64), np.uint16); idx = DeviceArray.zeros((N, 7, 7, 32), np.uint8)
dx = DeviceArray.zeros((N, 16, 16, 32), np.uint16); dw = DeviceArray.zeros((64, 32, 3, 3))
ws = DeviceArray((lib.nm_tc_conv3x3_wgrad_workspace(32, 64),), np.uint8)
flush = DeviceArray.zeros((256 * 1024 * 1024,), np.uint8)
def timeit(name, fn, flops, bytes_):
    e0, e1 = new_event(), new_event(); ms = C.c_float(); ts = []
    for i in range(8):
        flush.fill_zero()
        lib.nm_event_record(e0, stream()); fn(); lib.nm_event_record(e1, stream()); synchronize()
        lib.nm_event_elapsed_ms(e0, e1, C.byref(ms)); ts.append(ms.value)
    t = sorted(ts[2:])[len(ts[2:]) // 2]
    print(f"{name:28s} {t*1e3:8.1f} us  {flops/t/1e9:8.1f} TFLOP/s  {bytes_/t/1e6:8.1f} GB/s (algorithmic)")
F = 2.0 * N * 64 * 32 * 9 * 196
timeit("tc fprop+relu+pool", lambda: lib.nm_tc_conv3x3_fprop_pool_bf16(xp.p, wf.p, y.p, idx.p, N, 14, 14, 32, 64, 0, stream()), F, N*256*64 + N*49*128 + N*49*32)
timeit("tc dgrad", lambda: lib.nm_tc_conv3x3_dgrad_bf16(dyp.p, wg.p, dx.p, N, 14, 14, 32, 64, stream()), F, N*256*128 + N*256*64)
timeit("tc wgrad", lambda: lib.nm_tc_conv3x3_wgrad_bf16(xp.p, dyp.p, dw.p, ws.p, ws.nbytes, N, 14, 14, 32, 64, stream()), F, N*256*128 + N*256*64)
