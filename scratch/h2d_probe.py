import ctypes as C, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuromah_b200 import device
from neuromah_b200._lib import lib
device.init(0)
n = 64 << 20
pin = device.PinnedBuffer((n,), np.uint8)
pin.array[:] = 1
dev = device.DeviceArray((n,), np.uint8)
e0, e1 = device.new_event(), device.new_event()
ms = C.c_float()
for sz in (n, 12845056, 1 << 20):
    for it in range(4):
        lib.nm_event_record(e0, device.stream())
        lib.nm_h2d(dev.p, C.c_void_p(pin.ptr), sz, device.stream())
        lib.nm_event_record(e1, device.stream())
        device.synchronize()
        lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
        print(f"h2d {sz>>20} MiB it{it}: {ms.value:.3f} ms  {sz/ms.value/1e6:.1f} GB/s", flush=True)
# big pinned alloc like bench
t=time.time(); big = device.PinnedBuffer((25*4096,1,28,28), np.float32); print("alloc 321MB pinned", time.time()-t)
t=time.time(); big.array[:] = 0.5; print("fill", time.time()-t)
d2 = device.DeviceArray((4096,1,28,28), np.float32)
for it in range(6):
    lib.nm_event_record(e0, device.stream())
    lib.nm_h2d(d2.p, C.c_void_p(big.ptr + it*d2.nbytes), d2.nbytes, device.stream())
    lib.nm_event_record(e1, device.stream())
    device.synchronize()
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
    print(f"h2d batch it{it}: {ms.value:.3f} ms  {d2.nbytes/ms.value/1e6:.1f} GB/s", flush=True)
