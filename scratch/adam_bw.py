import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neuromah_b200 import device
from neuromah_b200._lib import lib
device.init(0)
n = 1 << 28
s = device.stream()
e0, e1 = device.new_event(), device.new_event()
ms = C.c_float()
for pad in (0, 1 << 15, 264 * 1024 + 4096, 1234568):
    bufs = [device.DeviceArray.zeros((n + k * pad,)) for k in range(4)]
    p, g, m, v = bufs
    print("pad", pad, [hex(b.ptr) for b in bufs])
    for name, bpp, call in (("adam", 28, lambda: lib.nm_adam_step_f32(p.p, g.p, m.p, v.p, n, 1e-3, 0.9, 0.999, 1e-7, 0.1, 0.001, 2, s)),
                            ("rmsprop", 20, lambda: lib.nm_rmsprop_step_f32(p.p, g.p, m.p, n, 1e-3, 0.9, 1e-7, 2, None, s)),
                            ("nadam", 28, lambda: lib.nm_nadam_step_f32(p.p, g.p, m.p, v.p, n, 1e-3, 0.9, 0.999, 1e-8, 0.1, 0.001, 2, s))):
        for _ in range(60): call()
        lib.nm_event_record(e0, s)
        for _ in range(10): call()
        lib.nm_event_record(e1, s)
        device.synchronize()
        lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
        print(f"   {name:8s} {ms.value/10:.4f} ms  {bpp*n/(ms.value/10*1e-3)/1e9:.0f} GB/s")
    del bufs, p, g, m, v
