import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuromah_b200 import device
from neuromah_b200._lib import lib
device.init(0)
B = 12845056
pin = device.PinnedBuffer((8 * B,), np.uint8); pin.array[:] = 1
dev = device.DeviceArray((8 * B,), np.uint8)
s0, s1 = device.new_stream(), device.new_stream()
e0, e1, ej = device.new_event(), device.new_event(), device.new_event()
ms = C.c_float()
def timed(label, fn, reps=40):
    fn(); lib.nm_stream_sync(s0); lib.nm_stream_sync(s1)
    lib.nm_event_record(e0, s0)
    for i in range(reps): fn(i)
    lib.nm_event_record(ej, s1); lib.nm_stream_wait_event(s0, ej)
    lib.nm_event_record(e1, s0)
    lib.nm_stream_sync(s0); lib.nm_stream_sync(s1)
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
    print(f"{label:40s} {ms.value/reps:.4f} ms/batch  {B/(ms.value/reps)/1e6:.1f} GB/s", flush=True)
def one(i=0):
    o = (i % 8) * B
    lib.nm_h2d(C.c_void_p(dev.ptr + o), C.c_void_p(pin.ptr + o), B, s0)
def two(i=0):
    o = (i % 8) * B; h = B // 2
    lib.nm_h2d(C.c_void_p(dev.ptr + o), C.c_void_p(pin.ptr + o), h, s0)
    lib.nm_h2d(C.c_void_p(dev.ptr + o + h), C.c_void_p(pin.ptr + o + h), B - h, s1)
def four(i=0):
    o = (i % 8) * B; q = B // 4
    for k in range(4):
        lib.nm_h2d(C.c_void_p(dev.ptr + o + k * q), C.c_void_p(pin.ptr + o + k * q), q, s0 if k % 2 == 0 else s1)
timed("one 12.8 MB copy per batch", one)
timed("two halves on two streams", two)
timed("four quarters on two streams", four)
timed("one 12.8 MB copy per batch (again)", one)
lab = device.PinnedBuffer((8 * 16384,), np.uint8); dlab = device.DeviceArray((8 * 16384,), np.uint8)
ev = [device.new_event() for _ in range(8)]
def x_and_labels(i=0):
    o = (i % 8) * B
    lib.nm_h2d(C.c_void_p(dev.ptr + o), C.c_void_p(pin.ptr + o), B, s0)
    lib.nm_h2d(C.c_void_p(dlab.ptr + (i % 8) * 16384), C.c_void_p(lab.ptr + (i % 8) * 16384), 16384, s0)
    lib.nm_event_record(ev[i % 8], s0)
timed("12.8 MB + 16 KB labels + event", x_and_labels)
def x_and_labels_first(i=0):
    o = (i % 8) * B
    lib.nm_h2d(C.c_void_p(dlab.ptr + (i % 8) * 16384), C.c_void_p(lab.ptr + (i % 8) * 16384), 16384, s0)
    lib.nm_h2d(C.c_void_p(dev.ptr + o), C.c_void_p(pin.ptr + o), B, s0)
    lib.nm_event_record(ev[i % 8], s0)
timed("16 KB labels + 12.8 MB + event", x_and_labels_first)
