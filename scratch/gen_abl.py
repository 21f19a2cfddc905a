"""Pace of the general conv kernel with parts switched off (NM_BUILD_ABLATE variants): python scratch/gen_abl.py <bits>"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = sys.argv[1]
if FLAGS != "0":
    os.environ["NM_LIB_PATH"] = os.path.join(ROOT, "neuromah_b200", f"libneuromah_b200_abl{FLAGS}.so")
sys.path.insert(0, ROOT)
import numpy as np
from neuromah_b200 import device
from neuromah_b200._lib import lib
device.init(0)
s = device.stream()
e0, e1 = device.new_event(), device.new_event()
ms = C.c_float()
out = []
for (n, h, cin, cout) in ((1024, 32, 32, 64), (1024, 32, 64, 64), (1024, 16, 64, 128), (1024, 16, 128, 128), (1024, 8, 256, 256)):
    x = device.DeviceArray.zeros((n, h + 2, h + 2, cin), np.uint16)
    w = device.DeviceArray.zeros((cout, 9, cin), np.uint16)
    y = device.DeviceArray.zeros((n, h + 2, h + 2, cout), np.uint16)
    for _ in range(3):
        lib.nm_tc_conv3x3_gen_bf16(x.p, w.p, None, None, y.p, n, h, h, cin, cout, 1, s)
    lib.nm_event_record(e0, s)
    for _ in range(20):
        lib.nm_tc_conv3x3_gen_bf16(x.p, w.p, None, None, y.p, n, h, h, cin, cout, 1, s)
    lib.nm_event_record(e1, s)
    device.synchronize()
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
    out.append(ms.value / 20 * 1e3)
print(f"{FLAGS:>8s} " + " ".join(f"{v:8.1f}" for v in out), flush=True)
