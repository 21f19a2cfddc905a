import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuromah_b200 import device
from neuromah_b200._lib import lib
from neuromah_b200.device import DeviceArray, stream
device.init(0)
n = 4096
dy = DeviceArray.zeros((n, 16, 16, 64), np.uint16)
wg = DeviceArray.zeros((32, 9, 64), np.uint16)
dx = DeviceArray.zeros((n, 16, 16, 32), np.uint16)
e0, e1 = device.new_event(), device.new_event()
ms = C.c_float()
for it in range(6):
    lib.nm_event_record(e0, stream())
    lib.nm_tc_conv3x3_dgrad_bf16(dy.p, wg.p, dx.p, n, 14, 14, 32, 64, stream())
    lib.nm_event_record(e1, stream())
    device.synchronize()
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
print(os.environ.get("NM_TC_DEBUG", "0"), os.environ.get("NM_TC_DGRAD_FOLD", "1"), "dgrad us:", round(ms.value * 1e3, 1))
