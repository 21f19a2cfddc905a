"""Where does the end-to-end step lose time against the 0.236 ms H2D floor?  Times every image copy on the copy stream
(CUDA events) inside the real Model.train(stream_batches=True) loop and the gaps between consecutive copies."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from neuromah_b200 import device
from neuromah_b200._lib import lib
import importlib; M = importlib.import_module("neuromah_b200.src.core.Model")
device.init(0)
spec, params = bench.cnn_spec_and_params()
model = bench.build_model(spec, params, mode="bf16")
model.epoch_end_accuracy = False
B, n_host = 4096, 32
rng = np.random.default_rng(0)
host_x = device.PinnedBuffer((n_host * B, 1, 28, 28), np.float32)
rng.random(out=host_x.array, dtype=np.float32)
labels = rng.integers(0, 10, n_host * B).astype(np.int32)
ev = []
real_h2d = lib.nm_h2d
class Wrap:
    def __getattr__(self, k): return getattr(lib, k)
    def nm_h2d(self, dst, src, nbytes, st):
        if nbytes > (1 << 20) and record[0]:
            a, b = device.new_event(), device.new_event()
            lib.nm_event_record(a, st); real_h2d(dst, src, nbytes, st); lib.nm_event_record(b, st)
            ev.append((a, b))
        else:
            real_h2d(dst, src, nbytes, st)
record = [False]
M.lib = Wrap()
for _ in range(2):
    model.train(host_x.array, labels, epochs=1, batch_size=B, verbose=0, stream_batches=True)
record[0] = True
device.synchronize(); t0 = time.perf_counter()
model.train(host_x.array, labels, epochs=4, batch_size=B, verbose=0, stream_batches=True)
device.synchronize(); dt = time.perf_counter() - t0
ms = C.c_float()
dur, gap = [], []
for i, (a, b) in enumerate(ev):
    lib.nm_event_elapsed_ms(a, b, C.byref(ms)); dur.append(ms.value)
    if i:
        lib.nm_event_elapsed_ms(ev[i - 1][1], a, C.byref(ms)); gap.append(ms.value)
dur, gap = np.array(dur), np.array(gap)
print(f"wall {1e3 * dt / (4 * n_host):.4f} ms/step over {len(ev)} copies")
print(f"copy duration: mean {dur.mean():.4f} median {np.median(dur):.4f} p90 {np.quantile(dur, .9):.4f} max {dur.max():.4f} ms")
print(f"gap end->next start: mean {gap.mean():.4f} median {np.median(gap):.4f} p90 {np.quantile(gap, .9):.4f} max {gap.max():.4f} ms")
