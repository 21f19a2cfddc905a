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
B = 4096
n_host = 32
rng = np.random.default_rng(0)
host_x = device.PinnedBuffer((n_host * B, 1, 28, 28), np.float32)
rng.random(out=host_x.array, dtype=np.float32)
labels = rng.integers(0, 10, n_host * B).astype(np.int32)
orig_upload = M._BatchFeeder._upload
stats = {"pageable": 0, "pinned": 0}
def patched(self, step):
    src = self.X[step * self.bs:min((step + 1) * self.bs, self.n)]
    stats["pinned" if device.pinned_address(src) is not None else "pageable"] += 1
    return orig_upload(self, step)
M._BatchFeeder._upload = patched
for trial in range(4):
    for epochs in (1, 3, 6):
        device.synchronize(); t0 = time.perf_counter()
        model.train(host_x.array, labels, epochs=epochs, batch_size=B, verbose=0, stream_batches=True)
        device.synchronize(); dt = time.perf_counter() - t0
        print(f"trial {trial} epochs {epochs}: {1e3 * dt / (epochs * n_host):.3f} ms/step  {stats}", flush=True)
