"""Timing experiments: run the bf16 training step with the ablation library (NM_BUILD_ABLATE=1 build) and print the
per-kernel CUDA-event times for a list of debug-flag values.   python scratch/ablate.py 0 1 2 3 ..."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = sys.argv[1]
os.environ["NM_LIB_PATH"] = os.path.join(ROOT, "neuromah_b200", f"libneuromah_b200_abl{FLAGS}.so")
sys.path.insert(0, ROOT)
import numpy as np
import bench
from neuromah_b200 import device, ops
from neuromah_b200._lib import lib
device.init(0)
dll = lib.load()
spec, params = bench.cnn_spec_and_params()
model = bench.build_model(spec, params, mode="bf16")
rng = np.random.default_rng(7)
B = 4096
x = device.DeviceArray.from_numpy(rng.random((4 * B, 1, 28, 28), dtype=np.float32))
y = device.DeviceArray.from_numpy(rng.integers(0, 10, 4 * B).astype(np.int32))
def run(flags, steps=30):
    for s in range(5):
        model.train_step(x[(s % 4) * B:(s % 4 + 1) * B], y[(s % 4) * B:(s % 4 + 1) * B])
    ops.timers = ops.KernelTimers()
    for s in range(steps):
        i = s % 4
        model.train_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B])
    device.synchronize()
    k = ops.timers.collect(); ops.timers = None
    return {n: 1e3 * sum(v) / len(v) for n, v in k.items()}
r = run(FLAGS)
if len(sys.argv) > 2:
    print("flags " + " ".join(f"{n.replace('tc_', '')[:14]:>14s}" for n in r))
print(f"{FLAGS:>5s} " + " ".join(f"{v:14.1f}" for v in r.values()), flush=True)
