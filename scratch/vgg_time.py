import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import ref_net
from neuromah_b200 import device, ops
from neuromah_b200._lib import lib
device.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
MODE = sys.argv[2] if len(sys.argv) > 2 else None          # None = FP32-exact, "bf16" = tensor-core plan (tc_plan_gen.py)
spec = ref_net.vgg_cifar_spec()
params = ref_net.init_params(spec, np.random.RandomState(0), init="he")
model = bench.build_model(spec, params, mode=MODE)
rng = np.random.default_rng(0)
x = device.DeviceArray.from_numpy(rng.random((B, 3, 32, 32), dtype=np.float32))
y = device.DeviceArray.from_numpy(rng.integers(0, 10, B).astype(np.int32))
for _ in range(3): model.train_step(x, y)
device.synchronize()
t = time.time()
K = 10
for _ in range(K): model.train_step(x, y)
device.synchronize()
dt = (time.time() - t) / K
print(f"VGG {MODE or 'fp32-exact'} batch {B}: {dt*1e3:.2f} ms/step  {B/dt:.0f} samples/s  {B*913.29e6/dt/1e12:.1f} TFLOP/s")
ops.timers = ops.KernelTimers()
for _ in range(3): model.train_step(x, y)
device.synchronize()
km = ops.timers.collect()
tot = {k: sum(v) / 3 for k, v in km.items()}
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:14]:
    print(f"   {k:40s} {v:8.3f} ms/step")
