"""One eager (no CUDA graph) VGG-style training step (bf16, or bf16x3 as second argument) between cudaProfilerStart/Stop, for an
ncu launch list:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/prof_vgg_step.py 1024 [bf16x3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import ref_net
from neuromah_b200 import device
from neuromah_b200._lib import lib
device.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spec = ref_net.vgg_cifar_spec()
params = ref_net.init_params(spec, np.random.RandomState(0), init="he")
model = bench.build_model(spec, params, mode=sys.argv[2] if len(sys.argv) > 2 else "bf16")
model.graph_steps = False
rng = np.random.default_rng(0)
x = device.DeviceArray.from_numpy(rng.random((B, 3, 32, 32), dtype=np.float32))
y = device.DeviceArray.from_numpy(rng.integers(0, 10, B).astype(np.int32))
for _ in range(2): model.train_step(x, y)
device.synchronize()
lib.nm_profiler_start()
model.train_step(x, y)
device.synchronize()
lib.nm_profiler_stop()
print("done")
