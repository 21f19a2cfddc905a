#!/usr/bin/env python
"""One training step of the MNIST CNN (bf16 tensor-core plan or FP32-exact mode) bracketed by
cudaProfilerStart/Stop, for `ncu --profile-from-start off` (B200_PROFILING.md).

    python profiles/prof_step.py [--mode bf16|fp32] [--batch 4096] [--steps 1]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/prof_step.py
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import bench
    from neuromah_b200 import device
    from neuromah_b200._lib import lib
    device.init(0)
    spec, params = bench.cnn_spec_and_params()
    model = bench.build_model(spec, params, mode=None if a.mode == "fp32" else "bf16")
    rng = np.random.default_rng(7)
    B = a.batch
    nb = 4
    x = device.DeviceArray.from_numpy(rng.random((nb * B, 1, 28, 28), dtype=np.float32))
    y = device.DeviceArray.from_numpy(rng.integers(0, 10, nb * B).astype(np.int32))
    for s in range(a.warmup):
        model.train_step(x[(s % nb) * B:(s % nb + 1) * B], y[(s % nb) * B:(s % nb + 1) * B])
    device.synchronize()
    lib.nm_profiler_start()
    for s in range(a.steps):
        i = (a.warmup + s) % nb
        model.train_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B])
    device.synchronize()
    lib.nm_profiler_stop()
    print("ok", flush=True)


if __name__ == "__main__":
    main()
