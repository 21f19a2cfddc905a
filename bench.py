#!/usr/bin/env python
"""bench.py -- training throughput of the Neuromah hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5                         # configs[1]: MNIST CNN, batch 4096 per GPU
    python bench.py --workload vgg --steps 20 --warmup 5                   # configs[3]: VGG-style stack, batch 1024 per GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps K --warmup W         # CPU arm (the reference's NumPy path)

A "step" is one pass of the hot path over one batch: forward, softmax-CE loss sums, backward,
(gradient exchange when N > 1), Adam applied twice per layer.  Workloads:
  mnist (default) = BASELINE.json configs[1]: the CNN of examples/test_Conv2D_MNIST.py:47-53 at batch 4096 per GPU
  vgg             = BASELINE.json configs[3]: C64 C64 P C128 C128 P C256 C256 P FC10 on 32x32x3 inputs, 1024 per GPU
weak scaling (global batch = per-GPU batch x N), synthetic data, random-init weights.

One JSON line on stdout (rank 0):
  value        device-resident throughput (inputs already in HBM), K graph-replayed steps between two CUDA events
  e2e          the same metric through the public API (Model.train(..., stream_batches=True)) from page-locked HOST
               buffers: per-step host->device copy of the batch (raw uint8 pixels, normalised on the device -- the
               example scripts' `X.astype(float32) / 255`) and a per-step device->host read of the running loss sums
  roofline     the dominant kernel (CUDA events on the launch stream in a second timed region) against the measured peak
  parity       correctness AT THIS CONFIGURATION, checked after the timed regions: every rank's parameters bit-identical,
               and the data-parallel step equal to a single-GPU run over the same global batch in virtual shards
  fp32_exact   (N = 1) the FP32-exact mode's throughput on the same workload (+ fp32_exact_tensor_core: mode "bf16x3")
  cpu_baseline (N = 1) the reference's CPU path timed on this box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "samples/s"

# ----------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------


def cnn_spec_and_params(seed=0):
    """The MNIST CNN of examples/test_Conv2D_MNIST.py:47-53 as a layer list + RandomNormal(0.01) initial weights
    (Initializer.py:90-98) in the reference's shapes.  Self-contained on purpose: the product arm of the bench does not
    touch oracle/ (only the CPU arm below does)."""
    spec = [
        {"kind": "conv", "ic": 1, "oc": 32, "k": 3, "padding": "same", "relu": True},
        {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"},
        {"kind": "conv", "ic": 32, "oc": 64, "k": 3, "padding": "same", "relu": True},
        {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"},
        {"kind": "flatten"},
        {"kind": "dense", "n_in": 64 * 7 * 7, "n_out": 10, "relu": False},
        {"kind": "softmax"},
    ]
    return spec, _init(spec, seed, "random_normal")


def vgg_spec_and_params(seed=0):
    """BASELINE.json configs[3] as SURVEY.md 8(d) defines it: all 3x3 'same' + ReLU, He init (Initializer.py:117-128)."""
    spec, ic = [], 3
    for oc in (64, 128, 256):
        spec += [{"kind": "conv", "ic": ic, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "conv", "ic": oc, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"}]
        ic = oc
    spec += [{"kind": "flatten"}, {"kind": "dense", "n_in": 256 * 4 * 4, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    return spec, _init(spec, seed, "he")


def mlp_spec_and_params(seed=0):
    """BASELINE.json configs[0]: the clean 784-128-10 MLP (Dense+ReLU, Dense, Softmax; examples/test_Dense_MNIST.py), RandomNormal init."""
    spec = [{"kind": "dense", "n_in": 784, "n_out": 128, "relu": True}, {"kind": "dense", "n_in": 128, "n_out": 10, "relu": False},
            {"kind": "softmax"}]
    return spec, _init(spec, seed, "random_normal")


def _init(spec, seed, init):
    rng = np.random.RandomState(seed)
    params = []
    for L in spec:
        if L["kind"] == "conv":
            shape, fan_in, b = (L["oc"], L["ic"], L["k"], L["k"]), L["ic"] * L["k"] * L["k"], np.zeros((L["oc"], 1))
        elif L["kind"] == "dense":
            shape, fan_in, b = (L["n_in"], L["n_out"]), L["n_in"], np.zeros((1, L["n_out"]))
        else:
            params.append(None)
            continue
        w = rng.normal(0, np.sqrt(2.0 / fan_in), size=shape) if init == "he" else rng.randn(*shape) * 0.01
        params.append({"weights": w, "biases": b})
    return params


WORKLOADS = {
    "mnist": dict(metric="train samples/sec (Conv2D MNIST CNN)", batch=4096, shape=(1, 28, 28), make=cnn_spec_and_params,
                  flops_per_sample=22.767e6,          # BASELINE.md section 3 (first-layer dgrad excluded)
                  desc="MNIST CNN (Conv3x3x32-ReLU-Pool2-Conv3x3x64-ReLU-Pool2-Dense10-Softmax), softmax-CE, "
                       "Adam x2 per layer (reference quirk Q1)",
                  cpu_sample=4096, cpu_steps=5, ref_budget=200_000),
    "vgg": dict(metric="train samples/sec (VGG-style Conv2D stack, CIFAR-10-shaped)", batch=1024, shape=(3, 32, 32),
                make=vgg_spec_and_params, flops_per_sample=913.29e6,
                desc="VGG-style stack C64 C64 P C128 C128 P C256 C256 P FC10 (3x3 'same' + ReLU, He init) on 32x32x3, "
                     "softmax-CE, Adam x2 per layer (reference quirk Q1)",
                cpu_sample=64, cpu_steps=4, ref_budget=4_000),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (launch-bound on a GPU: 0.078 GFLOP per step)
    "mlp": dict(metric="train samples/sec (Dense MNIST MLP 784-128-10)", batch=128, shape=(784,), make=mlp_spec_and_params,
                flops_per_sample=0.610e6,
                desc="MLP 784-128-10 (Dense+ReLU, Dense, Softmax), softmax-CE, Adam x2 per layer (reference quirk Q1), batch 128",
                cpu_sample=128, cpu_steps=200, ref_budget=400_000),
    # BASELINE.json configs[4]: one FedAvg client per GPU; a "step" of this workload is one ROUND = `local_steps` local
    # training steps at batch 4096 (no gradient exchange) + federated_average of the weight slab over NCCL
    "fedavg": dict(metric="train samples/sec (FedAvg rounds, Conv2D MNIST CNN, one client per GPU)", batch=4096, shape=(1, 28, 28),
                   make=cnn_spec_and_params, flops_per_sample=22.767e6, local_steps=8,
                   desc="FedAvg round = 8 local steps of the MNIST CNN at batch 4096 per client (Adam x2, state stays local) + "
                        "federated_average (Federated/utils.py:20-23) of the weight slab as a pre-scaled SUM all-reduce over NCCL",
                   cpu_sample=4096, cpu_steps=8, ref_budget=200_000),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the committed `ncu --set full`
    capture, as summarised in profiles/ncu_traffic.json by profiles/summarize.py (None when there is no capture)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(path)).get(name)
    except Exception:
        return None, None
    if not rec:
        return None, None
    return rec["dram_bytes"], f"{rec['source']} ({rec.get('note', 'ncu --set full --clock-control none')})"


_TAG = re.compile(r"\[(\d+)->(\d+)@(\d+)\]")
_DTAG = re.compile(r"tc_dense_gen_\w+\[(\d+)->(\d+)\]")


def kernel_work(name, b):
    """(algorithmic FLOPs, algorithmic HBM bytes) of one launch of the named operator at batch b.
    Operators of the FP32-exact tensor-core mode (names with _x3) move three bf16 planes per tensor: 3 x the bf16 bytes; their
    FLOPs are counted once (FP32-equivalent), although the tensor cores execute six bf16 products per term."""
    f, by = _kernel_work(name, b)
    return (f, 3.0 * by) if "_x3" in name else (f, by)


def _kernel_work(name, b):
    c2 = 2.0 * b * 64 * 32 * 9 * 14 * 14          # conv fprop = dgrad = wgrad = 2 B OC IC kH kW oh ow
    c1 = 2.0 * b * 32 * 1 * 9 * 28 * 28
    m = _DTAG.search(name)
    if m:                                         # general Dense GEMMs, tagged [n_in->n_out]: bf16 operands in, bf16 / fp32 out
        n_in, n_out = int(m.group(1)), int(m.group(2))
        flops = 2.0 * b * n_in * n_out
        if "wgrad" in name:
            return flops, 2.0 * b * (n_in + n_out) + 4.0 * n_in * n_out
        return flops, 2.0 * b * (n_in + n_out) + 2.0 * n_in * n_out
    m = _TAG.search(name)
    if m:                                         # general conv kernels, tagged [ic->oc@h]
        ic, oc, h = int(m.group(1)), int(m.group(2)), int(m.group(3))
        icp, ocp, grid = -(-ic // 64) * 64, -(-oc // 64) * 64, (h + 2) * (h + 2)
        flops = 2.0 * b * oc * ic * 9 * h * h
        if "fprop" in name:
            return flops, b * grid * 2.0 * (icp + ocp)
        if "dgrad" in name:
            return flops, b * grid * 2.0 * (ocp + 2 * icp)          # dY + ReLU mask of the layer below + dX
        if "wgrad" in name:
            return flops, b * grid * 2.0 * (icp + ocp)
    table = {
        # bf16 plan: activations bf16 NHWC with halo (256 grid pixels / image for the 14x14 maps)
        "tc_conv2_fprop_pool": (c2, b * (256 * 64 + 49 * 128 + 49 * 32) + 64 * 288 * 2),
        "tc_conv2_bwd": (2 * c2, b * (256 * 128 + 256 * 64 + 256 * 64) + 64 * 288 * 6),
        "tc_conv1_fwd_pool": (c1, b * (784 * 4 + 196 * 64 + 196 * 16)),
        "tc_conv1_bwd": (c1 / 4, b * (784 * 4 + 196 * 64 + 196 * 16)),
        "tc_dense_softmax_ce": (2.0 * b * 3136 * 10, b * (3136 * 2 + 100)),
        "tc_dense_wgrad": (2.0 * b * 3136 * 10, b * (3136 * 2 + 40)),
        "tc_dense_bwd_unpool": (2.0 * b * 3136 * 10, b * (49 * 32 + 196 * 128 + 40)),
        # FP32-exact mode: NCHW float32
        "conv2d_fprop_f32[32->64]": (c2, b * 4 * (32 * 196 + 64 * 196)),
        "conv2d_dgrad_f32[32->64]": (c2, b * 4 * (32 * 196 + 64 * 196)),
        "conv2d_wgrad_f32[32->64]": (c2, b * 4 * (32 * 196 + 64 * 196)),
        "conv2d_fprop_f32[1->32]": (c1, b * 4 * (784 + 32 * 784)),
        "conv2d_wgrad_f32[1->32]": (c1, b * 4 * (784 + 32 * 784)),
        "adam_step_f32": (0.0, 50186 * 28),
    }
    return table.get(name, (0.0, 0.0))


def build_model(spec, params, lr=0.001, mode=None):
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Conv2D, Layer_MaxPooling2D, Layer_Dense, Layer_Flatten
    from neuromah_b200.src.activations import Activation_ReLU, Activation_Softmax
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.optimizers import Optimizer_Adam
    from neuromah_b200.src.metrics import Accuracy_Categorical
    model = Model()
    for L, p in zip(spec, params):
        k = L["kind"]
        if k == "conv":
            layer = Layer_Conv2D(L["ic"], L["oc"], L["k"], padding=L["padding"],
                                 activation=Activation_ReLU() if L.get("relu") else None)
        elif k == "pool":
            layer = Layer_MaxPooling2D(L["size"], strides=L.get("stride"), padding=L.get("padding", "valid"))
        elif k == "flatten":
            layer = Layer_Flatten()
        elif k == "dense":
            layer = Layer_Dense(L["n_in"], L["n_out"], activation=Activation_ReLU() if L.get("relu") else None)
        else:
            layer = Activation_Softmax()
        if p is not None:
            layer.set_parameters(p["weights"], p["biases"])
        model.add(layer)
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=Optimizer_Adam(learning_rate=lr),
              accuracy=Accuracy_Categorical)
    model.finalize(mode=mode)
    return model


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md).

    The sampler is started at program start (nvidia-smi needs a second or two to come up on a fresh box) and
    every line carries a timestamp; ``stop(t0, t1)`` keeps the samples that fall inside the timed window
    (widened by ``slack`` seconds when the window is shorter than the sampling period)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=50):
        self.gpu, self.proc, self.path, self.period_ms = gpu_index, None, f"/tmp/nm_clocks_{os.getpid()}.csv", period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0, t1):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(2.5 * self.period_ms / 1e3)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), float(f[4]) if f[4] not in ("", "[N/A]") else 0.0,
                             [n for n, v in zip(names, f[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        try:
            os.remove(self.path)
        except OSError:
            pass
        window = "timed region"
        inside = [r for r in rows if t0 <= r[0] <= t1]
        if len(inside) < 2:
            slack = 4 * self.period_ms / 1e3
            inside = [r for r in rows if t0 - slack <= r[0] <= t1 + slack]
            window = f"timed region +/- {slack:.2f} s (region shorter than the sampling period)"
        reasons = sorted({n for r in inside for n in r[4]})
        return {"sm_mhz": statistics.median([r[1] for r in inside]) if inside else None,
                "sm_max_mhz": max([r[2] for r in inside]) if inside else (max([r[2] for r in rows]) if rows else None),
                "power_w_max": max([r[3] for r in inside]) if inside else None,
                "samples": len(inside), "window": window, "reasons": reasons}


def optimizer_bandwidth(n=1 << 28, reps=10):
    """the MNIST-CNN parameter slab (0.2 MB) is far too small to say anything about the optimizer
    kernels' bandwidth, so the same kernels are also timed on a synthetic slab of 2^28 parameters (1 GiB per buffer)."""
    import ctypes as C
    from neuromah_b200 import device
    from neuromah_b200._lib import lib
    peak = peaks()["hbm_gbs"]
    bufs = [device.DeviceArray.zeros((n,)) for _ in range(4)]
    p, g, m, v = bufs
    lib.nm_memset(g.p, 0x3c, g.nbytes, device.stream())          # small non-zero gradients (0x3c3c3c3c ~ 0.0115)
    s = device.stream()
    cases = {
        "adam_step_f32": (28, lambda: lib.nm_adam_step_f32(p.p, g.p, m.p, v.p, n, 1e-3, 0.9, 0.999, 1e-7, 0.1, 0.001, 2, s)),
        "sgd_step_f32": (12, lambda: lib.nm_sgd_step_f32(p.p, g.p, None, n, 1e-3, 0.0, 2, None, s)),
        "sgd_momentum_step_f32": (20, lambda: lib.nm_sgd_step_f32(p.p, g.p, m.p, n, 1e-3, 0.9, 2, None, s)),
        "rmsprop_step_f32": (20, lambda: lib.nm_rmsprop_step_f32(p.p, g.p, m.p, n, 1e-3, 0.9, 1e-7, 2, None, s)),
        "nadam_step_f32": (28, lambda: lib.nm_nadam_step_f32(p.p, g.p, m.p, v.p, n, 1e-3, 0.9, 0.999, 1e-8, 0.1, 0.001, 2, s)),
    }
    e0, e1 = device.new_event(), device.new_event()
    out = {"params": n, "peak_gbs": peak, "n_apply": 2}
    ms = C.c_float()
    for name, (bpp, call) in cases.items():
        for _ in range(3):
            call()
        lib.nm_event_record(e0, s)
        for _ in range(reps):
            call()
        lib.nm_event_record(e1, s)
        device.synchronize()
        lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
        gbs = bpp * n / (ms.value / reps * 1e-3) / 1e9
        out[name] = {"bytes_per_param": bpp, "ms": round(ms.value / reps, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
    del bufs
    return out


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own NumPy path on this box's host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_arm(wl, steps, warmup, sample):
    """(samples/s, ms_per_step, cores, kind, description).  Where the reference tree is present (the authoring container)
    the UNMODIFIED reference Python is stepped exactly like Model.train (Model.py:154-174) through oracle/ref_shim.py;
    on the GPU box (no /root/reference) the oracle port runs the same arithmetic.  Either way conv goes through the NumPy
    im2col + GEMM restatement of CPU_kernels/Conv2D.cpp (needs Eigen: unbuildable) and max-pool through the reference's
    own C++ (oracle/_ref) when it was built.  All host threads NumPy's BLAS will use."""
    from oracle import ref_net, ref_shim
    spec, params = wl["make"]()
    rng = np.random.default_rng(1)
    x = rng.random((sample,) + wl["shape"], dtype=np.float32)
    y = np.eye(10)[rng.integers(0, 10, sample)]
    if ref_shim.available():
        model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.001)
        step = lambda: ref_shim.train_step(model, x, y)      # noqa: E731
        kind, what = "reference", "the UNMODIFIED reference Python (Model / Dense / ReLU / softmax-CE / Adam) via oracle/ref_shim.py"
        pool = "the reference's compiled max-pool C++"
    else:
        net = ref_net.RefNet(spec, params)
        step = lambda: net.step(x, y)                        # noqa: E731
        kind, what = "port", "oracle port: reference Dense/ReLU/softmax-CE/Adam arithmetic in NumPy (float64)"
        pool = "reference C++ max-pool (oracle/_ref)" if net.pool is not None else "NumPy max-pool"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    desc = (f"{steps} steps x {sample} samples (the configured batch is {wl['batch']}) of the step of: {wl['desc']}; {what}, "
            f"conv = NumPy im2col+GEMM restatement of Conv2D.cpp (float32), {pool}")
    return steps * sample / dt, 1e3 * dt / steps, os.cpu_count(), kind, desc


def run_reference(args, wl, rank, world, out_fd):
    if rank != 0:
        return
    warm = max(args.warmup, 1)
    sample = max(16, min(wl["batch"], wl["ref_budget"] // (args.steps + warm)))      # K steps must end within a few minutes
    value, ms, cores, kind, desc = cpu_arm(wl, args.steps, warm, sample)
    line = {"metric": wl["metric"], "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "batch_per_gpu": args.batch or wl["batch"],
                       "global_batch": (args.batch or wl["batch"]) * args.gpus,
                       "parallelism": f"dp{args.gpus}", "mode": "reference CPU path (rank 0's host cores only)",
                       "per_step_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line, out_fd)


def _emit(line, out_fd):
    os.write(out_fd, (json.dumps(line) + "\n").encode())


# ----------------------------------------------------------------------------------------------------------------
# parity at the benchmarked configuration
# ----------------------------------------------------------------------------------------------------------------
def parity_check(wl, mode, B, rank, world, comm_factory, dist, peer_arg=None):
    """One training step at the benchmarked batch on fresh models, after the timed regions:
      * data parallel (this run's exchange) on rank r's shard of a seeded global batch of B x world samples;
      * the same global batch on ONE GPU in `world` virtual shards with the global batch as divisor, the
        gradient slabs summed on the host in float64, then the same Adam step.
    Reports the per-tensor relative L2 error of the summed gradient, the parameter difference after the update and whether
    every rank holds bit-identical parameters (crc32 of the slab gathered over the bootstrap group).  At N = 1 the global
    batch is split into two virtual shards against the un-split step (linearity of the gradient in the batch)."""
    from neuromah_b200 import device, parallel
    spec, params = wl["make"](seed=3)
    shards = max(world, 2)
    b = B if world > 1 else B // 2
    rng = np.random.default_rng(777)
    x = rng.random((shards * b,) + wl["shape"], dtype=np.float32)
    y = rng.integers(0, 10, shards * b).astype(np.int32)
    lr = 1e-3
    # (a) the run's own path
    m_dp = build_model(spec, params, lr=lr, mode=mode)
    comm = comm_factory()
    parallel.attach(m_dp, comm, shards * b, peer=peer_arg)
    lo, hi = (rank * b, (rank + 1) * b) if world > 1 else (0, shards * b)
    m_dp.train_step(x[lo:hi], y[lo:hi])
    device.synchronize()
    g_dp = m_dp.arena.host_grads().astype(np.float64)
    p_dp = m_dp.arena.host_params()
    # (b) virtual shards on this GPU
    m_v = build_model(spec, params, lr=lr, mode=mode)
    parallel.set_global_batch(m_v, shards * b)
    acc = np.zeros(m_v.arena.total, np.float64)
    for r in range(shards):
        out = m_v.forward(x[r * b:(r + 1) * b], training=True)
        m_v.backward(out, y[r * b:(r + 1) * b])
        acc += m_v.arena.grads.numpy()[:m_v.arena.total]
    m_v.arena.grads.copy_from(acc.astype(np.float32))
    m_v._optimizer_step()
    device.synchronize()
    g_v = np.concatenate([acc[o:o + s] for _, _, o, s, _ in m_v.arena.table])
    p_v = m_v.arena.host_params()
    off, per_tensor = 0, []
    for (_, _, _, size, _) in m_dp.arena.table:
        a, r_ = g_dp[off:off + size], g_v[off:off + size]
        per_tensor.append(float(np.linalg.norm(a - r_) / max(np.linalg.norm(r_), 1e-30)))
        off += size
    crc = zlib.crc32(p_dp.tobytes())
    identical = True
    if dist is not None:
        import torch
        t = torch.tensor([crc], dtype=torch.int64, device="cuda")
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        identical = len({int(o.item()) for o in outs}) == 1
    dp_abs = float(np.abs(p_dp - p_v).max())
    init = _flat_init(m_v, params)
    upd = float(np.linalg.norm((p_dp - p_v).astype(np.float64)) / max(np.linalg.norm((p_v - init).astype(np.float64)), 1e-30))
    comm.close()
    grad_tol = 2e-3 if mode == "bf16" else 1e-4          # same addends, different association (+ bf16 re-rounding of partials)
    ok = identical and max(per_tensor) <= grad_tol and dp_abs <= 2 * (1 + 1.9 / np.sqrt(1.999)) * lr * 1.003
    return {"ok": bool(ok), "replicas_bit_identical": bool(identical), "ranks": world, "virtual_shards": shards,
            "batch_per_shard": b, "grad_rel_l2_max": max(per_tensor), "grad_rel_l2_tol": grad_tol,
            "params_max_abs_diff": dp_abs, "update_rel_l2": upd,
            "what": ("data-parallel step (this run's gradient exchange) vs one GPU stepping the same global batch in "
                     "virtual shards" if world > 1 else "un-split step vs two virtual shards with the global batch as divisor")}


def fedavg_parity(model, comm, rank, world, dist, B, E, x, y):
    """One more round on the timed model: local steps, then the device average; every client's pre-average weights are
    gathered over the bootstrap group and the reference rule sum_i w_i n_i / N (Federated/utils.py:20-23, equal n_i) is
    applied on the host in float64.  Every rank must end with that model, bit-identical across ranks."""
    from neuromah_b200 import device
    from neuromah_b200.Federated.utils import federated_average, federated_average_device
    for _ in range(E):
        model.train_step(x, y)
    device.synchronize()
    mine = model.arena.host_params()
    allw = [mine]
    if dist is not None:
        import torch
        t = torch.from_numpy(mine).cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        allw = [o.cpu().numpy() for o in outs]
    ref = np.asarray(federated_average([w.astype(np.float64) for w in allw], [E * B] * world))
    federated_average_device(model.arena, E * B, E * B * world, comm if world > 1 else None)
    device.synchronize()
    got = model.arena.host_params()
    crc = zlib.crc32(got.tobytes())
    identical = True
    if dist is not None:
        import torch
        t = torch.tensor([crc], dtype=torch.int64, device="cuda")
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        identical = len({int(o.item()) for o in outs}) == 1
    err = float(np.abs(got - ref).max())
    spread = float(max(np.abs(w - allw[0]).max() for w in allw))
    return {"ok": bool(identical and err <= 1e-6), "replicas_bit_identical": bool(identical), "ranks": world,
            "max_abs_err_vs_reference_rule": err, "client_weight_spread_before_average": spread,
            "what": "device FedAvg (pre-scaled SUM all-reduce) vs federated_average() on the gathered client weights"}


def _flat_init(model, params):
    """Initial parameters in arena table order (the update vector is measured against them)."""
    flat = []
    trainable = [l for l in model.layers if hasattr(l, "weights")]
    ps = [p for p in params if p is not None]
    for layer, p in zip(trainable, ps):
        flat += [np.asarray(p["weights"], np.float32).ravel(), np.asarray(p["biases"], np.float32).ravel()]
    return np.concatenate(flat)


# ----------------------------------------------------------------------------------------------------------------
def main():
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (e.g. NCCL's version banner, which goes
    # to stdout) is sent to stderr for the duration of the run
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mnist", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (weak scaling) or global batch (--scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, BASELINE.json): --batch samples per GPU; strong: --batch is the GLOBAL batch, split N ways")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "bf16x3", "fp32"],
                    help="bf16 = fused tensor-core plan (tcgen05), bf16x3 = FP32-exact on the tensor cores (3 bf16 planes per operand), "
                         "fp32 = FP32-exact CUDA-core kernels")
    ap.add_argument("--input", default="u8", choices=["u8", "f32"],
                    help="host input of the end-to-end leg: raw uint8 pixels normalised on the device, or float32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32_exact / hbm_kernels / float32-input e2e sub-records")
    ap.add_argument("--dp", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: gradient exchange -- fused peer-memory all-reduce + Adam kernel, or NCCL all-reduce "
                         "(auto = peer memory when the GPUs can map each other)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = {"mnist": 1000, "vgg": 100, "fedavg": 100, "mlp": 2000}[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank, world, out_fd)

    import ctypes as C
    from neuromah_b200 import device, ops, parallel
    from neuromah_b200._lib import lib

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    clocks = ClockSampler(local_rank, period_ms=100)
    if rank == 0:                       # rank 0's GPU only: concurrent nvidia-smi pollers contend for the driver lock
        clocks.start()                  # early: nvidia-smi takes a while to produce its first sample
    device.init(local_rank)
    W, K, B = max(args.warmup, 3), args.steps, args.batch or wl["batch"]
    if args.scaling == "strong":
        if B % world:
            raise SystemExit(f"--scaling strong: global batch {B} is not divisible by {world} GPUs")
        B //= world
    mode = None if args.mode == "fp32" else args.mode

    def barrier():
        device.synchronize()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    spec, params = wl["make"]()
    model = build_model(spec, params, mode=mode)
    comm = parallel.Communicator(rank, world)
    peer_arg = None if args.dp == "auto" else args.dp == "peer"
    fed = args.workload == "fedavg"
    E = wl.get("local_steps", 1)         # training steps per timed "step" (a FedAvg round is E local steps + the average)
    if fed:                              # clients train independently from the same global model; only weights are exchanged
        if world > 1:
            comm.broadcast(model.arena.params, 0, model.arena.total)
    else:
        parallel.attach(model, comm, B * world, peer=peer_arg)
    model.epoch_end_accuracy = False     # time the K training steps only (no Model.py:199 sweep)
    from neuromah_b200.Federated.utils import federated_average_device

    def fed_average():
        federated_average_device(model.arena, E * B, E * B * world, comm if world > 1 else None)

    # synthetic data: up to 32 distinct host batches per rank in page-locked memory (cycled when K > 32).  Raw pixels are
    # uint8; the float32 copy is the example scripts' host-side normalisation of the same pixels.
    n_host = min(max(K * E, 1), 8 if args.workload == "vgg" else 32)
    rng = np.random.default_rng(100 + rank)
    host_u8 = device.PinnedBuffer((n_host * B,) + wl["shape"], np.uint8)
    host_u8.array[...] = rng.integers(0, 256, host_u8.array.shape, dtype=np.uint8)
    labels = rng.integers(0, 10, n_host * B).astype(np.int32)

    # ---------------- device-resident leg: inputs already in HBM (float32, as the reference's scripts hold them) ------
    n_dev = min(n_host, 8)
    x_dev = device.DeviceArray.from_numpy(host_u8.array[:n_dev * B].astype(np.float32) / 255.0)
    y_dev = device.DeviceArray.from_numpy(labels[:n_dev * B])
    x_b = [x_dev[i * B:(i + 1) * B] for i in range(n_dev)]      # stable views: the step's CUDA graph is keyed by pointer
    y_b = [y_dev[i * B:(i + 1) * B] for i in range(n_dev)]
    def do_step(s):
        for j in range(E):
            model.train_step(x_b[(s * E + j) % n_dev], y_b[(s * E + j) % n_dev])
        if fed:
            fed_average()

    for s in range(3 * n_dev):                                  # untimed priming: every batch eagerly, then captured
        model.train_step(x_b[s % n_dev], y_b[s % n_dev])
    for s in range(W):
        do_step(s)
    e0, e1 = device.new_event(), device.new_event()
    barrier()
    launches0 = device.launch_count()
    t_wall0 = time.time()
    lib.nm_event_record(e0, device.stream())
    for s in range(K):
        do_step(s)
    lib.nm_event_record(e1, device.stream())
    device.synchronize()
    t_wall1 = time.time()
    barrier()
    launches = device.launch_count() - launches0
    ms = C.c_float()
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
    elapsed_ms = max_over_ranks(ms.value)
    value = K * E * B * world / (elapsed_ms / 1e3)
    graphs = len(getattr(model, "_graphs", {}))

    clock_info = clocks.stop(t_wall0, t_wall1)   # before the end-to-end leg: a polling nvidia-smi stalls host-side syncs

    # ---------------- per-kernel pass: the same steps with a CUDA-event pair around every operator ----------
    # (a separate timed region: event records between kernels would serialise the programmatic dependent launches
    #  that the headline region relies on)
    K2 = min(K * E, 50)
    ops.timers = ops.KernelTimers()
    lib.nm_event_record(e0, device.stream())
    for s in range(K2):
        model.train_step(x_b[s % n_dev], y_b[s % n_dev])
    lib.nm_event_record(e1, device.stream())
    device.synchronize()
    lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
    step2_ms = ms.value / K2
    kernel_ms = ops.timers.collect()
    ops.timers = None
    del x_dev, x_b

    # ---------------- end-to-end leg: Model.train from pinned host memory ------------------------
    # K steps = (K // n_host) epochs over the n_host pinned batches + one partial pass; every step copies its
    # batch host->device and reads the running loss/accuracy sums back (16 bytes)
    def e2e_leg(host_arr, steps):
        n_warm = min(max(W, 4), n_host)   # >= 4 steps: both feeder slots are seen twice, i.e. their step graphs are captured
        for _ in range(2):
            model.train(host_arr[:n_warm * B], labels[:n_warm * B], epochs=1, batch_size=B, verbose=0, stream_batches=True)
        full, rem = divmod(steps, n_host)
        h2d = d2h = 0
        barrier()
        lib.nm_event_record(e0, device.stream())
        if fed:                            # a round = Model.train over the client's E local batches + the weight average
            for r in range(steps):
                lo = (r * E) % max(n_host - E + 1, 1)
                model.train(host_arr[lo * B:(lo + E) * B], labels[lo * B:(lo + E) * B], epochs=1, batch_size=B, verbose=0,
                            stream_batches=True)
                fed_average()
                h2d += model.h2d_bytes; d2h += model.d2h_bytes
            full = rem = 0
        if full:
            model.train(host_arr, labels, epochs=full, batch_size=B, verbose=0, stream_batches=True)
            h2d += model.h2d_bytes; d2h += model.d2h_bytes
        if rem:
            model.train(host_arr[:rem * B], labels[:rem * B], epochs=1, batch_size=B, verbose=0, stream_batches=True)
            h2d += model.h2d_bytes; d2h += model.d2h_bytes
        lib.nm_event_record(e1, device.stream())
        device.synchronize()
        barrier()
        lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
        t = max_over_ranks(ms.value)
        return {"value": steps * E * B * world / (t / 1e3), "unit": UNIT, "ms_per_step": t / steps,
                "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps}

    host_f32 = None
    if args.input == "f32" or (world == 1 and not args.no_extras):
        host_f32 = device.PinnedBuffer(host_u8.array.shape, np.float32)
        np.divide(host_u8.array, np.float32(255.0), out=host_f32.array, dtype=np.float32)
    if args.input == "u8":
        e2e = e2e_leg(host_u8.array, K)
        e2e["input"] = "raw uint8 pixels in pinned host memory; X / 255 on the device (examples/test_Conv2D_MNIST.py:27)"
    else:
        e2e = e2e_leg(host_f32.array, K)
        e2e["input"] = "float32 pixels in pinned host memory (normalised on the host)"
    e2e["api"] = f"Model.train(X_pinned_host, y, batch_size={B}, stream_batches=True)"
    if args.input == "u8" and host_f32 is not None:
        alt = e2e_leg(host_f32.array, min(K, 200))
        e2e["float32_host_input"] = {k: alt[k] for k in ("value", "ms_per_step", "h2d_bytes_per_step")}
    del host_f32

    # ---------------- parity at this configuration (every rank takes part) --------------------------------------
    parity = None
    if fed:
        x_par = device.DeviceArray.from_numpy(host_u8.array[:B].astype(np.float32) / 255.0)
        y_par = device.DeviceArray.from_numpy(labels[:B])
    if not args.no_parity:
        def comm_factory():
            c = parallel.Communicator(rank, world)
            return c
        try:
            # the check attaches with the same exchange as the timed model
            parity = (fedavg_parity(model, comm, rank, world, dist, B, E, x_par, y_par) if fed else
                      parity_check(wl, mode, B, rank, world, comm_factory, dist, peer_arg))
        except Exception as e:      # a failed check must not lose the measured line; it is reported as such
            parity = {"ok": False, "error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel ------------------------------------------------
    pk = peaks()
    tot = {k: sum(v) for k, v in kernel_ms.items()}
    top = max(tot, key=tot.get)
    avg_ms = tot[top] / len(kernel_ms[top])
    flops, nbytes = kernel_work(top, B)
    # a kernel timed alone in a region far shorter than a second runs at burst clocks: the burst peak is the denominator
    region_s = step2_ms * K2 / 1e3
    tf_peak = pk["tflops_burst"] if region_s < 1.0 else pk["tflops"]
    t_tensor, t_hbm = flops / (tf_peak * 1e12), nbytes / (pk["hbm_gbs"] * 1e9)
    if t_tensor >= t_hbm:
        achieved, peak, unit, bound = flops / (avg_ms * 1e-3) / 1e12, tf_peak, "TFLOP/s", "tensor"
    else:
        achieved, peak, unit, bound = nbytes / (avg_ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s", "hbm"
    per_kernel = {}
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f, nb = kernel_work(k, B)
        ms_k = v / len(kernel_ms[k])
        rec = {"ms": round(ms_k, 5), "launches_per_step": len(kernel_ms[k]) / K2}
        if f:
            rec["tflops"] = round(f / (ms_k * 1e-3) / 1e12, 2)
            rec["frac_tensor"] = round(f / (ms_k * 1e-3) / 1e12 / tf_peak, 3)
        if nb:
            rec["gbs"] = round(nb / (ms_k * 1e-3) / 1e9, 1)
            rec["frac_hbm"] = round(nb / (ms_k * 1e-3) / 1e9 / pk["hbm_gbs"], 3)
        per_kernel[k] = rec
    traffic, traffic_source = ncu_traffic(top)
    roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                "avg_ms": avg_ms, "algorithmic_flops": flops, "algorithmic_bytes": nbytes,
                "share_of_step": tot[top] / K2 / step2_ms,
                "timing": f"CUDA events around every operator in a second timed region of {K2} steps ({step2_ms:.4f} ms/step, "
                          f"{region_s * 1e3:.1f} ms in all; the headline region carries no per-kernel events)",
                "peak_source": pk["source"] + ((" bf16 dense burst (cuBLAS)" if region_s < 1.0 else " bf16 dense sustained (cuBLAS)")
                                               if bound == "tensor" else " HBM copy"),
                "kernels": per_kernel}
    if world > 1 and not fed:
        # the instrumented pass is eager: the exchange's push kernel (side stream, behind conv2's wgrad) then shares the SMs
        # with whichever kernel the compute stream runs next, and that kernel's event pair absorbs the wait
        roofline["note"] = ("per-kernel times at N > 1 are taken while the gradient push runs beside the backward kernels "
                            "(the event pair of the kernel launched next on the compute stream -- tc_conv1_bwd in the MNIST plan -- absorbs the "
                            "contention); the N = 1 line has the kernels timed alone")
    line = {"metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (three bf16 planes per operand, FP32-exact)"}.get(args.mode, "f32"), "data": "synthetic",
            "config": {"workload": wl["desc"],
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "gradient_exchange": ("none: clients train independently; weights averaged once per round (NCCL all-reduce of "
                                             "the pre-scaled slab)" if fed else "none (1 GPU)" if world == 1 else
                                             "gradient slab pushed over NVLink peer memory behind each wgrad (overlapping the rest of "
                                             "backward) + fused collect / rank-ordered sum / Adam kernel (nm_dp_*)"
                                             if comm.peer is not None else "NCCL all-reduce, buckets on a side stream"),
                       "mode": {"bf16": "bf16 tensor-core plan (bf16 operands, fp32 accumulate, fp32 master weights)",
                                "bf16x3": "FP32-exact tensor-core plan (three bf16 planes per operand, six tcgen05 products per contraction, "
                                          "fp32 accumulate, fp32 master weights)"}.get(args.mode, "fp32_exact"),
                       "step_tflops": wl["flops_per_sample"] * E * B * world / (elapsed_ms / K * 1e-3) / 1e12,
                       "cuda_graphs": graphs,
                       "l2": "inputs cycle over 8 distinct batches; the per-step activation working set "
                             "(>1 GB) is far larger than the 126 MB L2, so no explicit flush is needed"},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
    if parity is not None:
        line["parity"] = parity
    if world == 1 and not args.no_extras:
        # FP32-exact mode (the mode that meets rtol 1e-4 / atol 1e-5 against the NumPy path) on the same workload
        m32 = build_model(spec, params, mode=None)
        m32.epoch_end_accuracy = False
        n32 = 2
        x32 = device.DeviceArray.from_numpy(host_u8.array[:n32 * B].astype(np.float32) / 255.0)
        y32 = device.DeviceArray.from_numpy(labels[:n32 * B])
        for s in range(3):
            m32.train_step(x32[(s % n32) * B:(s % n32 + 1) * B], y32[(s % n32) * B:(s % n32 + 1) * B])
        k32 = 10 if args.workload == "mnist" else 4
        lib.nm_event_record(e0, device.stream())
        for s in range(k32):
            m32.train_step(x32[(s % n32) * B:(s % n32 + 1) * B], y32[(s % n32) * B:(s % n32 + 1) * B])
        lib.nm_event_record(e1, device.stream())
        device.synchronize()
        lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
        line["fp32_exact"] = {"value": k32 * B / (ms.value / 1e3), "unit": UNIT, "ms_per_step": ms.value / k32, "steps": k32,
                              "step_tflops": wl["flops_per_sample"] * B / (ms.value / k32 * 1e-3) / 1e12,
                              "mode": "FP32-exact CUDA-core kernels, reference layouts (Model.finalize()): rtol 1e-4 / atol 1e-5 "
                                      "against the NumPy path (tests/test_gpu_model.py)"}
        if args.workload in ("mlp", "mnist", "vgg"):
            # the same tolerance on the TENSOR cores: three bf16 planes per operand, six products per contraction (mode "bf16x3")
            mx3 = build_model(spec, params, mode="bf16x3")
            mx3.epoch_end_accuracy = False
            for s in range(8):                      # each of the two input slices is seen, then captured, then replayed
                mx3.train_step(x32[(s % n32) * B:(s % n32 + 1) * B], y32[(s % n32) * B:(s % n32 + 1) * B])
            kx3 = 200 if args.workload == "mlp" else 10
            lib.nm_event_record(e0, device.stream())
            for s in range(kx3):
                mx3.train_step(x32[(s % n32) * B:(s % n32 + 1) * B], y32[(s % n32) * B:(s % n32 + 1) * B])
            lib.nm_event_record(e1, device.stream())
            device.synchronize()
            lib.nm_event_elapsed_ms(e0, e1, C.byref(ms))
            line["fp32_exact_tensor_core"] = {
                "value": kx3 * B / (ms.value / 1e3), "unit": UNIT, "ms_per_step": ms.value / kx3, "steps": kx3,
                "cuda_graphs": len(getattr(mx3, "_graphs", {})),
                "step_tflops": wl["flops_per_sample"] * B / (ms.value / kx3 * 1e-3) / 1e12,
                "mode": "Model.finalize(mode='bf16x3'): every conv / hidden Dense contraction as 3 bf16 planes x 6 tcgen05 products with FP32 "
                        "accumulation, head on the FP32 CUDA-core kernels; rtol 1e-4 / atol 1e-5 against the NumPy path "
                        "(tests/test_gpu_tc_mlp_x3.py, tests/test_gpu_vgg_tc_x3_plan.py)"}
            del mx3
        del m32, x32
        line["hbm_kernels"] = optimizer_bandwidth()
    if world == 1 and not args.no_cpu_baseline:
        v, cms, cores, kind, desc = cpu_arm(wl, wl["cpu_steps"], 1, wl["cpu_sample"])
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    _emit(line, out_fd)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
