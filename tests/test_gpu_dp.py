"""Data-parallel gradient exchange over NVLink peer memory (nm_dp_*, neuromah_b200/csrc/nm_dp.cu).

world = 1 runs on any GPU box and pins the kernel's arithmetic against the plain Adam kernel bit for bit; the
multi-rank checks (tests/dp_worker.py) need >= 2 GPUs and are skipped otherwise (`gpurun --gpus 2`)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [1, 5, 1024, 4099, 50890])
def test_world1_fused_step_equals_adam_kernel_bitwise(n):
    from neuromah_b200 import device
    from neuromah_b200._lib import lib
    device.init(0)
    rng = np.random.default_rng(n)
    p0 = rng.standard_normal(n).astype(np.float32)
    st = device.DeviceArray.zeros((128,), np.uint8)
    lib.nm_adam_state_init(st.p, 1e-3, 1e-4, 0.9, 0.999, 1e-7, C.c_longlong(0), device.stream())
    st2 = device.DeviceArray.zeros((128,), np.uint8)
    lib.nm_adam_state_init(st2.p, 1e-3, 1e-4, 0.9, 0.999, 1e-7, C.c_longlong(0), device.stream())
    a = [device.DeviceArray.from_numpy(p0), None, device.DeviceArray.zeros((n,)), device.DeviceArray.zeros((n,))]
    b = [device.DeviceArray.from_numpy(p0), None, device.DeviceArray.zeros((n,)), device.DeviceArray.zeros((n,))]
    dp = C.c_void_p()
    lib.nm_dp_create(C.byref(dp), 0, 1, n)
    try:
        for step in range(5):                                   # odd and even parities, the step counter advancing on the device
            g = rng.standard_normal(n).astype(np.float32)
            a[1] = device.DeviceArray.from_numpy(g)
            b[1] = device.DeviceArray.from_numpy(g)
            lib.nm_adam_step_dev_f32(a[0].p, a[1].p, a[2].p, a[3].p, n, st.p, 2, device.stream())
            lib.nm_adam_advance_f32(st.p, device.stream())
            lib.nm_dp_allreduce_adam_dev_f32(dp, b[0].p, b[1].p, b[2].p, b[3].p, n, st2.p, 2, 1, device.stream())   # advance fused
            device.synchronize()
            for u, v in zip(a, b):
                assert np.array_equal(u.numpy(), v.numpy())
        bad, done = C.c_int(-1), C.c_uint(0)
        lib.nm_dp_status(dp, C.byref(bad), C.byref(done))
        assert bad.value == 0 and done.value == 5
        # all-reduce only: world 1 leaves the slab unchanged
        g = device.DeviceArray.from_numpy(p0)
        lib.nm_dp_allreduce_sum_f32(dp, g.p, n, device.stream())
        device.synchronize()
        assert np.array_equal(g.numpy(), p0)
    finally:
        lib.nm_dp_destroy(dp)


def test_slab_larger_than_mailbox_is_rejected():
    from neuromah_b200 import device
    from neuromah_b200._lib import lib
    device.init(0)
    dp = C.c_void_p()
    lib.nm_dp_create(C.byref(dp), 0, 1, 100)
    try:
        g = device.DeviceArray.zeros((4096,))
        with pytest.raises(ValueError):
            lib.nm_dp_allreduce_sum_f32(dp, g.p, 4096, device.stream())
    finally:
        lib.nm_dp_destroy(dp)


@pytest.mark.timeout(900)
def test_multi_gpu_peer_exchange_parity():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=840)
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
