"""Host-side logic added in round 2, runnable without a GPU: the deferred-free queue that keeps finalizers out of CUDA-graph
captures, bench.py's workload tables / per-kernel work model, and the ncu summariser that feeds `roofline.traffic`."""
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frees_are_deferred_while_a_capture_is_open(monkeypatch):
    """Round-1 driver failure: a finalizer's cudaFree inside a stream capture invalidated the capture.  While
    device._capture['depth'] > 0 a released block is queued; drain_deferred_frees() releases it afterwards."""
    from neuromah_b200 import device
    freed = []

    class FakeLib:
        def nm_free(self, p):
            freed.append(("device", p.value))

        def nm_host_free(self, p):
            freed.append(("host", p.value))
    monkeypatch.setattr(device, "lib", FakeLib())
    blk = object.__new__(device._Block)
    blk.ptr, blk.nbytes = 0x1000, 64
    monkeypatch.setitem(device._capture, "depth", 1)
    assert device.capturing()
    del blk                                             # finalizer runs "inside the capture"
    assert freed == [] and device._capture["device"] == [0x1000]
    device.drain_deferred_frees()                       # still capturing: nothing may be released
    assert freed == []
    monkeypatch.setitem(device._capture, "depth", 0)
    device.drain_deferred_frees()
    assert freed == [("device", 0x1000)] and device._capture["device"] == []
    blk2 = object.__new__(device._Block)                # outside a capture: released immediately
    blk2.ptr, blk2.nbytes = 0x2000, 64
    del blk2
    assert freed[-1] == ("device", 0x2000)


def test_allocation_inside_a_capture_is_refused(monkeypatch):
    from neuromah_b200 import device
    monkeypatch.setattr(device, "init", lambda *a, **k: 0)
    monkeypatch.setitem(device._capture, "depth", 1)
    with pytest.raises(RuntimeError, match="capture"):
        device._Block(1024)


def test_bench_workloads_and_work_model():
    sys.path.insert(0, ROOT)
    import bench
    counts = {}
    for name, wl in bench.WORKLOADS.items():
        spec, params = wl["make"]()
        counts[name] = sum(p["weights"].size + p["biases"].size for p in params if p is not None)
        assert spec[-1]["kind"] == "softmax"
    assert counts["mnist"] == 50186 and counts["fedavg"] == 50186
    assert counts["vgg"] == 1186378 and counts["mlp"] == 101770
    # FLOPs per sample of the tables = 3 x forward MACs x 2, minus the first layer's dgrad
    mnist_fwd = 2 * (32 * 9 * 784 + 64 * 32 * 9 * 196 + 3136 * 10)
    assert abs(3 * mnist_fwd - 2 * 32 * 9 * 784 - bench.WORKLOADS["mnist"]["flops_per_sample"]) < 2e3
    # tagged kernels: conv [ic->oc@h] and dense [n_in->n_out]
    f, b = bench.kernel_work("tc_conv_gen_fprop[64->128@16]", 1024)
    assert f == 2.0 * 1024 * 128 * 64 * 9 * 256 and b == 1024 * 18 * 18 * 2.0 * (64 + 128)
    f, b = bench.kernel_work("tc_conv_gen_wgrad[3->64@32]", 8)
    assert f == 2.0 * 8 * 64 * 3 * 9 * 1024 and b == 8 * 34 * 34 * 2.0 * (64 + 64)
    f, b = bench.kernel_work("tc_dense_gen_fprop[784->128]", 128)
    assert f == 2.0 * 128 * 784 * 128 and b > 0
    assert bench.kernel_work("tc_conv2_bwd", 4096)[0] == 2 * 2.0 * 4096 * 64 * 32 * 9 * 196
    assert bench.kernel_work("no_such_kernel", 1) == (0.0, 0.0)


def test_reference_arm_line_shape():
    """bench.py --impl reference prints ONE JSON line with the keys the driver reads (a tiny run: 2 steps of 16 samples)."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "mlp", "--steps", "3",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["per_step_sample"] == 128 and d["gpu_launches"] == 0


def test_ncu_summariser_writes_traffic(tmp_path):
    csv_text = ('"ID","Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum",'
                '"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",'
                '"sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active",'
                '"launch__registers_per_thread","launch__grid_size","launch__block_size"\n'
                '"","","us","Mbyte","Mbyte","%","%","%","%","register/thread","",""\n'
                '"0","void <unnamed>::k_a(int)","10.5","100.0","50.0","40","70","60","20","64","148","256"\n'
                '"1","void <unnamed>::k_b(int)","20.0","8.0","2.0","4","0","6","20","32","10","128"\n'
                '"2","void <unnamed>::k_a(int)","11.0","120.0","60.0","41","71","61","21","64","148","256"\n')
    script = os.path.join(ROOT, "profiles", "summarize.py")
    work = tmp_path / "profiles"
    work.mkdir()
    local = work / "summarize.py"
    local.write_text(open(script).read())
    r = subprocess.run([sys.executable, str(local), "--traffic", "op_a=k_a:1", "--source", "profiles/x.txt"], input=csv_text,
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "k_a(int)" in r.stdout and "k_b(int)" in r.stdout
    db = json.load(open(work / "ncu_traffic.json"))
    assert db["op_a"]["dram_bytes"] == pytest.approx(180.0e6) and db["op_a"]["source"] == "profiles/x.txt"


def test_committed_traffic_file_is_readable_by_bench():
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic("tc_conv2_bwd")
    assert t and 1e8 < t < 1e9 and src.startswith("profiles/")
    assert os.path.exists(os.path.join(ROOT, src.split(" ")[0]))
    assert bench.ncu_traffic("nope") == (None, None)
