import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import ref_net
from test_gpu_model import build
from neuromah_b200._lib import lib
spec = ref_net.mnist_cnn_spec()
params = ref_net.init_params(spec, np.random.RandomState(9), init="he")
rng = np.random.default_rng(33)
lab = rng.integers(0, 10, 4 * 64)
t = (np.random.default_rng(12345).random((10, 1, 28, 28)) > 0.7)
xu = np.clip(rng.integers(0, 160, (len(lab), 1, 28, 28)) + 95 * t[lab], 0, 255).astype(np.uint8)
xf = xu.astype(np.float32) / 255.0
y = np.eye(10)[lab]
runs = {}
for name, x, kw, graph in (("f32", xf, {}, True), ("f32b", xf, {}, True), ("u8", xu, {}, True), ("f32_eager", xf, {}, False), ("u8_eager", xu, {}, False), ("u8_stream", xu, {"stream_batches": True}, True)):
    m = build(spec, params, mode="bf16")
    m.graph_steps = graph
    n0 = [0]
    orig = lib.nm_graph_launch
    m.train(x, y, epochs=2, batch_size=64, verbose=0, **kw)
    runs[name] = m.arena.host_params()
    print(name, "graphs", len(getattr(m, "_graphs", {})), "loss", m.last_epoch["loss"])
for k in runs:
    print(k, "== f32:", np.array_equal(runs[k], runs["f32"]), np.abs(runs[k] - runs["f32"]).max())
print("u8_eager == f32_eager", np.array_equal(runs["u8_eager"], runs["f32_eager"]))
