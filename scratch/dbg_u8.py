import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import ref_net
from test_gpu_model import build
spec = ref_net.mnist_cnn_spec()
params = ref_net.init_params(spec, np.random.RandomState(9), init="he")
rng = np.random.default_rng(33)
xu = rng.integers(0, 256, (256, 1, 28, 28)).astype(np.uint8)
xf = xu.astype(np.float32) / 255.0
y = np.eye(10)[rng.integers(0, 10, 256)]
a = build(spec, params, mode="bf16"); b = build(spec, params, mode="bf16")
pa = np.asarray(a.forward(xf, training=True)); pb = np.asarray(b.forward(xu, training=True))
print("probs equal", np.array_equal(pa, pb), np.abs(pa - pb).max())
A1, B1 = np.asarray(a.layers[1].output), np.asarray(b.layers[1].output)
print("act1 equal", np.array_equal(A1, B1), np.abs(A1 - B1).max(), np.mean(A1 != B1))
bad = np.argwhere(A1 != B1)
print(bad[:10])
a.backward(pa, y); b.backward(pb, y)
ga, gb = a.arena.host_grads(), b.arena.host_grads()
print("grads equal", np.array_equal(ga, gb), np.abs(ga - gb).max())
off = 0
for (_, name, _, size, _) in a.arena.table:
    print(name, np.array_equal(ga[off:off+size], gb[off:off+size])); off += size
