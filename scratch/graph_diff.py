import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import synth_learnable
from oracle import ref_net
from test_gpu_model import build
from neuromah_b200.device import DeviceArray
spec = ref_net.mnist_cnn_spec()
params = ref_net.init_params(spec, np.random.RandomState(9), init="he")
x, y = synth_learnable(12, 2 * 64)
def run(graph, decay, steps=8):
    model = build(spec, params, mode="bf16", learning_rate=0.002)
    model.optimizer.decay = decay
    model.graph_steps = graph
    xb = [DeviceArray.from_numpy(x[i * 64:(i + 1) * 64]) for i in range(2)]
    yb = [DeviceArray.from_numpy(np.argmax(y[i * 64:(i + 1) * 64], axis=1).astype(np.int32)) for i in range(2)]
    snaps = []
    for s in range(steps):
        model.train_step(xb[s % 2], yb[s % 2])
        snaps.append((model.arena.host_params().copy(), model.arena.host_grads().copy()))
    return snaps
for decay in (0.0, 1e-3):
    e1, e2, g1, g2 = run(False, decay), run(False, decay), run(True, decay), run(True, decay)
    for s in range(8):
        print(f"decay {decay} step {s}: eager-eager p {np.abs(e1[s][0]-e2[s][0]).max():.2e}  graph-graph p {np.abs(g1[s][0]-g2[s][0]).max():.2e}  "
              f"eager-graph p {np.abs(e1[s][0]-g1[s][0]).max():.2e} g {np.abs(e1[s][1]-g1[s][1]).max():.2e}")
