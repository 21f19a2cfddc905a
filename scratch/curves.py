import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import golden, synth_learnable
from oracle import ref_net
from test_gpu_model import build
g = golden("cnn_200steps.npz")
spec = ref_net.mnist_cnn_spec()
x, y = synth_learnable(9, 64 * 200)
out = {}
for mode in (None, "bf16"):
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(8)), mode=mode, learning_rate=0.001)
    L = []
    for s in range(200):
        model._accum.fill_zero()
        model.train_step(x[s*64:(s+1)*64], y[s*64:(s+1)*64])
        L.append(model._read_accum()[0]/64)
    out[str(mode)] = np.array(L)
np.savez('/root/repo/gpurun_out/curves.npz', **out)
# forward test diag
params = ref_net.init_params(spec, np.random.RandomState(3), init="he")
model = build(spec, params, mode="bf16")
net = ref_net.RefNet(spec, params)
xx, yy = synth_learnable(4, 96)
o = np.asarray(model.forward(xx, training=False)); r = net.forward(xx)
print("probs maxabs", np.abs(o-r).max(), "rel", (np.abs(o-r)/(np.abs(r)+1e-12)).max())
for li in (1,3):
    a = np.asarray(model.layers[li].output); rr = net.outs[li]
    print(li, a.shape, rr.shape, np.abs(a-rr).max(), np.abs(rr).max())
print(model.predict(xx, batch_size=32).argmax(1)[:10], o.argmax(1)[:10])
