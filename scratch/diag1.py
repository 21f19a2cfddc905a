import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from conftest import synth
from oracle import ref_net
from test_gpu_model import build
spec = ref_net.mnist_cnn_spec()
params = ref_net.init_params(spec, np.random.RandomState(4))
model = build(spec, params, learning_rate=0.001)
net = ref_net.RefNet(spec, params)
x, y = synth(5, 512, (1, 28, 28))
names = ["c1w","c1b","c2w","c2b","dw","db"]
sizes = [288,32,18432,64,31360,10]
def split(v):
    out=[];o=0
    for s in sizes: out.append(v[o:o+s]); o+=s
    return out
for ep in range(2):
  for s in range(2):
    xb, yb = x[s*256:(s+1)*256], y[s*256:(s+1)*256]
    model.train_step(xb, yb); net.step(xb, yb)
    g1 = split(model.arena.host_grads().astype(np.float64)); g2 = split(net.flat("grads"))
    w1 = split(model.arena.host_params().astype(np.float64)); w2 = split(net.flat())
    for n,a,b,c,d in zip(names,g1,g2,w1,w2):
        i = np.argmax(np.abs(c-d))
        print(ep,s,n,"grad maxabs %.3e maxdiff %.3e | w maxdiff %.3e at g_ours %.3e g_ref %.3e"%(np.abs(b).max(), np.abs(a-b).max(), np.abs(c-d).max(), a[i], b[i]))
