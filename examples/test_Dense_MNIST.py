"""BASELINE.json config 1: the 784-128-10 MLP (ReLU, softmax cross-entropy, Adam) on synthetic 28x28 data, batch 128,
through the reference's API.  ``--as-written`` builds the stack exactly as the reference's examples/test_Dense_MNIST.py:38-49
does -- two Dropout layers, a last Dense with an embedded Softmax followed by Activation_Softmax again, loss and
accuracy passed as instances, batch 32 -- instead of BASELINE.json's clean 784-128-10 config.

    python examples/test_Dense_MNIST.py [--n 25600] [--epochs 2] [--as-written]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuromah_b200
neuromah_b200.install_alias()          # `import Neuromah.src...` below resolves to the B200 package
# the reference script's own import lines (examples/test_Dense_MNIST.py:2-9), module-path imports included
from Neuromah.src.core import Model
from Neuromah.src.layers import Layer_Dense
from Neuromah.src.losses.categorical_crossentropy import Loss_CategoricalCrossentropy
from Neuromah.src.optimizers import Optimizer_Adam
from Neuromah.src.activations import Activation_Softmax, Activation_ReLU
from Neuromah.src.metrics.Accuracy_Categorical import Accuracy_Categorical
from Neuromah.src.layers import Layer_Dropout

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=25600)
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--as-written", action="store_true",
                help="the reference script's stack verbatim (Dropout twice, softmax applied twice); FP32-exact mode only")
ap.add_argument("--dropout", action="store_true", help="the reference script's stack with a single softmax (Dense+ReLU, Dropout, "
                "Dense+ReLU, Dropout, Dense, Softmax): takes the tensor-core MLP plan in --mode bf16")
ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16", "bf16x3"],
                help="bf16 = tcgen05 MLP plan (neuromah_b200/tc_plan_mlp.py); bf16x3 = the same plan FP32-exact on the tensor cores")
args = ap.parse_args()
rng = np.random.default_rng(1)
labels = rng.integers(0, 10, args.n)
X = rng.random((args.n, 784), dtype=np.float32) * 0.5
for k in range(10):
    X[labels == k, 70 * k:70 * k + 60] += 0.5
y = np.eye(10)[labels]

model = Model()
if args.as_written:
    model.add(Layer_Dense(n_inputs=784, n_neurons=128, activation=Activation_ReLU()))
    model.add(Layer_Dropout(rate=0.25))
    model.add(Layer_Dense(n_inputs=128, n_neurons=64, activation=Activation_ReLU()))
    model.add(Layer_Dropout(rate=0.25))
    model.add(Layer_Dense(n_inputs=64, n_neurons=10, activation=Activation_Softmax()))
    model.add(Activation_Softmax())
    model.set(loss=Loss_CategoricalCrossentropy(model=model), optimizer=Optimizer_Adam(learning_rate=0.001),
              accuracy=Accuracy_Categorical(model=model))
    batch = 32
elif args.dropout:
    model.add(Layer_Dense(n_inputs=784, n_neurons=128, activation=Activation_ReLU()))
    model.add(Layer_Dropout(rate=0.25))
    model.add(Layer_Dense(n_inputs=128, n_neurons=64, activation=Activation_ReLU()))
    model.add(Layer_Dropout(rate=0.25))
    model.add(Layer_Dense(n_inputs=64, n_neurons=10))
    model.add(Activation_Softmax())
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=Optimizer_Adam(learning_rate=0.001), accuracy=Accuracy_Categorical)
    batch = 128
else:
    model.add(Layer_Dense(784, 128, activation=Activation_ReLU()))
    model.add(Layer_Dense(128, 10))
    model.add(Activation_Softmax())
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=Optimizer_Adam(learning_rate=0.001), accuracy=Accuracy_Categorical)
    batch = 128
if args.as_written and args.mode != "fp32":
    raise SystemExit("--as-written (softmax applied twice) runs in FP32-exact mode only; use --dropout --mode bf16")
model.finalize(mode=None if args.mode == "fp32" else args.mode)
t0 = time.perf_counter_ns()
model.train(X, y, epochs=args.epochs, batch_size=batch, validation_data=(X[:1024], y[:1024]))
print(f"Total training time: {time.perf_counter_ns() - t0} ns")
