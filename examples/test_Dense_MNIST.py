"""BASELINE.json config 1: the 784-128-10 MLP (ReLU, softmax cross-entropy, Adam) on synthetic 28x28 data, batch 128,
through the reference's API (examples/test_Dense_MNIST.py builds the same kind of stack; its Dropout layers and doubled
softmax -- SURVEY Q10 -- are left out, as BASELINE.json's config does).

    python examples/test_Dense_MNIST.py [--n 25600] [--epochs 2]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuromah_b200
neuromah_b200.install_alias()          # `import Neuromah.src...` below resolves to the B200 package
from Neuromah.src.layers import Layer_Dense
from Neuromah.src.optimizers import Optimizer_Adam
from Neuromah.src.losses import Loss_CategoricalCrossentropy
from Neuromah.src.core import Model
from Neuromah.src.activations import Activation_ReLU, Activation_Softmax
from Neuromah.src.metrics import Accuracy_Categorical

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=25600)
ap.add_argument("--epochs", type=int, default=2)
args = ap.parse_args()
rng = np.random.default_rng(1)
labels = rng.integers(0, 10, args.n)
X = rng.random((args.n, 784), dtype=np.float32) * 0.5
for k in range(10):
    X[labels == k, 70 * k:70 * k + 60] += 0.5
y = np.eye(10)[labels]

model = Model()
model.add(Layer_Dense(784, 128, activation=Activation_ReLU()))
model.add(Layer_Dense(128, 10))
model.add(Activation_Softmax())
model.set(loss=Loss_CategoricalCrossentropy, optimizer=Optimizer_Adam(learning_rate=0.001), accuracy=Accuracy_Categorical)
model.finalize()
t0 = time.perf_counter_ns()
model.train(X, y, epochs=args.epochs, batch_size=128, validation_data=(X[:1024], y[:1024]))
print(f"Total training time: {time.perf_counter_ns() - t0} ns")
