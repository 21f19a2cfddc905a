"""The reference's examples/test_Conv2D_MNIST.py on the B200 path, with synthetic MNIST-shaped data (no TensorFlow,
no network).  The model-building block is the reference's, call for call (examples/test_Conv2D_MNIST.py:42-63).

    python examples/test_Conv2D_MNIST.py [--mode bf16|bf16x3|fp32] [--n 16384] [--batch 4096] [--epochs 3]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuromah_b200
neuromah_b200.install_alias()          # `import Neuromah.src...` below resolves to the B200 package
from Neuromah.src.layers import Layer_Conv2D, Layer_MaxPooling2D, Layer_Dense, Layer_Flatten
from Neuromah.src.optimizers import Optimizer_Adam
from Neuromah.src.losses import Loss_CategoricalCrossentropy
from Neuromah.src.core import Model
from Neuromah.src.activations import Activation_ReLU, Activation_Softmax
from Neuromah.src.metrics import Accuracy_Categorical
from Neuromah.src.utils import TensorMonitor


def synthetic_digits(n, rng):
    """Learnable MNIST-shaped data: class k lights a bar at row 2k+4 on top of uniform noise in [0, 1)."""
    labels = rng.integers(0, 10, n)
    x = rng.random((n, 1, 28, 28), dtype=np.float32) * 0.5
    for k in range(10):
        x[labels == k, 0, 2 * k + 4:2 * k + 7, 6:22] += 0.5
    return x, np.eye(10)[labels]


ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="bf16", choices=["bf16", "bf16x3", "fp32"],
                help="bf16 = tensor-core plan; bf16x3 = FP32-exact on the tensor cores (3 bf16 planes per operand); fp32 = FP32-exact CUDA-core kernels")
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--log-dir", default="/tmp/tensor_logs")
args = ap.parse_args()
rng = np.random.default_rng(0)
train_images, train_labels_onehot = synthetic_digits(args.n, rng)
val_images, val_labels = synthetic_digits(1024, rng)

model = Model(device='cuda')
model.add(Layer_Conv2D(in_channels=1, out_channels=32, kernel_size=3, padding='same', activation=Activation_ReLU()))
model.add(Layer_MaxPooling2D(pool_size=2, strides=2))
model.add(Layer_Conv2D(in_channels=32, out_channels=64, kernel_size=3, padding='same', activation=Activation_ReLU()))
model.add(Layer_MaxPooling2D(pool_size=2, strides=2))
model.add(Layer_Flatten())
model.add(Layer_Dense(n_inputs=64 * 7 * 7, n_neurons=10))
model.add(Activation_Softmax())
model.set(loss=Loss_CategoricalCrossentropy,
          optimizer=Optimizer_Adam(learning_rate=0.001),
          accuracy=Accuracy_Categorical,
          tensorMonitor=TensorMonitor(log_dir=args.log_dir))
model.finalize(mode=None if args.mode == "fp32" else args.mode)

t0 = time.perf_counter()
model.train(train_images, train_labels_onehot, validation_data=(val_images, val_labels), epochs=args.epochs,
            batch_size=args.batch, verbose=1)
dt = time.perf_counter() - t0
print(f"{args.epochs * args.n / dt:,.0f} samples/s over {args.epochs} epochs incl. per-epoch accuracy sweep and validation")
pred = model.predict(val_images[:8], batch_size=8)
print("predictions:", pred.argmax(1), "labels:", val_labels[:8].argmax(1))
