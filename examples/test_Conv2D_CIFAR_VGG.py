"""BASELINE.json config 4 on the B200 path: a VGG-style Conv2D stack (C(3->64) C(64->64) P C(64->128)
C(128->128) P C(128->256) C(256->256) P Flatten Dense(4096->10) Softmax, all 3x3 'same' + ReLU, He init) on synthetic
CIFAR-10-shaped 32x32x3 data, built with the reference's own layer / model API (the building block reads like
examples/test_Conv2D_MNIST.py:42-63 of the reference, with more layers).

    python examples/test_Conv2D_CIFAR_VGG.py [--mode bf16|bf16x3|fp32] [--n 8192] [--batch 1024] [--epochs 2]

mode bf16 = the general tensor-core plan (neuromah_b200/tc_plan_gen.py: streamed-filter tcgen05 conv fprop / dgrad /
wgrad); mode fp32 = the FP32-exact per-layer kernels.
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


def synthetic_cifar(n, rng):
    """Learnable CIFAR-shaped data: a fixed random binary 3x32x32 template per class under uniform noise in [0, 1)."""
    labels = rng.integers(0, 10, n)
    templ = (np.random.default_rng(4321).random((10, 3, 32, 32), dtype=np.float32) > 0.7).astype(np.float32)
    x = 0.6 * rng.random((n, 3, 32, 32), dtype=np.float32) + 0.4 * templ[labels]
    return x.astype(np.float32), np.eye(10)[labels]


ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="bf16", choices=["bf16", "bf16x3", "fp32"],
                help="bf16 = tensor-core plan; bf16x3 = FP32-exact on the tensor cores (3 bf16 planes per operand); fp32 = FP32-exact CUDA-core kernels")
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--epochs", type=int, default=2)
args = ap.parse_args()
rng = np.random.default_rng(0)
train_images, train_labels_onehot = synthetic_cifar(args.n, rng)
val_images, val_labels = synthetic_cifar(512, rng)

model = Model(device='cuda')
in_channels = 3
for out_channels in (64, 128, 256):
    model.add(Layer_Conv2D(in_channels=in_channels, out_channels=out_channels, kernel_size=3, padding='same',
                           weight_initializer='he', activation=Activation_ReLU()))
    model.add(Layer_Conv2D(in_channels=out_channels, out_channels=out_channels, kernel_size=3, padding='same',
                           weight_initializer='he', activation=Activation_ReLU()))
    model.add(Layer_MaxPooling2D(pool_size=2, strides=2))
    in_channels = out_channels
model.add(Layer_Flatten())
model.add(Layer_Dense(n_inputs=256 * 4 * 4, n_neurons=10, weight_initializer='he'))
model.add(Activation_Softmax())
model.set(loss=Loss_CategoricalCrossentropy,
          optimizer=Optimizer_Adam(learning_rate=0.001),
          accuracy=Accuracy_Categorical)
model.finalize(mode=None if args.mode == "fp32" else args.mode)

t0 = time.perf_counter()
model.train(train_images, train_labels_onehot, validation_data=(val_images, val_labels), epochs=args.epochs,
            batch_size=args.batch, verbose=1)
dt = time.perf_counter() - t0
print(f"{args.epochs * args.n / dt:,.0f} samples/s over {args.epochs} epochs incl. per-epoch accuracy sweep and validation")
pred = model.predict(val_images[:8], batch_size=8)
print("predictions:", pred.argmax(1), "labels:", val_labels[:8].argmax(1))
