"""Generate golden vectors by running the UNMODIFIED reference (authoring container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_shim.py (Dense, ReLU, Softmax, CCE, fused
gradient, Adam, SGD, Model and the compiled max-pool are the reference's own code; conv
goes through the oracle restatement because Conv2D.cpp needs Eigen -- see DESIGN.md) and
writes small .npz fixtures next to this file.  Inputs are regenerated from seeds by the
tests, so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_net, ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def data(seed, n, shape, classes=10):
    rng = np.random.default_rng(seed)
    x = rng.random((n,) + shape, dtype=np.float32)
    y = np.eye(classes)[rng.integers(0, classes, n)]
    return x, y


def trainable(model):
    seen, out = set(), []
    for l in model.trainable_layers:
        if id(l) not in seen:
            seen.add(id(l))
            out.append(l)
    return out


def flat_params(model):
    return np.concatenate([np.concatenate([l.weights.ravel(), l.biases.ravel()]) for l in trainable(model)])


def mlp_200():
    """BASELINE.json configs[0]: 784-128-10, batch 128, Adam lr 1e-3, 200 steps."""
    spec = ref_net.mlp_spec()
    params = ref_net.init_params(spec, np.random.RandomState(0))
    model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.001)
    x, y = data(1, 128 * 200, (784,))
    losses, accs = [], []
    for s in range(200):
        r = ref_shim.train_step(model, x[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128])
        losses.append(r["loss"]); accs.append(r["acc"])
    np.savez_compressed(os.path.join(HERE, "mlp_200steps.npz"), losses=np.array(losses), accs=np.array(accs),
                        final_params=flat_params(model).astype(np.float32))


def mlp_sgd():
    spec = ref_net.mlp_spec(64, 32, 10)
    params = ref_net.init_params(spec, np.random.RandomState(5))
    model = ref_shim.build_model(spec, params, optimizer="sgd", lr=0.5, decay=0.01)
    x, y = data(6, 32 * 20, (64,))
    losses = []
    for s in range(20):
        losses.append(ref_shim.train_step(model, x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])["loss"])
    np.savez_compressed(os.path.join(HERE, "mlp_sgd_20steps.npz"), losses=np.array(losses),
                        final_params=flat_params(model).astype(np.float32))


def single_layer_optimizers():
    """RMSprop / Adagrad / Nadam keep their state keyed by the bare parameter name (RMSprop.py:99-101), so the reference
    only runs them on a model with ONE trainable layer: softmax regression 64 -> 10, batch 32, 20 steps, with the
    duplicate application per step of SURVEY Q1."""
    spec = [{"kind": "dense", "n_in": 64, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    out = {}
    for kind, lr, decay in (("rmsprop", 0.01, 1e-3), ("adagrad", 0.05, 0.0), ("nadam", 0.01, 1e-2)):
        params = ref_net.init_params(spec, np.random.RandomState(31))
        model = ref_shim.build_model(spec, params, optimizer=kind, lr=lr, decay=decay)
        x, y = data(32, 32 * 20, (64,))
        losses = []
        for s in range(20):
            losses.append(ref_shim.train_step(model, x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])["loss"])
        out[f"{kind}_losses"] = np.array(losses)
        out[f"{kind}_params"] = flat_params(model).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "optimizers_single_layer_20steps.npz"), **out)


def dropout_mlp():
    """Dense(64->32, ReLU) -> Dropout(0.3) -> Dense(32->10) -> Softmax, Adam, batch 16, 10 steps.  The masks the
    reference drew (np.random.binomial, Dropout.py:27) are stored so the device layer can replay them (fixed_mask)."""
    spec = [{"kind": "dense", "n_in": 64, "n_out": 32, "relu": True}, {"kind": "dropout", "rate": 0.3},
            {"kind": "dense", "n_in": 32, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    params = ref_net.init_params(spec, np.random.RandomState(41))
    model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.01)
    x, y = data(42, 16 * 10, (64,))
    np.random.seed(123)
    losses, masks = [], []
    for s in range(10):
        losses.append(ref_shim.train_step(model, x[s * 16:(s + 1) * 16], y[s * 16:(s + 1) * 16])["loss"])
        masks.append(np.array(model.layers[1].binary_mask, dtype=np.float32))
    infer = model.forward(x[:16], training=False)
    np.savez_compressed(os.path.join(HERE, "dropout_mlp_10steps.npz"), losses=np.array(losses), masks=np.stack(masks),
                        final_params=flat_params(model).astype(np.float64), infer=np.asarray(infer, dtype=np.float64))


def dense_example_stack():
    """The stack of examples/test_Dense_MNIST.py:38-43 as written (reduced widths): Dense+ReLU, Dropout, Dense+ReLU,
    Dropout, Dense with an embedded Softmax, then Activation_Softmax again (the doubled softmax), loss and
    accuracy passed as INSTANCES (:47-49).  Adam, batch 32, 6 steps; the dropout masks are stored for replay."""
    ref_shim.load()
    from NeuromahRef.src.core import Model
    from NeuromahRef.src.layers import Layer_Dense, Layer_Dropout
    from NeuromahRef.src.activations import Activation_ReLU, Activation_Softmax
    from NeuromahRef.src.losses import Loss_CategoricalCrossentropy
    from NeuromahRef.src.optimizers import Optimizer_Adam
    from NeuromahRef.src.metrics import Accuracy_Categorical
    rs = np.random.RandomState(61)
    W = [rs.randn(64, 32) * 0.1, rs.randn(32, 16) * 0.1, rs.randn(16, 10) * 0.1]
    model = Model()
    d1 = Layer_Dense(n_inputs=64, n_neurons=32, activation=Activation_ReLU())
    d2 = Layer_Dense(n_inputs=32, n_neurons=16, activation=Activation_ReLU())
    d3 = Layer_Dense(n_inputs=16, n_neurons=10, activation=Activation_Softmax())
    for d, w in zip((d1, d2, d3), W):
        d.weights = w.copy(); d.biases = np.zeros((1, w.shape[1]))
    drop1, drop2 = Layer_Dropout(rate=0.25), Layer_Dropout(rate=0.25)
    for l in (d1, drop1, d2, drop2, d3, Activation_Softmax()):
        model.add(l)
    model.set(loss=Loss_CategoricalCrossentropy(model=model), optimizer=Optimizer_Adam(learning_rate=0.01),
              accuracy=Accuracy_Categorical(model=model))
    model.finalize()
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = data(62, 32 * 6, (64,))
    np.random.seed(321)
    losses, m1, m2 = [], [], []
    for s in range(6):
        losses.append(ref_shim.train_step(model, x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])["loss"])
        m1.append(np.array(drop1.binary_mask, dtype=np.float32)); m2.append(np.array(drop2.binary_mask, dtype=np.float32))
    out = model.forward(x[:32], training=False)
    np.savez_compressed(os.path.join(HERE, "dense_example_stack_6steps.npz"), losses=np.array(losses), masks1=np.stack(m1),
                        masks2=np.stack(m2), w0=W[0], w1=W[1], w2=W[2], final_params=flat_params(model).astype(np.float64),
                        infer=np.asarray(out, dtype=np.float64))


def cnn_small():
    """MNIST CNN of examples/test_Conv2D_MNIST.py:47-53, batch 4, 3 Adam steps, wired conv backward."""
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(2))
    model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.001, conv_backward="wired")
    x, y = data(3, 4 * 3, (1, 28, 28))
    out = {}
    for s in range(3):
        r = ref_shim.train_step(model, x[s * 4:(s + 1) * 4], y[s * 4:(s + 1) * 4])
        out[f"loss{s}"] = r["loss"]
        if s == 0:
            for i, l in enumerate(model.layers):
                out[f"out{i}"] = np.asarray(l.output, dtype=np.float32)
                if i < len(model.layers) - 1 and l.dinputs is not None:
                    out[f"din{i}"] = np.asarray(l.dinputs, dtype=np.float32)
            g = []
            for l in trainable(model):
                p = l.get_parameters()
                g += [np.ravel(p["weights"][1]), np.ravel(p["biases"][1])]
            out["grads0"] = np.concatenate(g).astype(np.float64)
        out[f"params{s}"] = flat_params(model).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cnn_small_3steps.npz"), **out)


def vgg_small_spec():
    """BASELINE.json config 4's stack narrowed 8x so the fixture stays small:
    C(3->8) C(8->8) P C(8->16) C(16->16) P Flatten Dense(1024->10) Softmax on 32x32x3 inputs."""
    spec, ic = [], 3
    for oc in (8, 16):
        spec += [{"kind": "conv", "ic": ic, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "conv", "ic": oc, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"}]
        ic = oc
    return spec + [{"kind": "flatten"}, {"kind": "dense", "n_in": 16 * 8 * 8, "n_out": 10, "relu": False}, {"kind": "softmax"}]


def vgg_small():
    """The VGG-style stack through the unmodified reference Python (He-initialised parameters handed in), batch 6,
    2 Adam steps, wired conv backward: conv -> conv chains (ReLU backward between two convs, dgrad of a non-first conv
    feeding a conv) that the MNIST CNN never exercises."""
    spec = vgg_small_spec()
    params = ref_net.init_params(spec, np.random.RandomState(31), init="he")
    model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.001, conv_backward="wired")
    x, y = data(32, 6 * 2, (3, 32, 32))
    out = {}
    for s in range(2):
        r = ref_shim.train_step(model, x[s * 6:(s + 1) * 6], y[s * 6:(s + 1) * 6])
        out[f"loss{s}"] = r["loss"]
        if s == 0:
            out["probs0"] = np.asarray(model.layers[-1].output, dtype=np.float64)
            g = []
            for l in trainable(model):
                p = l.get_parameters()
                g += [np.ravel(p["weights"][1]), np.ravel(p["biases"][1])]
            out["grads0"] = np.concatenate(g).astype(np.float64)
        out[f"params{s}"] = flat_params(model).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "vgg_small_2steps.npz"), **out)


def cnn_stub():
    """Same net with the reference's SHIPPED stub conv backward, 2 steps."""
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(2))
    model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.001, conv_backward="stub")
    x, y = data(3, 4 * 2, (1, 28, 28))
    out = {}
    for s in range(2):
        out[f"loss{s}"] = ref_shim.train_step(model, x[s * 4:(s + 1) * 4], y[s * 4:(s + 1) * 4])["loss"]
    out["final_params"] = flat_params(model).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cnn_stub_2steps.npz"), **out)


def cnn_200():
    """MNIST CNN, batch 64, Adam 1e-3, 200 steps on LEARNABLE synthetic data: the loss curve the
    north star asks to match (FP32-exact and bf16 modes)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import synth_learnable
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(8))
    model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.001, conv_backward="wired")
    x, y = synth_learnable(9, 64 * 200)
    losses, accs = [], []
    for s in range(200):
        r = ref_shim.train_step(model, x[s * 64:(s + 1) * 64], y[s * 64:(s + 1) * 64])
        losses.append(r["loss"]); accs.append(r["acc"])
    np.savez_compressed(os.path.join(HERE, "cnn_200steps.npz"), losses=np.array(losses), accs=np.array(accs))


def ops():
    """Op-level vectors straight from the reference classes."""
    ref_shim.load()
    from NeuromahRef.src.layers import Layer_Dense, Layer_MaxPooling2D
    from NeuromahRef.src.activations import Activation_ReLU, Activation_Softmax
    from NeuromahRef.src.losses import Loss_CategoricalCrossentropy
    from NeuromahRef.src.losses.Activation_Softmax_Loss_CategoricalCrossentropy import \
        Activation_Softmax_Loss_CategoricalCrossentropy
    from NeuromahRef.src.optimizers import Optimizer_Adam
    rng = np.random.default_rng(11)
    out = {}
    # Dense fwd/bwd with ReLU + L1/L2
    np.random.seed(7)
    d = Layer_Dense(20, 12, activation=Activation_ReLU(), weight_regularizer_l1=0.01, weight_regularizer_l2=0.02)
    x = rng.standard_normal((9, 20)).astype(np.float32); dv = rng.standard_normal((9, 12))
    d.forward(x, True); d.backward(dv)
    out.update(dense_w=d.weights, dense_x=x, dense_dv=dv, dense_out=d.output, dense_dw=d.dweights,
               dense_db=d.dbiases, dense_dx=d.dinputs)
    # Softmax + CCE + fused gradient
    z = rng.standard_normal((7, 10)) * 3
    sm = Activation_Softmax(); sm.forward(z)
    yy = rng.integers(0, 10, 7)
    loss = Loss_CategoricalCrossentropy(model=None)
    out.update(sm_in=z, sm_out=sm.output, cce=loss.forward(sm.output, yy), cce_onehot=loss.forward(sm.output, np.eye(10)[yy]),
               labels=yy)
    f = Activation_Softmax_Loss_CategoricalCrossentropy(); f.backward(sm.output, np.eye(10)[yy])
    out["fused_grad"] = f.dinputs
    loss.backward(sm.output, yy); out["cce_bwd"] = loss.dinputs
    sm.backward(dvs := rng.standard_normal((7, 10))); out.update(sm_dv=dvs, sm_bwd=sm.dinputs)
    # max pool: 'same' padding, overlapping windows, ties
    xp = np.round(rng.standard_normal((2, 3, 7, 9)) * 2).astype(np.float32)       # many ties
    for tag, (ps, st, pad) in {"a": (2, 2, "valid"), "b": (3, 2, "same"), "c": ((2, 3), (1, 2), "same")}.items():
        p = Layer_MaxPooling2D(ps, strides=st, padding=pad); p.forward(xp)
        dvp = rng.standard_normal(p.output.shape).astype(np.float32); p.backward(dvp)
        out.update({f"pool_{tag}_out": p.output, f"pool_{tag}_idx": p.max_indices, f"pool_{tag}_dv": dvp,
                    f"pool_{tag}_dx": p.dinputs})
    out["pool_x"] = xp
    # Adam: 3 iterations on one mock layer, decay on
    class L:
        def __init__(s): s.w = rng.standard_normal((5, 4)); s.g = None
        def get_parameters(s): return {"weights": (s.w, s.g)}
    lay = L(); out["adam_w0"] = lay.w.copy()
    opt = Optimizer_Adam(learning_rate=0.01, decay=0.1)
    gs = []
    for it in range(3):
        lay.g = rng.standard_normal((5, 4)) * (10.0 ** -it); gs.append(lay.g.copy())
        opt.pre_update_params(); opt.update_params(lay); opt.update_params(lay); opt.post_update_params()
        out[f"adam_w{it + 1}"] = lay.w.copy()
    out["adam_g"] = np.stack(gs)
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)


def losses_family():
    """The other output activations / losses of the Model.train step, from the UNMODIFIED reference classes:
    Activation_Sigmoid / Activation_Linear, Loss_MeanSquaredError / MeanAbsoluteError / BinaryCrossentropy,
    Accuracy_Regression / Accuracy_Categorical(binary=True).  Per-op vectors on the operands of
    tests/conftest.py::losses_family_inputs and three small models (conftest.LOSSES_FAMILY_CASES) stepped exactly as
    Model.train does (batch 32)."""
    ref_shim.load()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import LOSSES_FAMILY_CASES, losses_family_case, losses_family_inputs
    from NeuromahRef.src.core import Model
    from NeuromahRef.src.layers import Layer_Dense
    from NeuromahRef.src.activations import Activation_ReLU, Activation_Sigmoid, Activation_Linear
    from NeuromahRef.src.losses import Loss_MeanSquaredError, Loss_MeanAbsoluteError, Loss_BinaryCrossentropy
    from NeuromahRef.src.optimizers import Optimizer_Adam, Optimizer_SGD
    from NeuromahRef.src.metrics import Accuracy_Categorical, Accuracy_Regression
    out = {}
    v = {k: a.astype(np.float64) for k, a in losses_family_inputs().items()}
    sg = Activation_Sigmoid()
    sg.forward(v["x"], training=True)
    sg.backward(v["dv"])
    out["sigmoid_out"], out["sigmoid_dx"], out["sigmoid_pred"] = sg.output, sg.dinputs, sg.predictions(sg.output)
    for name, cls, t in (("mse", Loss_MeanSquaredError, "t_reg"), ("mae", Loss_MeanAbsoluteError, "t_reg"),
                         ("bce", Loss_BinaryCrossentropy, "t_bin")):
        loss = cls(model=None)
        out[f"{name}_fwd"] = loss.forward(v["pred"], v[t])
        loss.backward(v["pred"], v[t])
        out[f"{name}_bwd"] = loss.dinputs
        out[f"{name}_fwd4"] = loss.forward(v["pred4"], v["t4"])
        loss.backward(v["pred4"], v["t4"])
        out[f"{name}_bwd4"] = loss.dinputs
    losses = {"mse": Loss_MeanSquaredError, "mae": Loss_MeanAbsoluteError, "bce": Loss_BinaryCrossentropy}
    for name, c in LOSSES_FAMILY_CASES.items():
        w, x, y = losses_family_case(name)
        model = Model()
        d1 = Layer_Dense(n_inputs=8, n_neurons=16, activation=Activation_ReLU())
        d2 = Layer_Dense(n_inputs=16, n_neurons=c["n_out"])
        for d, wi in zip((d1, d2), w):
            d.weights = wi.copy(); d.biases = np.zeros((1, wi.shape[1]))
        for l in (d1, d2, Activation_Sigmoid() if c["act"] == "sigmoid" else Activation_Linear()):
            model.add(l)
        opt = Optimizer_Adam(learning_rate=c["lr"]) if c["opt"] == "adam" else Optimizer_SGD(learning_rate=c["lr"])
        acc = Accuracy_Categorical(binary=True, model=model) if c["acc"] == "binary" else Accuracy_Regression(model=model)
        model.set(loss=losses[c["loss"]](model=model), optimizer=opt, accuracy=acc)
        model.finalize()
        model.accuracy.init(y)                      # Model.train does this first (:119)
        model.loss.new_pass(); model.accuracy.new_pass()
        ls, accs = [], []
        for s in range(c["steps"]):
            r = ref_shim.train_step(model, x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])
            ls.append(r["loss"]); accs.append(r["acc"])
        out[f"{name}_losses"], out[f"{name}_accs"] = np.array(ls), np.array(accs)
        out[f"{name}_final_params"] = flat_params(model).astype(np.float64)
        out[f"{name}_acc_running"] = np.array([model.accuracy.accumulated_sum, model.accuracy.accumulated_count], dtype=np.float64)
        out[f"{name}_infer"] = np.asarray(model.forward(x[:32], training=False), dtype=np.float64)
        if c["acc"] == "regression":
            out[f"{name}_precision"] = np.float64(model.accuracy.precision)
    np.savez_compressed(os.path.join(HERE, "losses_family.npz"), **out)


def ctor_surface():
    """Constructor differential: every call of tests/conftest.py::CTOR_CASES on the UNMODIFIED reference layer classes --
    the exception type it raises, or the attributes (names, array shapes, plain values) of the object it builds."""
    import json
    ref_shim.load()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import CTOR_CASES, ctor_outcome
    import NeuromahRef.src.layers as L
    from NeuromahRef.src.activations import Activation_ReLU
    from NeuromahRef.src.initializers.Initializer import HeInitializer, XavierInitializer
    ns = {"relu": Activation_ReLU, "xavier": XavierInitializer, "he": HeInitializer, "object": object}
    np.random.seed(0)
    out = [{"cls": c, "kwargs": repr(kw), "outcome": ctor_outcome(getattr(L, c), kw, ns)} for c, kw in CTOR_CASES]
    with open(os.path.join(HERE, "ctor_surface.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (mlp_200, mlp_sgd, cnn_small, vgg_small, cnn_stub, cnn_200, ops, single_layer_optimizers, dropout_mlp, dense_example_stack,
               losses_family, ctor_surface):
        if not only or fn.__name__ in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
