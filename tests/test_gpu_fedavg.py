"""BASELINE.json config 5 on the device: federated_average (Federated/utils.py:20-23) over the flat parameter slab.

One GPU is enough to pin the arithmetic: the per-client pre-scale n_i / N that precedes the SUM all-reduce
(``federated_average_device``; the all-reduce itself is covered at world size 2 by tests/test_parallel_gloo.py on CPU and
by tests/dp_worker.py on GPUs) and the server-side weighted sum over several client slabs resident on one GPU
(``federated_average_slabs``), both against the oracle's restatement of the reference rule; plus a full FedAvg round of
real local training (FLTrainer) whose aggregate is checked against the rule applied to the clients' reported weights."""
import numpy as np
import pytest

from conftest import synth_learnable
from oracle import np_ref as R
from oracle import ref_net
from test_gpu_model import build

pytestmark = pytest.mark.gpu


def test_prescale_of_the_device_fedavg_is_the_reference_weight():
    from neuromah_b200.Federated.utils import federated_average_device
    spec = ref_net.mnist_cnn_spec()
    m = build(spec, ref_net.init_params(spec, np.random.RandomState(1), init="he"))
    w = m.arena.host_params().astype(np.float64)
    federated_average_device(m.arena, 300, 1200, None)             # this client's term of the sum: w * n_i / N
    ref = np.asarray(R.federated_average([w, np.zeros_like(w)], [300, 900]))
    np.testing.assert_allclose(m.arena.host_params(), ref, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("alias_out", [False, True])
def test_slab_fedavg_matches_reference_rule(alias_out):
    from neuromah_b200.Federated.utils import federated_average_slabs
    spec = ref_net.mnist_cnn_spec()
    clients = [build(spec, ref_net.init_params(spec, np.random.RandomState(10 + i), init="he")) for i in range(4)]
    counts = [4096, 1000, 37, 8192]
    updates = [c.arena.host_params().astype(np.float64) for c in clients]
    out = clients[2] if alias_out else build(spec, ref_net.init_params(spec, np.random.RandomState(99)))
    federated_average_slabs(out.arena, [c.arena for c in clients], counts)
    ref = np.asarray(R.federated_average(updates, counts))
    np.testing.assert_allclose(out.arena.host_params(), ref, rtol=2e-6, atol=1e-8)
    # the layers' weight views see the aggregate (they alias the slab)
    np.testing.assert_allclose(np.asarray(out.layers[0].weights).ravel(), ref[:288], rtol=2e-6, atol=1e-8)


def test_fedavg_round_with_real_local_training():
    """Three clients (FLTrainer with a model attached) train locally on their own shards from the same global weights,
    report ({"weights": list}, num_samples) as the reference's FLTrainer does (FLTrainer.py:25-31), and the aggregate equals
    federated_average on those reports; the averaged model then evaluates far better than chance (0.1) on every shard."""
    from neuromah_b200.Federated.FLTrainer import FLTrainer
    from neuromah_b200.Federated.utils import federated_average, federated_average_slabs
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(3), init="he")
    sizes = [512, 256, 384]
    models = [build(spec, params, mode="bf16", learning_rate=0.002) for _ in sizes]
    global_w = models[0].arena.host_params()
    reports, shards = [], []
    for i, (m, n) in enumerate(zip(models, sizes)):
        x, y = synth_learnable(40 + i, n)
        shards.append((x, y))
        m.epoch_end_accuracy = False
        reports.append(FLTrainer({"weights": global_w.tolist()}, model=m, data=(x, y), batch_size=128, local_epochs=3).train_local_model())
    assert [r[1] for r in reports] == sizes
    host_avg = np.asarray(federated_average([r[0]["weights"] for r in reports], sizes))
    federated_average_slabs(models[0].arena, [m.arena for m in models], sizes)
    np.testing.assert_allclose(models[0].arena.host_params(), host_avg, rtol=2e-6, atol=1e-8)
    for x, y in shards:
        assert models[0].evaluate(x, y, batch_size=128, verbose=0)["acc"] > 0.3
