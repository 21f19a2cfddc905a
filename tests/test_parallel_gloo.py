"""CPU, world_size 2, gloo: the data-parallel normalisation rule.

Each rank runs the ORACLE on its shard with the global batch as divisor, the gradient
vectors are SUM-all-reduced over gloo, and every rank must land on the single-process
batch-B gradients and post-Adam weights.  This pins the host-side sharding / normalisation
logic (neuromah_b200.parallel.shard_bounds + the global-batch rule) without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import synth
    from neuromah_b200.parallel import shard_bounds
    from oracle import ref_net
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(21))
    B = 8
    x, y = synth(22, B, (1, 28, 28))
    lo, hi = shard_bounds(B, rank, world)
    net = ref_net.RefNet(spec, params)
    net.step(x[lo:hi], y[lo:hi], global_batch=B, update=False)
    g = torch.from_numpy(net.flat("grads").copy())
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    # install the reduced gradients and apply the optimizer
    off = 0
    for i, p in enumerate(net.params):
        if p is None:
            continue
        for name in ("weights", "biases"):
            n = p[name].size
            net.grads[i][name] = g[off:off + n].numpy().reshape(net.grads[i][name].shape)
            off += n
    net.update()
    q.put((rank, g.numpy(), net.flat()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sum_allreduce_equals_single_process():
    import torch.multiprocessing as mp
    from conftest import synth
    from oracle import ref_net
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = ref_net.mnist_cnn_spec()
    net = ref_net.RefNet(spec, ref_net.init_params(spec, np.random.RandomState(21)))
    x, y = synth(22, 8, (1, 28, 28))
    net.step(x, y)
    for _, g, w in results:
        np.testing.assert_allclose(g, net.flat("grads"), rtol=1e-4, atol=1e-7)   # f32 conv grads, different sum order
        np.testing.assert_allclose(w, net.flat(), rtol=1e-4, atol=1e-5)              # north-star tolerance


def _fedavg_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(50 + rank)
    w = rng.standard_normal(1000).astype(np.float32)          # this client's weight slab after local training
    n_local = [300, 100][rank]
    n_total = 400
    # federated_average_device's rule (Federated/utils.py:20-23): scale by n_i / N, then SUM all-reduce
    t = torch.from_numpy(w * np.float32(n_local / n_total))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    q.put((rank, w, n_local, t.numpy()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_client_fedavg_equals_reference_rule():
    """Config 5: pre-scaled SUM all-reduce over two clients == federated_average(updates, num_samples)."""
    import torch.multiprocessing as mp
    from oracle import np_ref as R
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fedavg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.asarray(R.federated_average([r[1].astype(np.float64) for r in results], [r[2] for r in results]))
    for _, _, _, avg in results:
        np.testing.assert_allclose(avg, ref, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(results[0][3], results[1][3])     # every client ends with the same model
