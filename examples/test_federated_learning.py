"""BASELINE.json config 5: FedAvg with one client per GPU, each training the MNIST CNN locally on its own shard with the
B200 path, then `federated_average` (Federated/utils.py:20-23) as a pre-scaled SUM all-reduce of the weight slab over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        examples/test_federated_learning.py [--rounds 3] [--local-steps 4] [--batch 4096]
(N = 1 works too: a single client.)
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from neuromah_b200 import device, parallel
from neuromah_b200.Federated.utils import federated_average_device

ap = argparse.ArgumentParser()
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--local-steps", type=int, default=4)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
device.init(local_rank)
spec, params = bench.cnn_spec_and_params(seed=0)
model = bench.build_model(spec, params, mode=None if args.mode == "fp32" else "bf16")
comm = parallel.Communicator(rank, world)
if world > 1:
    comm.broadcast(model.arena.params, 0, model.arena.total)            # every client starts from the global model
rng = np.random.default_rng(100 + rank)                                  # each client owns a different shard
n = args.local_steps * args.batch
labels = rng.integers(0, 10, n)
X = rng.random((n, 1, 28, 28), dtype=np.float32) * 0.5
for k in range(10):
    X[labels == k, 0, 2 * k + 4:2 * k + 7, 6:22] += 0.5
model.epoch_end_accuracy = False
for rnd in range(1, args.rounds + 1):
    t0 = time.perf_counter()
    model.train(X, labels, epochs=1, batch_size=args.batch, verbose=0)                       # local training (Adam state stays local)
    federated_average_device(model.arena, n, n * world, comm if world > 1 else None)          # sum_i w_i * n_i / N
    device.synchronize()
    dt = time.perf_counter() - t0
    w = model.arena.host_params()
    if rank == 0:
        print(f"round {rnd}: {world} client(s) x {args.local_steps} local steps of batch {args.batch}: {dt * 1e3:.1f} ms, "
              f"local loss {model.last_epoch['loss']:.4f}, |w| = {float(np.linalg.norm(w)):.5f}", flush=True)
if world > 1:
    import torch
    t = torch.from_numpy(model.arena.host_params()).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("all clients hold the same averaged weights:", bool(torch.equal(lo, hi)))
    dist.destroy_process_group()
