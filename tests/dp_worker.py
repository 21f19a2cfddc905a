"""One rank of the multi-GPU data-parallel checks (launched by tests/test_gpu_dp.py under torch.distributed.run).

Checks, per mode (fp32-exact and the bf16 tensor-core plan):
  1. the fused peer-memory exchange (nm_dp_allreduce_adam_dev_f32) leaves every rank with BIT-IDENTICAL parameters;
  2. its summed gradient equals the NCCL all-reduce's (same addends, different association: fp32 rounding only);
  3. `world` ranks on shards of a global batch land on the single-process full-batch gradient (SURVEY.md 8e rule);
  4. graph replay of the data-parallel step gives the same parameters as eager launches.
Prints DP_OK on rank 0 when everything holds.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from conftest import synth
    from neuromah_b200 import device, parallel
    device.init(local)
    comm = parallel.Communicator(rank, world)
    b = 64
    B = b * world
    steps = 4
    xs, ys = [], []
    for s in range(steps):
        x, y = synth(300 + s, B, (1, 28, 28))
        xs.append(x); ys.append(np.argmax(y, axis=1).astype(np.int32))
    lo, hi = parallel.shard_bounds(B, rank, world)

    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.numpy() for o in out]

    for mode in (None, "bf16"):
        spec, params = bench.cnn_spec_and_params(seed=5)

        def run(kind, graph=False):
            model = bench.build_model(spec, params, mode=mode)
            model.graph_steps = graph
            if kind == "single":
                parallel.set_global_batch(model, B)
                sl = slice(0, B)
            else:
                parallel.attach(model, comm, B, peer=(kind == "peer"))
                assert kind == "nccl" or comm.peer is not None, "peer-memory exchange unavailable on this box"
                sl = slice(lo, hi)
            grads = []
            xd = [device.DeviceArray.from_numpy(x[sl]) for x in xs]
            yd = [device.DeviceArray.from_numpy(y[sl]) for y in ys]
            reps = 3 if graph else 1                       # graph mode: the second sighting of a pointer captures, the third replays
            for _ in range(reps):
                for s in range(steps):
                    if kind == "nccl":
                        saved, comm.peer = comm.peer, None     # force the NCCL path on the same communicator
                    try:
                        model.train_step(xd[s], yd[s])
                    finally:
                        if kind == "nccl":
                            comm.peer = saved
                    if not graph:
                        grads.append(model.arena.grads.numpy()[:model.arena.total].copy())
            device.synchronize()
            comm.check()
            return model.arena.params.numpy()[:model.arena.total].copy(), grads

        p_peer, g_peer = run("peer")
        p_nccl, g_nccl = run("nccl")
        p_one, g_one = run("single")
        # 1. replicas bit-identical
        for other in gather(p_peer):
            assert np.array_equal(other, p_peer), f"{mode}: replicas diverged under the peer-memory exchange"
        # 2. same sum as NCCL up to fp32 association
        # (first step only: afterwards the two runs' weights differ in the last bits, which Adam amplifies)
        np.testing.assert_allclose(g_peer[0], g_nccl[0], rtol=2e-5, atol=1e-7, err_msg=f"{mode}: peer vs NCCL gradient")
        # 3. shards of the global batch == the single-process full batch (first step: identical weights going in)
        tol = dict(rtol=1e-4, atol=1e-5) if mode is None else dict(rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(g_peer[0], g_one[0], err_msg=f"{mode}: data-parallel vs single-process gradient", **tol)
        # Adam turns a gradient that is zero up to rounding into an update of up to +-lr, so parameters are compared loosely
        assert np.max(np.abs(p_peer - p_one)) <= steps * 2 * 1.01e-3, f"{mode}: parameters drifted from the single-process run"
        assert np.max(np.abs(p_peer - p_nccl)) <= steps * 2 * 1.01e-3
        if rank == 0:
            d = np.abs(p_peer - p_one)
            print(f"[{mode or 'fp32'}] |params(dp) - params(single)|: median {np.median(d):.2e}  p99 {np.quantile(d, 0.99):.2e}  max {d.max():.2e}", flush=True)
        # 4. graph replay (bf16 plan only) == eager
        if mode == "bf16":
            p_graph, _ = run("peer", graph=True)
            p_eager3 = None
            # the graph run makes 3 passes over the batches; repeat eagerly for the same number of steps
            model_steps = 3 * steps
            model = bench.build_model(spec, params, mode=mode)
            model.graph_steps = False
            parallel.attach(model, comm, B, peer=True)
            xd = [device.DeviceArray.from_numpy(x[lo:hi]) for x in xs]
            yd = [device.DeviceArray.from_numpy(y[lo:hi]) for y in ys]
            for i in range(model_steps):
                model.train_step(xd[i % steps], yd[i % steps])
            device.synchronize()
            p_eager3 = model.arena.params.numpy()[:model.arena.total].copy()
            np.testing.assert_allclose(p_graph, p_eager3, rtol=1e-4, atol=2e-5, err_msg="graph replay vs eager (data parallel)")
            for other in gather(p_graph):
                assert np.array_equal(other, p_graph), "replicas diverged under graph replay"
    dist.barrier()
    if rank == 0:
        print("DP_OK", flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
