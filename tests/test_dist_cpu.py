"""CPU: the N > 1 host path with world_size-2 gloo: episode sharding without a data-path collective, barrier and
max-over-ranks timing reduction (what bench.py does under torchrun on GPUs with NCCL)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from gs_dynamics_b200 import dist as gdist, workloads
    r, lr, w = gdist.init(backend="gloo")
    mine = gdist.shard(5, r, w)                       # 5 episodes over 2 ranks
    # every rank builds only its own episodes (seed = episode id): no parameter is shared, no gradient exchange
    sizes = [workloads.tracking_problem(64, seed=e, num_knn=3)["params"]["means3D"].shape[0] for e in mine]
    gdist.barrier()
    t_local = 1.0 + rank                               # pretend device time
    t_max, = gdist.reduce_max([t_local])
    n_total, = gdist.reduce_sum([float(len(mine))])
    q.put((rank, mine, sizes, t_max, n_total))
    gdist.finalize()


def test_two_rank_episode_sharding_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]
    assert all(r[3] == 2.0 for r in res)               # max over ranks
    assert all(r[4] == 5.0 for r in res)               # every unit processed exactly once


def _bucket_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from gs_dynamics_b200 import dist as gdist, gnn_train
    gdist.init(backend="gloo")
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))   # same init on both ranks
    bucket = gnn_train.GradientBucket(lin.parameters())
    bucket.zero()
    x = torch.full((4, 5), float(rank + 1))
    lin(x).sum().backward()                                  # rank-dependent gradients, written into the flat bucket
    local = bucket.flat.clone()
    bucket.all_reduce_mean()                                 # the one data-path collective of DP GNN training
    q.put((rank, local.tolist(), bucket.flat.tolist(), [p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket._views())]))
    gdist.finalize()


def test_gradient_bucket_allreduce_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    mean = (torch.tensor(res[0][1]) + torch.tensor(res[1][1])) / 2
    for r in res:
        assert torch.allclose(torch.tensor(r[2]), mean, atol=1e-6)
        assert all(r[3])                                     # parameter .grad tensors alias the flat bucket
