"""GNN training iteration (bench.py's gnn_train object) alone, for A/B runs of GSD_TRAIN_GEMM=cublas|tc:
  MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 GSD_TRAIN_GEMM=tc python tools/train_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gs_dynamics_b200 import dist as gdist

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
gdist.init(backend="nccl", device=dev)
r = bench.run_gnn_train(dev, 1)
print(os.environ.get("GSD_TRAIN_GEMM", "cublas"), "graph %.2f ms, eager %.2f ms per iteration, loss %.3g -> %.3g" % (
    r["ms_per_iteration"], r["eager_ms_per_iteration"], r["loss_first_mean_over_ranks"], r["loss_last_mean_over_ranks"]))
