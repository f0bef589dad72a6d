"""Times the first-frame (t = 0) tracking iteration — get_loss(is_initial_timestep=True) + backward + FusedAdam on all parameter
groups, eager launches (the t = 0 path is not graph-capturable), no densification — at G Gaussians:  python tools/t0_bench.py [G]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gs_dynamics_b200 import tracking as TR


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    dev = torch.device("cuda", 0)
    params, v, opt, dataset, _ = bench.build_gpu_problem(G, 0, dev)
    opt = TR.initialize_optimizer(params, v)          # first-frame learning rates: every group live
    step = TR.TrackingStep(params, v, opt, dataset, is_initial_timestep=True, use_graph=False)
    step.prepare()
    rng = random.Random(0)
    for _ in range(10):
        step.step(rng.randrange(len(dataset)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        step.step(rng.randrange(len(dataset)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("t=0 iteration (all groups, two renders, autograd path) at G=%d: %.3f ms = %.0f it/s" % (G, ms, 1e3 / ms))


if __name__ == "__main__":
    main()
