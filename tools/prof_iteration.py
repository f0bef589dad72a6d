"""N eager steady-state tracking iterations (FusedTrackingStep, use_graph=False) for ncu: python tools/prof_iteration.py G [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gs_dynamics_b200 import tracking as TR, workloads

G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
params, variables, opt, dataset, _ = workloads.tracking_problem_gpu(G, 0, torch.device("cuda"))
step = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=False)
step.prepare([0])
for i in range(n):
    step.step(0)
torch.cuda.synchronize()
print("done", G, n, step.check_capacity())
