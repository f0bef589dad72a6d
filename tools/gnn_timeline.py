"""Timeline of ONE CUDA-graph replay of the GNN rollout step (sloth cfg, 2 000 particles + tool): kernels of both branches with
their start offsets and durations, from CUPTI activity records through torch.profiler (see tools/graph_timeline.py).

  python tools/gnn_timeline.py [replays] > profiles/r2_gnn_timeline.txt"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gs_dynamics_b200 import gnn, workloads as GO

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cfg = GO.sloth_cfg(512)
dev = torch.device('cuda')
gi = GO.make_graph_inputs(2000, 1, 'sloth')
model = gnn.DynamicsPredictor(dict(cfg), dev).to(dev).eval()
model.load_state_dict(GO.make_state_dict(cfg, 0, head_scale=1e-3))
ro = gnn.GnnRollout(model, gi['state'][0, :, :2000].to(dev), gi['state'][0, :, 2000:].to(dev), 0.075, 8, True, use_graph=True)
d = torch.tensor([0.005, 0, 0], device=dev)
for _ in range(5):
    ro.step(d)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(n):
        ro.step(d)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and ("gsd_" in e.name or "emcpy" in e.name or "emset" in e.name)]
ev.sort(key=lambda e: e.time_range.start)
per = len(ev) // n
its = [ev[k * per:(k + 1) * per] for k in range(n)] if per * n == len(ev) else [ev]
print("# %d replays of %d device activities; start offset / duration in us (mean over replays, matched by launch order)" % (len(its), len(its[0])))
t_end = 0.0
for k in range(len(its[0])):
    st = sum(it[k].time_range.start - it[0].time_range.start for it in its) / len(its)
    du = sum(it[k].time_range.end - it[k].time_range.start for it in its) / len(its)
    t_end = max(t_end, st + du)
    print("%8.1f %8.1f  %s" % (st, du, re.sub(r"\(.*", "", its[0][k].name).replace("void ", "")[:70]))
print("# first start -> last end: %.1f us" % t_end)
