"""Timeline of ONE CUDA-graph replay of the steady-state tracking iteration (kernels of both branches with their start offsets
and durations as they really overlap), from CUPTI activity records through torch.profiler — ncu serialises kernels and nsys is
not in the image.  Averages over the profiled replays; a number taken under the profiler is not a bench value.

  python tools/graph_timeline.py [G] [replays] > profiles/r2_graph_timeline_100k.txt"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gs_dynamics_b200 import tracking as TR, workloads

G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda")
params, variables, opt, dataset, _ = workloads.tracking_problem_gpu(G, 0, dev)
step = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=True)
if os.environ.get("TL_FORK"):          # experiment switches (see tools/step_ablate.py)
    step.priors_fork = os.environ["TL_FORK"]
if os.environ.get("TL_HI_PRIO"):
    step.capture_stream = torch.cuda.Stream(priority=-1)
step.prepare()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for i in range(20):
    step.step(i % 4)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(n):
        flush.zero_()
        step.step(0)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "gsd_" in e.name]
ev.sort(key=lambda e: e.time_range.start)
# split into replays: every replay launches the same kernels; the torch.cuda.synchronize() between replays separates them in time
per = len(ev) // n
its = [sorted(ev[k * per:(k + 1) * per], key=lambda e: e.time_range.start) for k in range(n)] if per * n == len(ev) else [ev]
print("# %d replays of %d kernels, G = %d; start offset / duration in us (mean over replays); '|' marks the side branch" % (len(its), len(its[0]), G))
# kernels are matched by NAME across replays (every kernel of the iteration occurs once; concurrent branches may start in a
# different order from replay to replay), rows are ordered by mean start
t_end = 0.0
rows = []
names = [re.sub(r"\(.*", "", e.name).replace("void ", "") for e in its[0]]
for nm in names:
    st = du = 0.0
    for it in its:
        t0 = min(e.time_range.start for e in it)
        e = next(e for e in it if re.sub(r"\(.*", "", e.name).replace("void ", "") == nm)
        st += e.time_range.start - t0
        du += e.time_range.end - e.time_range.start
    st /= len(its); du /= len(its)
    rows.append((st, du, nm))
    t_end = max(t_end, st + du)
rows.sort()
prev_end = 0.0
for st, du, name in rows:
    side = any(s in name for s in ("gsd_track_node_prep", "gsd_track_fg", "gsd_track_bg", "gsd_track_finish", "gsd_ssim_finish", "gsd_blend_bwd_prefix"))
    gap = "" if side else "  (gap %5.1f)" % (st - prev_end)
    if not side:
        prev_end = st + du
    print("%8.1f %8.1f  %s %s%s" % (st, du, "|" if side else " ", name, gap))
print("# first kernel start -> last kernel end: %.1f us" % t_end)
