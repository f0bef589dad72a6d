"""Where does the graph-replayed iteration spend its time?  Times FusedTrackingStep (CUDA-graph replay, 100k by default) as it
is and with single stages switched off by monkeypatching (results are then wrong: timing experiment only).

  python tools/step_ablate.py [G] [steps] [variants, comma separated]

  base        the product iteration
  no_priors   the side branch's prior kernels replaced by cached tensors: what the overlap with the render branch costs
  fork_after_fwd / main_hi_prio   where the priors branch forks, and the render branch captured on a high-priority stream"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gs_dynamics_b200 import tracking as TR, workloads

G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def measure(label, patch=None, fork='start', hi_prio=False, cap=None, carve=None, pieces=None):
    if pieces is None:
        os.environ.pop('GSD_PRIORS_PIECES', None)
    else:
        os.environ['GSD_PRIORS_PIECES'] = str(pieces)
    if carve is None:
        os.environ.pop('GSD_PRIORS_CARVEOUT', None)
    else:
        os.environ['GSD_PRIORS_CARVEOUT'] = str(carve)
    if cap is None:
        os.environ.pop('GSD_PRIORS_CTAS_PER_SM', None)
    else:
        os.environ['GSD_PRIORS_CTAS_PER_SM'] = str(cap)
    params, variables, opt, dataset, _ = workloads.tracking_problem_gpu(G, 0, dev)
    orig = TR._track_priors_launch
    if patch == "no_priors":
        cache = {}

        def fake(*a, **k):
            if "v" not in cache:
                cache["v"] = orig(*a, **k)
            return cache["v"]
        TR._track_priors_launch = fake
    if patch == "copy_instead":     # the priors replaced by a plain 44 MB -> 44 MB copy: same DRAM traffic / L2 footprint, no arithmetic
        cache = {"src": torch.zeros(44 * 1024 * 1024 // 4, device=dev), "dst": torch.empty(44 * 1024 * 1024 // 4, device=dev)}

        def fake(*a, **k):
            if "v" not in cache:
                cache["v"] = orig(*a, **k)
            cache["dst"].copy_(cache["src"])
            return cache["v"]
        TR._track_priors_launch = fake
    try:
        step = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=True)
        step.priors_fork = fork
        if hi_prio:
            step.capture_stream = torch.cuda.Stream(priority=-1)
        step.prepare()
    finally:
        TR._track_priors_launch = orig
    n_cams = len(dataset)
    for i in range(20):
        step.step(i % n_cams)
    tot = 0.0
    for i in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step.step((i * 7 + 1) % n_cams)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print("%-12s %8.1f us / iteration  (%6.0f it/s)" % (label, 1e3 * tot / steps, steps / tot * 1e3), flush=True)


def priors_alone():
    params, variables, opt, dataset, _ = workloads.tracking_problem_gpu(G, 0, dev)
    x = params['means3D'].detach()
    q = torch.nn.functional.normalize(params['unnorm_rotations'].detach())
    w = (200.0, 4.0, 1000.0, TR.FLOOR_WEIGHT, 200.0)
    for _ in range(3):
        TR._track_priors_launch(x, q, variables, w)
    tot = 0.0
    for _ in range(50):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        TR._track_priors_launch(x, q, variables, w)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print("priors alone (memsets + prep + fg + finish, eager, cold L2): %.1f us" % (1e3 * tot / 50), flush=True)


variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["base", "no_priors", "base"]
priors_alone()
for vname in variants:
    if vname == "base":
        measure("base")
    elif vname == "no_priors":
        measure("no_priors", "no_priors")
    elif vname == "copy_instead":
        measure("copy_instead", "copy_instead")
    elif vname == "after_fwd":
        measure("fork_after_fwd", fork='after_forward')
    elif vname == "hi_prio":
        measure("main_hi_prio", hi_prio=True)
    elif vname == "both":
        measure("after_fwd+hi_prio", fork='after_forward', hi_prio=True)
    elif vname == "no_morton":   # packed priors tables in the caller's (random) order
        orig_m = TR._morton_order
        TR._morton_order = lambda pts: torch.arange(pts.shape[0], device=pts.device)
        measure("no_morton")
        TR._morton_order = orig_m
    elif vname.startswith("hi_carve"):      # render branch on a high-priority stream + shared-memory carveout of the priors kernel
        measure(vname, hi_prio=True, carve=int(vname[8:]))
    elif vname.startswith("carve"):
        measure(vname, carve=int(vname[5:]))
    elif vname.startswith("pieces"):        # the priors kernel launched in N consecutive pieces
        measure(vname, pieces=int(vname[6:]))
    elif vname.startswith("cap"):
        measure(vname, cap=int(vname[3:]))
