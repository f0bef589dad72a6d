"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): rasterizer forward + both backward modes,
one fused tracking iteration, tile sort with long lists, the tcgen05 GEMM.  python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gs_dynamics_b200 import rasterizer as R, tracking as TR, workloads, gnn, scenes
from tests.helpers import make_camera, make_scene, settings_from

dev = torch.device("cuda")
# rasterizer: ragged image, long tile lists (multi-chunk, look-back, replay), both colour sets
cam = make_camera(1, 100, 70)
sc, act = make_scene(6000, 3, scale_boost=0.8, box_scale=0.5)
a = {k: v.cuda() for k, v in act.items()}
st = settings_from(cam, [0.1, 0.2, 0.3])
c, r, d, s = R.raster_forward(st, a["means3D"], a["opacities"], a["colors_precomp"], a["scales"], a["rotations"], colors1=sc["seg_colors"].cuda())
dL = torch.randn_like(c)
g1 = R.raster_backward(s, dL)
g2 = R.raster_backward(s, dL, need_means2D=False, geom_only=True)
cam2 = make_camera(0, 48, 32)
sc2, act2 = make_scene(20000, 5, box_scale=0.25)
a2 = {k: v.cuda() for k, v in act2.items()}
c2, _, _, s2 = R.raster_forward(settings_from(cam2, [0, 0, 0]), a2["means3D"], a2["opacities"], a2["colors_precomp"], a2["scales"], a2["rotations"])
R.raster_backward(s2, torch.randn_like(c2))
# one fused tracking iteration (eager) at a small size
params, variables, opt, dataset, _ = workloads.tracking_problem_gpu(3000, 0, dev)
step = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=False)
step.prepare([0])
step.step(0)
# targets from bytes (unpack kernel) and the two-launch variant of the backward (gsd_raster_backward + gsd_track_update)
H, W = dataset[0]['im'].shape[1:]
step.set_target_u8(0, torch.randint(0, 256, (H, W, 3), dtype=torch.uint8).pin_memory(), torch.randint(0, 2, (H, W), dtype=torch.uint8).pin_memory())
step.step(0)
step.fuse_update = False
step.prefix_on_side = False
step.step(0)
# tcgen05 GEMM, both tile widths
for M, N in ((300, 128), (700, 512)):
    x = torch.randn(M, 64, device=dev); w = torch.randn(N, 64, device=dev)
    y = gnn._tc_linear(x, gnn._tc_split(w), torch.randn(N, device=dev), relu=True)
torch.cuda.synchronize()
print("sanitize case done", float(c.sum()), float(g1["means3D"].abs().sum()), float(y.sum()))
