"""Short fwd+bwd loop of the rasterizer for ncu (launch list / full capture). Development tool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gs_dynamics_b200 import rasterizer as R
from tests.helpers import make_camera, make_scene, settings_from
G = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
n_sets = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cam = make_camera(0, 640, 480); sc, act = make_scene(G, 0)
a = {k: v.cuda() for k, v in act.items()}; st = settings_from(cam, [0, 0, 0])
seg = sc["seg_colors"].cuda() if n_sets == 2 else None
c, r, d, s = R.raster_forward(st, a["means3D"], a["opacities"], a["colors_precomp"], a["scales"], a["rotations"], colors1=seg)
cap = int(s.status[0].item() * 1.25)
dL = torch.randn(3 * n_sets, 480, 640, device="cuda")
for it in range(iters):
    c, r, d, s = R.raster_forward(st, a["means3D"], a["opacities"], a["colors_precomp"], a["scales"], a["rotations"], colors1=seg, capacity=cap)
    g = R.raster_backward(s, dL)
torch.cuda.synchronize()
print("done", int(s.status[0].item()))
