"""GPU-side diagnostics for the rasterizer: stage-by-stage comparison against the CPU oracle + first timings.
Writes a report to gpurun_out/diag_raster.txt. Development tool (not part of the product or the test suite)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gs_dynamics_b200 import rasterizer as R, scenes
from tests.helpers import make_camera, make_scene, oracle_forward, oracle_backward, settings_from, rel_err

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "diag_raster.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); out.write(s + "\n"); out.flush()

def align(x, a=256): return (x + a - 1) // a * a

def decode_geom(buf, G):
    b = buf.cpu().numpy(); off = 0; res = {}
    for name, dt, n in (("xy", np.float32, 2), ("conic_o", np.float32, 4), ("ext", np.float32, 2), ("depth", np.float32, 1),
                        ("rect", np.uint32, 2), ("tiles", np.uint32, 1), ("offsets", np.uint32, 1)):
        nb = G * n * 4
        res[name] = b[off:off + nb].view(dt).reshape(G, n) if n > 1 else b[off:off + nb].view(dt)
        off += align(nb)
    return res

def stage_compare(G, w, h, seed, boost, box):
    P("=== stage compare G=%d %dx%d seed=%d" % (G, w, h, seed))
    cam = make_camera(0, w, h); sc, act = make_scene(G, seed, scale_boost=boost, box_scale=box)
    bg = [0.1, 0.2, 0.3]
    fo = oracle_forward(act, cam, torch.tensor(bg), debug=True)
    a = {k: v.cuda() for k, v in act.items()}
    st = settings_from(cam, bg)
    color, radii, depth, state = R.raster_forward(st, a["means3D"], a["opacities"], a["colors_precomp"], a["scales"], a["rotations"])
    torch.cuda.synchronize()
    P("status", state.status.cpu().tolist(), "oracle R", fo["R"])
    geom = decode_geom(state.keep[10], G)
    vis = fo["radii"] > 0
    P("radii mismatches", int((radii.cpu().numpy() != fo["radii"]).sum()), "visible", int(vis.sum()))
    P("tiles mismatches", int((geom["tiles"] != fo["tiles_touched"]).sum()))
    P("xy maxdiff", float(np.abs(geom["xy"][vis] - fo["xy"][vis]).max()), "conic maxrel",
      float((np.abs(geom["conic_o"][vis] - fo["conic_o"][vis]) / (np.abs(fo["conic_o"][vis]) + 1e-12)).max()),
      "depth maxdiff", float(np.abs(geom["depth"][vis] - fo["gdepth"][vis]).max()))
    d = np.abs(color.cpu().numpy() - fo["color"])
    P("color maxdiff", float(d.max()), "frac>1e-4", float((d > 1e-4).mean()), "frac>1e-5", float((d > 1e-5).mean()))
    if d.max() > 1e-3:
        idx = np.unravel_index(np.argmax(d), d.shape); P("worst at", idx, "gpu", float(color.cpu().numpy()[idx]), "oracle", float(fo["color"][idx]))
        ys, xs = np.where(d.max(0) > 1e-3); P("bad px count", len(ys), "first", list(zip(ys[:10].tolist(), xs[:10].tolist())))
    dd = np.abs(depth.cpu().numpy() - fo["depth"]); P("depth img maxdiff", float(dd.max()))
    g = torch.Generator().manual_seed(seed); dL = torch.randn(3, h, w, generator=g)
    bo = oracle_backward(act, cam, torch.tensor(bg), dL)
    gr = R.raster_backward(state, dL.cuda()); torch.cuda.synchronize()
    for k, ko in (("means3D", "means3D"), ("means2D", "means2D"), ("colors0", "colors"), ("opacities", "opacities"), ("scales", "scales"), ("rotations", "rotations")):
        gg = gr[k].cpu().numpy().reshape(bo[ko].shape)
        P("grad", k, "rel", rel_err(gg, bo[ko]), "finite", bool(np.isfinite(gg).all()), "max", float(np.abs(bo[ko]).max()))

def timing(G, n_sets, iters=30):
    cam = make_camera(0, 640, 480); sc, act = make_scene(G, 0)
    a = {k: v.cuda() for k, v in act.items()}; st = settings_from(cam, [0, 0, 0])
    seg = sc["seg_colors"].cuda() if n_sets == 2 else None
    c, r, d, s = R.raster_forward(st, a["means3D"], a["opacities"], a["colors_precomp"], a["scales"], a["rotations"], colors1=seg)
    cap = int(s.status[0].item() * 1.25)
    dL = torch.randn(3 * n_sets, 480, 640, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for it in range(iters + 5):
        ev[0].record()
        c, r, d, s = R.raster_forward(st, a["means3D"], a["opacities"], a["colors_precomp"], a["scales"], a["rotations"], colors1=seg, capacity=cap)
        ev[1].record()
        g = R.raster_backward(s, dL)
        ev[2].record(); torch.cuda.synchronize()
        if it >= 5: tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    P("timing G=%d n_sets=%d R=%d: fwd %.1f us  bwd %.1f us" % (G, n_sets, int(s.status[0].item()), 1000 * tf / iters, 1000 * tb / iters))

if __name__ == "__main__":
    P(torch.cuda.get_device_name(0))
    stage_compare(2000, 128, 96, 1, 1.2, 0.6)
    stage_compare(20000, 640, 480, 3, 0.0, 1.0)
    for G in (50000, 100000):
        for ns in (1, 2):
            timing(G, ns)
