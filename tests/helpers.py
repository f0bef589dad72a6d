"""Shared builders for the rasterizer tests: seeded scenes + cameras, oracle calls."""
import numpy as np
import torch

from gs_dynamics_b200 import scenes


def make_camera(cam_id=0, w=640, h=480, near=0.01, far=100.0):
    W0, H0, cams = scenes.demo_cameras()
    k, w2c = cams[cam_id]
    k = k.copy()
    k[0] *= w / W0
    k[1] *= h / H0
    return scenes.camera_matrices(w, h, k, w2c, near, far)


def make_scene(G, seed=0, scale_boost=0.0, box_scale=1.0):
    sc = scenes.synthetic_scene(G, seed, box_scale=box_scale)
    if scale_boost:
        sc["log_scales"] = sc["log_scales"] + scale_boost
    return sc, scenes.activate(sc)


def oracle_forward(act, cam, bg, colors=None, debug=False):
    from oracle import raster_c
    col = act["colors_precomp"] if colors is None else colors
    return raster_c.forward(act["means3D"], col, act["opacities"], act["scales"], act["rotations"],
                            cam["viewmatrix"], cam["projmatrix"], bg, cam["tanfovx"], cam["tanfovy"],
                            cam["image_height"], cam["image_width"], debug=debug)


def oracle_backward(act, cam, bg, dL, colors=None):
    from oracle import raster_c
    col = act["colors_precomp"] if colors is None else colors
    return raster_c.backward(act["means3D"], col, act["opacities"], act["scales"], act["rotations"],
                             cam["viewmatrix"], cam["projmatrix"], bg, cam["tanfovx"], cam["tanfovy"],
                             cam["image_height"], cam["image_width"], dL)


def settings_from(cam, bg, device="cuda"):
    from gs_dynamics_b200.rasterizer import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=cam["image_height"], image_width=cam["image_width"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
        bg=torch.as_tensor(bg, dtype=torch.float32).to(device), scale_modifier=1.0,
        viewmatrix=cam["viewmatrix"].to(device), projmatrix=cam["projmatrix"].to(device), sh_degree=0,
        campos=cam["campos"].to(device), prefiltered=False)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def analytic_two_gaussians(w=64, h=64, f=100.0):
    """Known-answer case computable by hand from the published splatting formulas (Kerbl et al. 2023, EWA splatting; the
    constants of SURVEY.md §2.1): camera at the origin looking down +z with the principal point at the image centre, two
    ISOTROPIC Gaussians on the optical axis at depths z0 < z1.  Then J = diag(f/z), cov2D = ((f s / z)^2 + 0.3) I, the
    projected centre is pixel (w-1)/2, (h-1)/2 and for a pixel at squared distance d2 from it
        alpha_i = min(0.99, o_i * exp(-d2 / (2 ((f s_i / z_i)^2 + 0.3)))),   dropped when < 1/255
        colour  = a0 c0 + (1 - a0) a1 c1 + (1 - a0)(1 - a1) bg,   depth = a0 z0 + (1 - a0) a1 z1,   T = (1 - a0)(1 - a1)
    (no pixel gets near the 1e-4 termination).  Returns the inputs and the expected maps, with a validity mask that leaves out
    the pixels within 2 % of the 1/255 threshold."""
    import numpy as np
    import torch
    from gs_dynamics_b200 import scenes
    k = [[f, 0.0, w / 2.0], [0.0, f, h / 2.0], [0.0, 0.0, 1.0]]
    cam = scenes.camera_matrices(w, h, k, np.eye(4), near=0.01)
    z = np.array([1.0, 1.5])
    s = np.array([0.03, 0.06])
    o = np.array([0.7, 0.9])
    c = np.array([[0.9, 0.2, 0.1], [0.1, 0.5, 0.8]])
    bg = np.array([0.05, 0.10, 0.15])
    act = dict(means3D=torch.tensor([[0.0, 0.0, z[0]], [0.0, 0.0, z[1]]], dtype=torch.float32),
               colors_precomp=torch.tensor(c, dtype=torch.float32), opacities=torch.tensor(o[:, None], dtype=torch.float32),
               scales=torch.tensor(np.repeat(s[:, None], 3, 1), dtype=torch.float32),
               rotations=torch.tensor([[1.0, 0, 0, 0], [1.0, 0, 0, 0]], dtype=torch.float32))
    ys, xs = np.mgrid[0:h, 0:w]
    d2 = (xs - (w - 1) / 2.0) ** 2 + (ys - (h - 1) / 2.0) ** 2
    var = (f * s / z) ** 2 + 0.3
    a = [np.minimum(0.99, o[i] * np.exp(-d2 / (2 * var[i]))) for i in range(2)]
    near_thr = np.zeros_like(d2, dtype=bool)
    for i in range(2):
        near_thr |= np.abs(a[i] - 1 / 255.0) < 0.02 / 255.0
        a[i] = np.where(a[i] >= 1 / 255.0, a[i], 0.0)
    T = (1 - a[0]) * (1 - a[1])
    color = a[0][None] * c[0][:, None, None] + ((1 - a[0]) * a[1])[None] * c[1][:, None, None] + T[None] * bg[:, None, None]
    depth = a[0] * z[0] + (1 - a[0]) * a[1] * z[1]
    radii = np.ceil(3.0 * np.sqrt(var)).astype(np.int32)
    return dict(cam=cam, act=act, bg=torch.tensor(bg, dtype=torch.float32), color=color, depth=depth, final_T=T, radii=radii,
                valid=~near_thr)


def analytic_one_gaussian_gradients(w=48, h=48, f=90.0, seed=0):
    """Closed-form GRADIENTS for one isotropic Gaussian on the optical axis (camera as in analytic_two_gaussians) and a seeded
    upstream gradient dL/dC.  With G = exp(-(dx^2/vx + dy^2/vy)/2), v. = (f s./z)^2 + 0.3, a = o G < 0.99 and
    C = a c + (1 - a) bg on the pixels with a >= 1/255 (E = sum_ch dL (c - bg) restricted to them):
        dL/dc   = sum_px a dL                      dL/do  = sum_px G E
        dL/dX   = (f/z) sum_px a E dx / vx          (dx = px - centre; cov2D is stationary in X, Y on the axis)
        dL/dsx  = sum_px a E (dx^2 / (2 vx^2)) * 2 (f/z)^2 sx,   same for y;   dL/dsz = 0  (J's third column vanishes on the axis)
    Pixels within 2 % of the 1/255 threshold get dL = 0 so that the contributor set is unambiguous."""
    import numpy as np
    import torch
    from gs_dynamics_b200 import scenes
    k = [[f, 0.0, w / 2.0], [0.0, f, h / 2.0], [0.0, 0.0, 1.0]]
    cam = scenes.camera_matrices(w, h, k, np.eye(4), near=0.01)
    z, s, o = 1.2, 0.05, 0.6
    c = np.array([0.8, 0.3, 0.2])
    bg = np.array([0.1, 0.4, 0.05])
    act = dict(means3D=torch.tensor([[0.0, 0.0, z]], dtype=torch.float32), colors_precomp=torch.tensor(c[None], dtype=torch.float32),
               opacities=torch.tensor([[o]], dtype=torch.float32), scales=torch.tensor([[s, s, s]], dtype=torch.float32),
               rotations=torch.tensor([[1.0, 0, 0, 0]], dtype=torch.float32))
    ys, xs = np.mgrid[0:h, 0:w]
    dx, dy = xs - (w - 1) / 2.0, ys - (h - 1) / 2.0
    v = (f * s / z) ** 2 + 0.3
    Gs = np.exp(-(dx * dx + dy * dy) / (2 * v))
    a = o * Gs
    on = a >= 1 / 255.0
    rng = np.random.default_rng(seed)
    dL = rng.normal(size=(3, h, w))
    dL[:, np.abs(a - 1 / 255.0) < 0.02 / 255.0] = 0.0
    E = (dL * (c - bg)[:, None, None]).sum(0) * on
    g = dict(colors=(dL * (a * on)[None]).sum((1, 2)), opacity=(Gs * E).sum(),
             mean=np.array([(f / z) * (a * E * dx / v).sum(), (f / z) * (a * E * dy / v).sum()]),
             scales=np.array([(a * E * dx * dx / (2 * v * v)).sum() * 2 * (f / z) ** 2 * s,
                              (a * E * dy * dy / (2 * v * v)).sum() * 2 * (f / z) ** 2 * s, 0.0]))
    return dict(cam=cam, act=act, bg=torch.tensor(bg, dtype=torch.float32), dL=torch.tensor(dL, dtype=torch.float32), grads=g)


def pct_rel_err(a, b, q=99.9, eps_frac=1e-3):
    """Per-element relative error |a - b| / (|b| + eps) with eps = eps_frac * max|b|, at percentile q and at the maximum:
    unlike rel_err (max-abs error over max-abs value) this does not hide errors on small entries."""
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    if a.size == 0:
        return 0.0, 0.0
    e = np.abs(a - b) / (np.abs(b) + eps_frac * (np.abs(b).max() + 1e-30))
    return float(np.percentile(e, q)), float(e.max())
