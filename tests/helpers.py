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
