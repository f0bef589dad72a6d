"""Synthetic workloads of SURVEY.md §8(d): the 4 calibrated demo cameras and the scene S(G, seed).

Pure numpy / CPU torch; used by bench.py, smoke() and the tests to build identical seeded inputs for the
CUDA path and for the oracle.
"""
import json
import os

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def demo_cameras():
    """The 4 cameras of the reference's assets/demo (w2c, OpenCV convention), 640x480."""
    with open(os.path.join(_DATA, "demo_cameras.json")) as f:
        d = json.load(f)
    return d["w"], d["h"], [(np.array(c["k"], np.float64), np.array(c["w2c"], np.float64)) for c in d["cams"]]


def camera_matrices(w, h, k, w2c, near=0.01, far=100.0):
    """view/proj matrices with the conventions of the reference's setup_camera
    (/root/reference/src/tracking/helpers.py:10-33): viewmatrix = w2c^T, projmatrix = (P_gl @ w2c)^T,
    both float32 [1,4,4]; tanfov = size / (2 f)."""
    fx, fy, cx, cy = float(k[0][0]), float(k[1][1]), float(k[0][2]), float(k[1][2])
    w2c_t = torch.tensor(np.asarray(w2c), dtype=torch.float32)
    campos = torch.inverse(w2c_t)[:3, 3].contiguous()
    view = w2c_t.t().contiguous().unsqueeze(0)
    p_gl = torch.tensor([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                         [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                         [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                         [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32)
    proj = view.bmm(p_gl.t().unsqueeze(0)).contiguous()
    return dict(image_height=int(h), image_width=int(w), tanfovx=w / (2 * fx), tanfovy=h / (2 * fy),
                viewmatrix=view, projmatrix=proj, campos=campos)


def synthetic_scene(G, seed=0, box_scale=1.0):
    """S(G, seed) of SURVEY.md §8(d): raw (un-activated) tracking parameters as float32 CPU tensors."""
    rng = np.random.default_rng(seed)
    lo = np.array([-0.25, -0.25, -0.10]) * box_scale
    hi = np.array([0.25, 0.25, 0.0]) * box_scale
    means = rng.uniform(lo, hi, size=(G, 3)) + np.array([0.280, 0.073, 0.0])
    sigma = rng.uniform(0.002, 0.006, size=(G, 1))
    log_scales = np.repeat(np.log(sigma), 3, axis=1)
    rots = rng.normal(size=(G, 4))
    logit_op = rng.normal(size=(G, 1))
    rgb = rng.uniform(size=(G, 3))
    seg = np.tile(np.array([[1.0, 0.0, 0.0]]), (G, 1))
    t = lambda a: torch.tensor(a, dtype=torch.float32).contiguous()
    return dict(means3D=t(means), rgb_colors=t(rgb), seg_colors=t(seg), unnorm_rotations=t(rots),
                logit_opacities=t(logit_op), log_scales=t(log_scales))


def activate(params):
    """means/colours/rotations/opacities/scales as the rasterizer consumes them
    (/root/reference/src/tracking/helpers.py:36-45), CPU reference math."""
    return dict(means3D=params["means3D"], colors_precomp=params["rgb_colors"],
                rotations=torch.nn.functional.normalize(params["unnorm_rotations"]),
                opacities=torch.sigmoid(params["logit_opacities"]), scales=torch.exp(params["log_scales"]))
