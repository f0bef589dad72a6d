"""Synthetic benchmark workloads of SURVEY.md §8(d) as CPU tensors (moved to the GPU by the caller).

Config 2 / 3 (tracking): steady-state t > 0 iteration of train_gs.py at fixed G on the 4 demo cameras (640x480):
scene S(G, seed) is optimised towards images of S(G, seed+1); K = 20 neighbours from a KD-tree on the initial means;
prev_* from a 1 mm / 0.01 perturbation of the current state; cameras drawn i.i.d. uniform (random.seed(0)).
"""
import numpy as np
import torch

from . import scenes


def tracking_problem(G, seed=0, num_knn=20, n_cams=4, width=640, height=480):
    from scipy.spatial import cKDTree
    W0, H0, cams = scenes.demo_cameras()
    sc = scenes.synthetic_scene(G, seed)
    target = scenes.activate(scenes.synthetic_scene(G, seed + 1))
    cam_list = []
    for cid in range(n_cams):
        k, w2c = cams[cid % len(cams)]
        k = k.copy()
        k[0] *= width / W0
        k[1] *= height / H0
        cam_list.append(dict(k=k, w2c=w2c, w=width, h=height, id=cid, mats=scenes.camera_matrices(width, height, k, w2c, 1.0, 100.0)))
    params = dict(sc)
    params["cam_m"] = torch.zeros(50, 3)
    params["cam_c"] = torch.zeros(50, 3)
    pts = sc["means3D"].numpy().astype(np.float64)
    d, idx = cKDTree(pts).query(pts, k=num_knn + 1)
    sq, idx = d[:, 1:] ** 2, idx[:, 1:]
    rng = np.random.default_rng(seed + 100)
    rot = torch.nn.functional.normalize(sc["unnorm_rotations"])
    pts_prev = sc["means3D"] + torch.tensor(rng.normal(scale=1e-3, size=(G, 3)), dtype=torch.float32)
    rot_prev = torch.nn.functional.normalize(rot + torch.tensor(rng.normal(scale=1e-2, size=(G, 4)), dtype=torch.float32))
    nbr = torch.tensor(idx, dtype=torch.int64)
    prev_inv = rot_prev.clone()
    prev_inv[:, 1:] *= -1
    variables = dict(
        neighbor_indices=nbr, neighbor_weight=torch.tensor(np.exp(-2000 * sq), dtype=torch.float32),
        neighbor_dist=torch.tensor(np.sqrt(sq), dtype=torch.float32), prev_offset=(pts_prev[nbr] - pts_prev[:, None]).contiguous(),
        prev_inv_rot_fg=prev_inv.contiguous(), prev_pts=pts_prev, prev_rot=rot_prev,
        init_bg_pts=torch.zeros(0, 3), init_bg_rot=torch.zeros(0, 4), max_2D_radius=torch.zeros(G), scene_radius=1.0)
    return dict(G=G, params=params, variables=variables, cams=cam_list, target=target, seg_colors=sc["seg_colors"])


def seg_target_from_mask(mask):
    """(seg, 0, 1-seg) colour coding of the reference's dataset loader (train_utils.py:71-75)."""
    return torch.stack((mask, torch.zeros_like(mask), 1 - mask))
