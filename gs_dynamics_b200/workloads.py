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


def tracking_problem_gpu(G, seed, device, prob=None):
    """tracking_problem(G, seed) moved to the device and wired to the library the way train_gs.py leaves it after frame 0
    (train_utils.py:354-374: kNN tables, frozen learning rates): returns (params, variables, optimizer, dataset, host_targets).
    Targets are rendered by the CUDA rasterizer from S(G, seed + 1); host_targets are their pinned host copies."""
    from . import rasterizer as R
    from . import tracking as TR
    prob = tracking_problem(G, seed) if prob is None else prob
    params = {k: torch.nn.Parameter(v.to(device).contiguous()) for k, v in prob["params"].items()}
    params["rgb_colors"].requires_grad = False
    v = {k: (t.to(device).contiguous() if isinstance(t, torch.Tensor) else t) for k, t in prob["variables"].items()}
    v["neighbor_indices_i32"] = v["neighbor_indices"].to(torch.int32).contiguous()
    v["in_ptr"], v["in_edge"] = TR.build_in_edges(v["neighbor_indices_i32"])
    v["fg_index"] = None
    v["bg_index"] = torch.zeros(0, dtype=torch.int32, device=device)
    TR.pack_edge_records(v)
    opt = TR.initialize_optimizer(params, v)
    for g in opt.param_groups:  # steady state: lrs frozen after t = 0 (train_utils.py:370-373)
        if g["name"] in ("logit_opacities", "log_scales", "cam_m", "cam_c", "rgb_colors"):
            g["lr"] = 0.0
    tgt = {k: t.to(device) for k, t in prob["target"].items()}
    ones = torch.ones_like(tgt["colors_precomp"])
    dataset, host = [], []
    for c in prob["cams"]:
        cam = TR.setup_camera(c["w"], c["h"], c["k"], c["w2c"], near=1.0, far=100, device=device)
        with torch.no_grad():
            out, _, _, _ = R.raster_forward(cam, tgt["means3D"], tgt["opacities"], tgt["colors_precomp"], tgt["scales"],
                                            tgt["rotations"], colors1=ones)
        im = out[:3].clone()
        seg = seg_target_from_mask(out[3]).contiguous()
        dataset.append({"cam": cam, "im": im, "seg": seg, "id": c["id"]})
        host.append((im.cpu().pin_memory(), seg.cpu().pin_memory()))
    return params, v, opt, dataset, host


def seg_target_from_mask(mask):
    """(seg, 0, 1-seg) colour coding of the reference's dataset loader (train_utils.py:71-75)."""
    return torch.stack((mask, torch.zeros_like(mask), 1 - mask))


# ----------------------------------------------------------------------------------------------------
# GNN workloads (configs 1 and 4 of SURVEY.md §8d): model configs of src/config/{rope,sloth}.yaml, seeded weights and
# rollout-style graph inputs (dynamics_module.py:106-125)
# ----------------------------------------------------------------------------------------------------
def model_dims(cfg):
    motion = cfg.get('motion_dim', 0)
    in_dim = cfg['n_his'] * cfg['state_dim'] + (cfg['n_his'] - 1) * motion + cfg['attr_dim'] + cfg['action_dim']
    rel_dim = cfg['rel_attr_dim'] * 2 + cfg['rel_group_dim'] + cfg['rel_distance_dim'] * cfg['n_his']
    return in_dim, rel_dim


def make_state_dict(cfg, seed, head_scale=1.0):
    """Deterministic weights from numpy's PCG64 stream (stable across versions): U(-1/sqrt(fan_in), 1/sqrt(fan_in)).
    head_scale multiplies the last layer of the motion head: an untrained head moves particles by ~0.1 m per step, which
    empties the neighbour graph within a few rollout steps; the rollout benchmark uses 1e-3 (mm-scale motion, like a
    trained model) so that all 50 steps run on a realistic 8-NN graph."""
    rng = np.random.default_rng(seed)
    in_dim, rel_dim = model_dims(cfg)
    nf = cfg['nf_effect']
    shapes = {}
    for name, d in (("particle_encoder", in_dim), ("relation_encoder", rel_dim)):
        shapes[f"{name}.model.0"] = (cfg['nf_particle'] if name[0] == 'p' else cfg['nf_relation'], d)
        h = shapes[f"{name}.model.0"][0]
        shapes[f"{name}.model.2"] = (h, h)
        shapes[f"{name}.model.4"] = (nf, h)
    shapes["particle_propagator.linear"] = (nf, 2 * nf)
    shapes["relation_propagator.linear"] = (nf, 3 * nf)
    shapes["non_rigid_predictor.linear_0"] = (nf, nf)
    shapes["non_rigid_predictor.linear_1"] = (nf, nf)
    shapes["non_rigid_predictor.linear_2"] = (3, nf)
    sd = {}
    for k, (o, i) in shapes.items():
        b = 1.0 / np.sqrt(i)
        sd[k + ".weight"] = torch.tensor(rng.uniform(-b, b, size=(o, i)), dtype=torch.float32)
        sd[k + ".bias"] = torch.tensor(rng.uniform(-b, b, size=(o,)), dtype=torch.float32)
    if head_scale != 1.0:
        sd["non_rigid_predictor.linear_2.weight"] *= head_scale
        sd["non_rigid_predictor.linear_2.bias"] *= head_scale
    return sd


def sloth_cfg(nf=512):
    return dict(verbose=False, nf_particle=nf, nf_relation=nf, nf_effect=nf, attr_dim=2, state_dim=1, motion_dim=3, action_dim=3,
                pstep=3, rel_attr_dim=2, rel_group_dim=1, rel_distance_dim=3, n_his=3)


def rope_cfg(nf=512):
    return dict(verbose=False, nf_particle=nf, nf_relation=nf, nf_effect=nf, attr_dim=2, state_dim=0, action_dim=3, pstep=3,
                rel_attr_dim=2, rel_group_dim=1, rel_distance_dim=3, n_his=3)


def make_graph_inputs(n_obj, seed, kind="sloth", n_his=3):
    """Seeded rollout-style inputs (dynamics_module.py:106-125): n_obj object particles + 1 tool particle (last)."""
    rng = np.random.default_rng(seed)
    if kind == "rope":
        x = np.arange(n_obj) * 0.009
        base = np.stack([x, np.zeros(n_obj), np.zeros(n_obj)], 1) + rng.normal(scale=0.002, size=(n_obj, 3))
    else:
        base = rng.uniform([0, 0, 0], [0.5, 0.5, 0.1], size=(n_obj, 3))
    N = n_obj + 1
    states = np.zeros((1, n_his, N, 3), np.float32)
    for h in range(n_his):
        states[0, h, :n_obj] = base + rng.normal(scale=0.001, size=(n_obj, 3)) * (n_his - 1 - h)
        states[0, h, n_obj] = np.array([0.25, 0.25, 0.12]) + 0.005 * h * np.array([1.0, 0, 0])
    action = np.zeros((1, N, 3), np.float32)
    action[0, n_obj] = [0.005, 0, 0]
    attrs = np.zeros((1, N, 2), np.float32)
    attrs[0, :n_obj, 0] = 1
    attrs[0, n_obj:, 1] = 1
    t = torch.tensor
    return dict(state=t(states), action=t(action), attrs=t(attrs), p_instance=torch.ones(1, n_obj, 1),
                state_mask=torch.ones(N, dtype=torch.bool), eef_mask=torch.tensor([False] * n_obj + [True]))


def make_training_batch(B, n_obj, seed, kind="sloth", n_future=3, n_his=3, n_pad=2, learnable=False):
    """Seeded training-style batch (the dict DynDataset yields, /root/reference/src/data/dataset.py:240-420, without graph
    matrices): B elements of n_obj object particles + n_pad padding particles (state_mask False, as the dataset pads to
    max_nobj) + 1 tool particle (last).  Futures: the objects drift by a smooth seeded field, the tool advances 5 mm per step.
    learnable=True makes the futures a function of the inputs (the objects follow 60 % of the tool's displacement, plus 0.1 mm of
    noise) instead of a per-element random drift: data-parallel runs on DIFFERENT batches then share one mapping to learn, and the
    averaged gradient lowers every rank's loss (the random drift can only be memorised batch by batch)."""
    rng = np.random.default_rng(seed)
    N = n_obj + n_pad + 1
    n_p = n_obj + n_pad
    out = {k: [] for k in ("state", "action", "attrs", "p_instance", "state_mask", "eef_mask", "state_future", "tool_future",
                           "action_future")}
    for b in range(B):
        gi = make_graph_inputs(n_obj, seed * 1000 + b, kind, n_his)
        st = torch.zeros(n_his, N, 3)
        st[:, :n_obj] = gi["state"][0, :, :n_obj]
        st[:, -1] = gi["state"][0, :, n_obj]
        act = torch.zeros(N, 3)
        act[-1] = gi["action"][0, n_obj]
        attrs = torch.zeros(N, 2)
        attrs[:n_obj, 0] = 1
        attrs[-1, 1] = 1
        p_inst = torch.zeros(n_p, 1)
        p_inst[:n_obj] = 1
        smask = torch.zeros(N, dtype=torch.bool)
        smask[:n_obj] = True
        smask[-1] = True
        emask = torch.zeros(N, dtype=torch.bool)
        emask[-1] = True
        drift = torch.tensor(rng.normal(scale=0.002, size=(n_future, 1, 3)), dtype=torch.float32).cumsum(0)
        noise = torch.tensor(rng.normal(scale=0.0005, size=(n_future, n_obj, 3)), dtype=torch.float32)
        if learnable:
            drift = 0.6 * act[-1][None, None, :] * torch.arange(1, n_future + 1, dtype=torch.float32)[:, None, None]
            noise = 0.2 * noise
        sf = torch.zeros(n_future, n_p, 3)
        sf[:, :n_obj] = st[-1, :n_obj][None] + drift + noise
        tf = torch.zeros(max(n_future - 1, 1), N, 3)
        af = torch.zeros(max(n_future - 1, 1), N, 3)
        for f in range(n_future - 1):
            tf[f, -1] = st[-1, -1] + act[-1] * (f + 1)
            af[f, -1] = act[-1]
        for k, v in (("state", st), ("action", act), ("attrs", attrs), ("p_instance", p_inst), ("state_mask", smask),
                     ("eef_mask", emask), ("state_future", sf), ("tool_future", tf), ("action_future", af)):
            out[k].append(v)
    return {k: torch.stack(v) for k, v in out.items()}
