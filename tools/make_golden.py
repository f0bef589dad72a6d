"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE (/root/reference/src) in the build container.
The GPU box has no /root/reference: tests read only the committed fixtures.  Run:  python tools/make_golden.py

tracking_golden.npz — outputs of the reference's own helpers (tracking/helpers.py: quat_mult, l1_loss_v1/v2,
weighted_l2_loss_v1/v2; tracking/external.py: calc_ssim, build_rotation) composed exactly as get_loss does at
train_utils.py:182-228, on seeded inputs built by oracle.tracking_oracle.make_prior_case.
"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
import numpy as np, torch

# ---- stubs for modules the reference imports at module top but that this path never calls
for name in ("open3d", "diff_gaussian_rasterization", "dgl", "dgl.geometry"):
    m = types.ModuleType(name)
    sys.modules[name] = m
sys.modules["diff_gaussian_rasterization"].GaussianRasterizationSettings = object
sys.modules["diff_gaussian_rasterization"].GaussianRasterizer = object
sys.modules["dgl.geometry"].farthest_point_sampler = None
sys.path.insert(0, os.path.join(REF, "src", "tracking"))
import helpers as ref_h      # /root/reference/src/tracking/helpers.py
import external as ref_e     # /root/reference/src/tracking/external.py

# build_rotation hard-codes device='cuda' (external.py:28); run it on the CPU by dropping that kwarg
_zeros = torch.zeros
def _zeros_cpu(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)

def ref_build_rotation(q):
    torch.zeros = _zeros_cpu
    try:
        return ref_e.build_rotation(q)
    finally:
        torch.zeros = _zeros

from oracle import tracking_oracle as T

out = {}
# ---- photometric
g = torch.Generator().manual_seed(0)
x = torch.rand(3, 50, 70, generator=g).requires_grad_(True)
y = torch.rand(3, 50, 70, generator=g)
loss = 0.8 * ref_h.l1_loss_v1(x, y) + 0.2 * (1.0 - ref_e.calc_ssim(x, y))
loss.backward()
out.update(ph_x=x.detach().numpy(), ph_y=y.numpy(), ph_loss=loss.item(), ph_l1=ref_h.l1_loss_v1(x, y).item(),
           ph_ssim=ref_e.calc_ssim(x, y).item(), ph_grad=x.grad.numpy())

# ---- priors, composed as train_utils.py:200-228
for tag, (G, K, seed, fb) in dict(a=(96, 6, 1, 0.25), b=(64, 4, 2, 0.0)).items():
    c = T.make_prior_case(G, K, seed, fb)
    m3 = c["means3D"].clone().requires_grad_(True)
    rq = c["rotations"].clone().requires_grad_(True)
    is_fg = c["is_fg"]
    fg_pts, fg_rot = m3[is_fg], rq[is_fg]
    rel_rot = ref_h.quat_mult(fg_rot, c["prev_inv_rot_fg"])
    rot = ref_build_rotation(rel_rot)
    neighbor_pts = fg_pts[c["neighbor_indices"]]
    curr_offset = neighbor_pts - fg_pts[:, None]
    curr_offset_in_prev_coord = (rot.transpose(2, 1)[:, None] @ curr_offset[:, :, :, None]).squeeze(-1)
    L = {}
    L["rigid"] = ref_h.weighted_l2_loss_v2(curr_offset_in_prev_coord, c["prev_offset"], c["neighbor_weight"])
    L["rot"] = ref_h.weighted_l2_loss_v2(rel_rot[c["neighbor_indices"]], rel_rot[:, None], c["neighbor_weight"])
    curr_offset_mag = torch.sqrt((curr_offset ** 2).sum(-1) + 1e-20)
    L["iso"] = ref_h.weighted_l2_loss_v1(curr_offset_mag, c["neighbor_dist"], c["neighbor_weight"])
    L["floor"] = torch.clamp(fg_pts[:, 1], min=0).mean()
    if (~is_fg).any():
        L["bg"] = ref_h.l1_loss_v2(m3[~is_fg], c["init_bg_pts"]) + ref_h.l1_loss_v2(rq[~is_fg], c["init_bg_rot"])
    else:
        L["bg"] = m3.sum() * 0.0
    w = dict(rigid=200.0, rot=4.0, iso=1000.0, floor=2.0, bg=200.0)   # train_gs.py defaults + train_utils.py:237
    total = sum(w[k] * v for k, v in L.items())
    total.backward()
    out.update({f"pr_{tag}_{k}": v.item() for k, v in L.items()})
    out.update({f"pr_{tag}_total": total.item(), f"pr_{tag}_gx": m3.grad.numpy(), f"pr_{tag}_gq": rq.grad.numpy(),
                f"pr_{tag}_cfg": np.array([G, K, seed, fb])})
dst = os.path.join(ROOT, "tests", "golden", "tracking_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes")

# =====================================================================================================================
# gnn_golden.npz — the reference's DynamicsPredictor.forward (src/gnn/model.py) and construct_edges_from_states
# (src/data/dataset.py) on seeded inputs (oracle.gnn_oracle.make_graph_inputs) with seeded weights
# (oracle.gnn_oracle.make_state_dict, numpy PCG64), nf = 128 so that the fixture stays small.
# =====================================================================================================================
sys.path.insert(0, os.path.join(REF, "src"))
from gnn.model import DynamicsPredictor as RefPredictor          # /root/reference/src/gnn/model.py:70
from data.dataset import construct_edges_from_states as ref_edges  # /root/reference/src/data/dataset.py:88
from oracle import gnn_oracle as GO

gout = {}
for tag, (kind, n_obj, topk, adj, conn, seed) in dict(sloth=("sloth", 300, 6, 0.075, True, 11), rope=("rope", 200, 5, 0.08, False, 12)).items():
    cfg = GO.sloth_cfg(128) if kind == "sloth" else GO.rope_cfg(128)
    model = RefPredictor(dict(cfg), torch.device("cpu"))
    model.load_state_dict(GO.make_state_dict(cfg, seed))
    model.eval()
    gi = GO.make_graph_inputs(n_obj, seed, kind)
    Rr, Rs = ref_edges(gi["state"][0, -1], adj, mask=gi["state_mask"], tool_mask=gi["eef_mask"], topk=topk, connect_all=conn)
    with torch.no_grad():
        pos, mot = model(state=gi["state"], attrs=gi["attrs"], Rr=Rr[None], Rs=Rs[None], p_instance=gi["p_instance"], action=gi["action"])
    gout.update({f"{tag}_recv": Rr.argmax(-1).numpy().astype(np.int32), f"{tag}_send": Rs.argmax(-1).numpy().astype(np.int32),
                 f"{tag}_pred_pos": pos.numpy(), f"{tag}_pred_motion": mot.numpy(),
                 f"{tag}_cfg": np.array([n_obj, topk, adj, float(conn), seed])})
dst = os.path.join(ROOT, "tests", "golden", "gnn_golden.npz")
np.savez_compressed(dst, **gout)
print("wrote", dst, os.path.getsize(dst), "bytes")

# ------------------------------------------------------------------------------------------------------------------
# skinning_golden.npz — the reference's render/utils.py (interpolate_motions, mat2quat, quat2mat, relations_to_matrix)
# run on the CPU (its `device` argument) on the seeded scenes of oracle.skinning_oracle.make_skinning_inputs, which
# include an isolated bone (rank 0), a two-bone pair (rank 1) and a coplanar neighbourhood (rank 2).
# ------------------------------------------------------------------------------------------------------------------
import warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.join(REF, "src"))
from render.utils import interpolate_motions as ref_interp, mat2quat as ref_m2q, quat2mat as ref_q2m, relations_to_matrix as ref_r2m
from oracle import skinning_oracle as SO

sout = {}
for tag, (nb, npart, seed) in {"a": (40, 300, 0), "b": (150, 1000, 1)}.items():
    d = SO.make_skinning_inputs(nb, npart, seed)
    x, q, w = ref_interp(d["bones"], d["motions"], d["relations"], d["xyz"], quat=d["quat"], device="cpu")
    for k, v in d.items():
        sout[f"{tag}_{k}"] = v.numpy()
    sout[f"{tag}_xyz_out"], sout[f"{tag}_quat_out"], sout[f"{tag}_weights"] = x.numpy(), q.numpy(), w.numpy()
    # caller-supplied weights variant (utils.py:207 skipped)
    g = torch.Generator().manual_seed(seed)
    wg = torch.rand(npart, nb, generator=g)
    wg = wg / wg.sum(1, keepdim=True)
    x2, q2, _ = ref_interp(d["bones"], d["motions"], d["relations"], d["xyz"], quat=d["quat"], weights=wg, device="cpu")
    sout[f"{tag}_weights_given"], sout[f"{tag}_xyz_out_given"], sout[f"{tag}_quat_out_given"] = wg.numpy(), x2.numpy(), q2.numpy()
g = torch.Generator().manual_seed(5)
qq = torch.nn.functional.normalize(torch.randn(64, 4, generator=g), dim=-1)
qq[0] = torch.tensor([0., 1., 0., 0.]); qq[1] = torch.tensor([0., 0., 1., 0.]); qq[2] = torch.tensor([0., 0., 0., 1.])  # trace = -1 branches
Rm = ref_q2m(qq)
sout["m2q_quat_in"], sout["m2q_rot"], sout["m2q_quat_out"] = qq.numpy(), Rm.numpy(), ref_m2q(Rm).numpy()
# relations_to_matrix on a small one-hot pair
recv = torch.tensor([0, 0, 1, 2, 2, 3]); send = torch.tensor([0, 1, 1, 2, 0, 3])
Rr = torch.zeros(1, 6, 4); Rs = torch.zeros(1, 6, 4)
Rr[0, torch.arange(6), recv] = 1; Rs[0, torch.arange(6), send] = 1
sout["r2m_Rr"], sout["r2m_Rs"], sout["r2m_rel"] = Rr.numpy(), Rs.numpy(), ref_r2m(Rr, Rs).numpy()
dst = os.path.join(ROOT, "tests", "golden", "skinning_golden.npz")
np.savez_compressed(dst, **sout)
print("wrote", dst, os.path.getsize(dst), "bytes")

# ------------------------------------------------------------------------------------------------------------------
# formats_golden.npz — on-disk formats: the reference's .splat writer (real_world/gs/convert.py), save_params / params2cpu
# (tracking/helpers.py:122-140) and the path mapping of train_utils.py:10-29, run as they are on seeded inputs.
# ------------------------------------------------------------------------------------------------------------------
import importlib.util, tempfile
spec = importlib.util.spec_from_file_location("ref_convert", os.path.join(REF, "src", "real_world", "gs", "convert.py"))
ref_convert = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_convert)
rng = np.random.default_rng(7)
n = 64
fin = dict(pts=rng.normal(size=(n, 3)), colors=rng.uniform(-0.1, 1.1, size=(n, 3)), scales=rng.uniform(0.001, 0.02, size=(n, 3)),
           quats=rng.normal(size=(n, 4)), opacities=rng.uniform(0, 1, size=(n, 1)))
fout = {f"splat_{k}": v for k, v in fin.items()}
with tempfile.TemporaryDirectory() as td:
    ref_convert.save_to_splat(fin["pts"], fin["colors"], fin["scales"], fin["quats"], fin["opacities"], os.path.join(td, "a.splat"))
    fout["splat_bytes"] = np.frombuffer(open(os.path.join(td, "a.splat"), "rb").read(), dtype=np.uint8)
    # save_params: frame 0 holds every key, later frames only means3D / rgb_colors / unnorm_rotations
    frames = []
    for t in range(3):
        p = {k: torch.tensor(rng.normal(size=s)).float() for k, s in dict(means3D=(n, 3), rgb_colors=(n, 3), seg_colors=(n, 3), unnorm_rotations=(n, 4),
                                                                           logit_opacities=(n, 1), log_scales=(n, 3), cam_m=(50, 3), cam_c=(50, 3)).items()}
        for k, v in p.items():
            fout[f"params_in_{t}_{k}"] = v.numpy()
        frames.append(ref_h.params2cpu(p, t == 0))
    cwd = os.getcwd(); os.chdir(td)
    try:
        ref_h.save_params(frames, "seqA", "expA")
        saved = dict(np.load(os.path.join(td, "output", "expA", "seqA", "params.npz")))
    finally:
        os.chdir(cwd)
    for k, v in saved.items():
        fout[f"params_saved_{k}"] = v
sys.path.insert(0, os.path.join(REF, "src", "tracking"))
import train_utils as ref_tu   # /root/reference/src/tracking/train_utils.py (stubs above satisfy its imports)
paths = ["camera_1/foreground/foreground_000012.png", "camera_3/color/color_7.png", "a/b/camera_0/foreground/foreground_000100.png"]
fout["paths_in"] = np.array(paths)
fout["paths_seg"] = np.array([ref_tu.map_to_segmentation_path(p) for p in paths])
fout["paths_depth"] = np.array([ref_tu.map_to_depth_path(p) for p in ["camera_1/color_12.png", "x/y/im_000003.png"]])
import random as _random
_random.seed(0)
_ds, _todo = [{"id": i} for i in range(4)], []
fout["get_batch_ids"] = np.array([ref_tu.get_batch(_todo, _ds)["id"] for _ in range(40)])   # train_utils.py:81-85 as it behaves
assert _todo == []
dst = os.path.join(ROOT, "tests", "golden", "formats_golden.npz")
np.savez_compressed(dst, **fout)
print("wrote", dst, os.path.getsize(dst), "bytes")

# ------------------------------------------------------------------------------------------------------------------
# gnn_train_golden.npz — the reference's training unroll (src/train.py:183-211) composed from ITS model and ITS loss
# functions (train.mse_loss, train.length_loss) on a seeded batch (workloads.make_training_batch), nf = 128:
# loss, per-step loss terms, and gradients of selected parameters after loss_sum.backward().
# ------------------------------------------------------------------------------------------------------------------
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].use = lambda *a, **k: None
import train as ref_train                                          # /root/reference/src/train.py

GRAD_KEYS = ["particle_encoder.model.0.weight", "relation_encoder.model.0.weight", "relation_encoder.model.4.bias",
             "relation_propagator.linear.bias", "particle_propagator.linear.bias", "non_rigid_predictor.linear_2.weight"]
tout = {}
for tag, (kind, B, n_obj, topk, adj, conn, seed, n_future) in dict(sloth=("sloth", 2, 60, 5, 0.075, True, 21, 3),
                                                                    rope=("rope", 3, 40, 4, 0.08, False, 22, 2)).items():
    cfg = GO.sloth_cfg(128) if kind == "sloth" else GO.rope_cfg(128)
    model = RefPredictor(dict(cfg), torch.device("cpu"))
    model.load_state_dict(GO.make_state_dict(cfg, seed, head_scale=0.05))
    model.train()
    batch = GO.make_training_batch(B, n_obj, seed, kind, n_future)
    Rr, Rs = GO.batch_edges(batch, adj, topk, conn)
    # cross-check the padded one-hots against the reference's own edge builder
    for b in range(B):
        rr, rs_ = ref_edges(batch["state"][b, -1], adj, mask=batch["state_mask"][b], tool_mask=batch["eef_mask"][b], topk=topk, connect_all=conn)
        assert torch.equal(rr, Rr[b, :rr.shape[0]]) and torch.equal(rs_, Rs[b, :rs_.shape[0]]) and Rr[b, rr.shape[0]:].abs().sum() == 0
    data = dict(batch)
    data["Rr"], data["Rs"] = Rr, Rs
    loss_funcs = [(ref_train.mse_loss, 1.0), (ref_train.length_loss, 0.01)]
    loss_sum, items = 0, []
    future_state, future_tool, future_action = data['state_future'], data['tool_future'], data['action_future']
    for fi in range(n_future):                                      # train.py:183-211, line for line
        gt_state = future_state[:, fi].clone()
        pred_state, pred_motion = model(**data)
        pred_state_p = pred_state[:, :gt_state.shape[1], :3].clone()
        loss = [weight * func(pred_state_p, gt_state, **data) for func, weight in loss_funcs]
        loss_sum += sum(loss)
        items.append([l.item() for l in loss])
        if fi < n_future - 1:
            next_tool = future_tool[:, fi].clone()
            next_action = future_action[:, fi].clone()
            next_state = next_tool.unsqueeze(1)
            next_state[:, -1, :pred_state_p.shape[1]] = pred_state_p
            next_state = torch.cat([data['state'][:, 1:], next_state], dim=1)
            data["state"] = next_state
            data["action"] = next_action
    loss_sum.backward()
    grads = dict(model.named_parameters())
    tout.update({f"{tag}_cfg": np.array([B, n_obj, topk, adj, float(conn), seed, n_future]), f"{tag}_loss": loss_sum.item(),
                 f"{tag}_items": np.array(items), f"{tag}_pred_last": pred_state_p.detach().numpy(),
                 f"{tag}_grad_abs_sum": np.array([float(p.grad.abs().sum()) for _, p in model.named_parameters()])})
    for k in GRAD_KEYS:
        tout[f"{tag}_grad_{k}"] = grads[k].grad.numpy()
dst = os.path.join(ROOT, "tests", "golden", "gnn_train_golden.npz")
np.savez_compressed(dst, **tout)
print("wrote", dst, os.path.getsize(dst), "bytes")

# rigid_loss / umeyama (train.py:30-38, gnn/utils.py:7-40) on a seeded case: value, gradient, R, t
g = torch.Generator().manual_seed(5)
Bq, Nq = 3, 25
X = torch.rand(Bq, Nq, 3, generator=g)
ang = torch.tensor([0.3, -0.2, 0.5])
Rz = torch.stack([torch.stack([torch.cos(ang), -torch.sin(ang), torch.zeros(3)], 1), torch.stack([torch.sin(ang), torch.cos(ang), torch.zeros(3)], 1),
                  torch.tensor([[0., 0., 1.]]).repeat(3, 1)], 1)
Y = (X @ Rz.transpose(1, 2) + torch.tensor([0.1, -0.05, 0.02]) + 0.01 * torch.randn(Bq, Nq, 3, generator=g)).requires_grad_(True)
qmask = torch.rand(Bq, Nq, generator=g) > 0.2
stq = X[:, None].repeat(1, 3, 1, 1)
lq = ref_train.rigid_loss(Y, None, state=stq, obj_mask=qmask)
lq.backward()
from gnn.utils import umeyama_algorithm as ref_umeyama          # /root/reference/src/gnn/utils.py:7
cq, Rq, tq = ref_umeyama(X, Y.detach(), qmask.float(), fixed_scale=False)
rout = dict(rig_X=X.numpy(), rig_Y=Y.detach().numpy(), rig_mask=qmask.numpy(), rig_loss=lq.item(), rig_grad=Y.grad.numpy(),
            ume_c=cq.numpy(), ume_R=Rq.numpy(), ume_t=tq.numpy())
dst = os.path.join(ROOT, "tests", "golden", "gnn_rigid_golden.npz")
np.savez_compressed(dst, **rout)
print("wrote", dst, os.path.getsize(dst), "bytes")

# ------------------------------------------------------------------------------------------------------------------
# densify_golden.npz — the reference's densify / cat_params_to_optimizer / remove_points / update_params_and_optimizer
# (tracking/external.py:138-299) and initialize_optimizer (tracking/train_utils.py:152-164) run AS THEY ARE on the CPU (their
# hard-coded device="cuda" dropped, torch.normal replaced by fixture noise so that the split samples are reproducible on any
# device) through a clone / split / prune round (i = 500), the round with the big-point prune and the opacity reset
# (i = 3000) and the remove_thresh_5k round (i = 5000).  Inputs of every round are seeded by (round, current point count).
# ------------------------------------------------------------------------------------------------------------------
def densify_round_inputs(rnd, n):
    """(means2D.grad [n,3], seen [n]) of one round: shared by this generator and tests/test_densify_gpu.py."""
    r = np.random.default_rng(1000 + rnd)
    g = r.normal(scale=3e-4, size=(n, 3)).astype(np.float32)
    seen = r.uniform(size=n) < 0.8
    return g, seen


def densify_initial_state(n=400, seed=5):
    r = np.random.default_rng(seed)
    p = dict(means3D=r.uniform(-0.2, 0.2, size=(n, 3)), rgb_colors=r.uniform(size=(n, 3)),
             seg_colors=np.stack([np.ones(n), np.zeros(n), np.zeros(n)], -1), unnorm_rotations=r.normal(size=(n, 4)),
             logit_opacities=r.normal(scale=3.0, size=(n, 1)), log_scales=np.log(r.uniform(0.01, 0.14, size=(n, 3))),
             cam_m=np.zeros((50, 3)), cam_c=np.zeros((50, 3)))
    v = dict(max_2D_radius=r.uniform(0, 30, size=n), means2D_gradient_accum=r.uniform(0, 0.1, size=n) * (r.uniform(size=n) < 0.9),
             denom=np.floor(r.uniform(0, 600, size=n)) * (r.uniform(size=n) < 0.95))
    moments = {k: (r.normal(scale=1e-3, size=a.shape), r.uniform(0, 1e-6, size=a.shape)) for k, a in p.items()}
    noise = r.normal(size=(40 * n, 3))
    return ({k: a.astype(np.float32) for k, a in p.items()}, {k: a.astype(np.float32) for k, a in v.items()},
            {k: (m.astype(np.float32), s.astype(np.float32)) for k, (m, s) in moments.items()}, noise.astype(np.float32))


dp, dv, dm, dnoise = densify_initial_state()
params = {k: torch.nn.Parameter(torch.tensor(a).contiguous().requires_grad_(True)) for k, a in dp.items()}
params['rgb_colors'].requires_grad = False                      # train_utils.py:132
variables = {k: torch.tensor(a) for k, a in dv.items()}
variables['scene_radius'] = 1.0
optimizer = ref_tu.initialize_optimizer(params, variables)
for k, p_ in params.items():                                     # Adam state as after some iterations
    if p_.requires_grad:
        optimizer.state[p_] = dict(step=torch.tensor(7.0), exp_avg=torch.tensor(dm[k][0]), exp_avg_sq=torch.tensor(dm[k][1]))
dout = {"noise": dnoise, "scene_radius": np.float32(1.0)}
_normal, _zeros_real = torch.normal, torch.zeros
state = {"used": 0}
def _fixture_normal(mean=None, std=None, **kw):
    out = std * torch.tensor(dnoise[state["used"]:state["used"] + std.shape[0]])
    state["used"] += std.shape[0]
    return out
torch.normal, torch.zeros = _fixture_normal, _zeros_cpu
try:
    for rnd, it in enumerate((499, 500, 3000, 5000, 5001)):
        n_now = params['means3D'].shape[0]
        g2d, seen = densify_round_inputs(rnd, n_now)
        if it == 5000:   # between the opacity reset (i = 3000) and this round the optimiser has moved the opacities
            params['logit_opacities'].data.copy_(torch.tensor(np.random.default_rng(77).normal(scale=2.0, size=(n_now, 1)).astype(np.float32)))
        m2d = torch.zeros(n_now, 3, requires_grad=True)
        m2d.grad = torch.tensor(g2d)
        variables['means2D'], variables['seen'] = m2d, torch.tensor(seen)
        with torch.no_grad():
            params, variables, num_pts = ref_e.densify(params, variables, optimizer, it, 0.005, 0.25, 0.05)
        dout[f"r{rnd}_iter"], dout[f"r{rnd}_n_in"], dout[f"r{rnd}_n_out"], dout[f"r{rnd}_noise_used"] = it, n_now, num_pts, state["used"]
        for k, p_ in params.items():
            dout[f"r{rnd}_p_{k}"] = p_.detach().numpy().copy()
            st = optimizer.state.get(p_, None)
            if st is not None and "exp_avg" in st:
                dout[f"r{rnd}_m_{k}"], dout[f"r{rnd}_v_{k}"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
        for k in ("means2D_gradient_accum", "denom", "max_2D_radius"):
            dout[f"r{rnd}_var_{k}"] = variables[k].numpy().copy()
finally:
    torch.normal, torch.zeros = _normal, _zeros_real
assert dout["r1_n_out"] != dout["r1_n_in"] and dout["r2_n_out"] != dout["r2_n_in"]
dst = os.path.join(ROOT, "tests", "golden", "densify_golden.npz")
np.savez_compressed(dst, **dout)
print("wrote", dst, os.path.getsize(dst), "bytes", {k: int(dout[k]) for k in dout if k.endswith("_n_out")})
