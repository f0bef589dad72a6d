"""GPU parity tests of the Gaussian skinning step (gsd_skin_bone_transforms / gsd_skin_apply through the C ABI) against the
float64 oracle and the reference's committed outputs.

Tolerances: bone rotations |Δ| <= 2e-6 (fp64 SVD of the same fp32 covariance); positions <= 2e-6 m abs (scene 0.3 m; 40-2000
term fp32 sums); blended quaternions <= 2e-5; dense weights <= 1e-6 abs vs fp64 (the REFERENCE's own fp32 weights are 1e-4 off
because its cdist expands |x-b|^2 through a matmul)."""
import os

import numpy as np
import pytest
import torch

from oracle import skinning_oracle as SO

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "skinning_golden.npz"))
C = lambda a: torch.tensor(a).cuda()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_golden_reference_outputs(tag):
    from gs_dynamics_b200 import skinning as SK
    d = {k: C(GOLD[f"{tag}_{k}"]) for k in ("bones", "motions", "relations", "xyz", "quat")}
    x, q, w = SK.interpolate_motions(d["bones"], d["motions"], d["relations"], d["xyz"], quat=d["quat"])
    assert np.abs(x.cpu().numpy() - GOLD[f"{tag}_xyz_out"]).max() < 3e-6
    assert np.abs(q.cpu().numpy() - GOLD[f"{tag}_quat_out"]).max() < 3e-5
    assert np.abs(w.cpu().numpy() - GOLD[f"{tag}_weights"]).max() < 2e-4
    x2, q2, w2 = SK.interpolate_motions(d["bones"], d["motions"], d["relations"], d["xyz"], quat=d["quat"], weights=C(GOLD[f"{tag}_weights_given"]))
    assert np.abs(x2.cpu().numpy() - GOLD[f"{tag}_xyz_out_given"]).max() < 3e-6
    assert np.abs(q2.cpu().numpy() - GOLD[f"{tag}_quat_out_given"]).max() < 3e-5


@pytest.mark.parametrize("n_bones,n_particles,seed", [(40, 300, 3), (150, 5000, 4), (600, 8000, 5), (7, 1, 6)])
def test_vs_float64_oracle(n_bones, n_particles, seed):
    from gs_dynamics_b200 import skinning as SK
    d = SO.make_skinning_inputs(n_bones, n_particles, seed)
    ox, oq, ow, oR = SO.interpolate_motions(d["bones"], d["motions"], d["relations"], d["xyz"], quat=d["quat"], dtype=torch.float64)
    g = {k: v.cuda() for k, v in d.items()}
    tf, R = SK.bone_transforms(g["bones"], g["motions"], g["relations"])
    assert np.abs(R.cpu().numpy() - oR.numpy()).max() < 2e-6
    x, q, w = SK.interpolate_motions(g["bones"], g["motions"], g["relations"], g["xyz"], quat=g["quat"])
    assert np.abs(x.cpu().numpy() - ox.numpy()).max() < 2e-6
    assert np.abs(q.cpu().numpy() - oq.numpy()).max() < 2e-5
    assert np.abs(w.cpu().numpy() - ow.numpy()).max() < 1e-6
    # rot passthrough when no quaternion is given; weights skipped on request
    rot_in = torch.eye(3, device="cuda").repeat(n_particles, 1, 1)
    x3, r3, w3 = SK.interpolate_motions(g["bones"], g["motions"], g["relations"], g["xyz"], rot=rot_in, return_weights=False)
    assert r3 is rot_in and w3 is None and torch.equal(x3, x)


def test_edge_index_relations_drop_the_tool_node():
    """relations given as the rollout's EdgeIndex (tool node last, connect_all edges) == dense relations[:nobj, :nobj]."""
    from gs_dynamics_b200 import skinning as SK, gnn
    d = SO.make_skinning_inputs(200, 1000, 9, special=False)
    bones = d["bones"].cuda()
    states = torch.cat([bones, torch.tensor([[0.15, 0.15, 0.2]]).cuda()], 0)
    mask = torch.ones(201, dtype=torch.bool).cuda()
    tool = torch.zeros(201, dtype=torch.bool).cuda(); tool[-1] = True
    e = gnn.construct_edges_index(states, 0.08, mask, tool, topk=8, connect_all=True)
    Rr, Rs = e.dense()
    rel = SK.relations_to_matrix(Rr, Rs)[:200, :200]
    a = SK.interpolate_motions(bones, d["motions"].cuda(), e, d["xyz"].cuda(), quat=d["quat"].cuda(), return_weights=False)
    b = SK.interpolate_motions(bones, d["motions"].cuda(), rel, d["xyz"].cuda(), quat=d["quat"].cuda(), return_weights=False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    ox, oq, _, _ = SO.interpolate_motions(d["bones"], d["motions"], rel.cpu(), d["xyz"], quat=d["quat"])
    assert np.abs(a[0].cpu().numpy() - ox.numpy()).max() < 2e-6 and np.abs(a[1].cpu().numpy() - oq.numpy()).max() < 2e-5


def test_rigid_motion_is_reproduced_exactly():
    """size-independent property at full size (100k Gaussians, 2000 bones): if every bone moves by one rigid transform, every
    Gaussian moves by that transform and its rotation is composed with it."""
    from gs_dynamics_b200 import skinning as SK
    g = torch.Generator().manual_seed(0)
    bones = torch.rand(2000, 3, generator=g) * torch.tensor([0.5, 0.5, 0.1])
    xyz = torch.rand(100000, 3, generator=g) * torch.tensor([0.5, 0.5, 0.1])
    quat = torch.nn.functional.normalize(torch.randn(100000, 4, generator=g), dim=-1)
    qg = torch.nn.functional.normalize(torch.tensor([[0.9, 0.1, -0.3, 0.2]]), dim=-1)
    Rg = SO.quat2mat(qg)[0]
    t = torch.tensor([0.01, -0.02, 0.005])
    motions = bones @ Rg.T + t - bones
    d = torch.cdist(bones, bones)
    rel = (d < 0.05).long()
    x, q, _ = SK.interpolate_motions(bones.cuda(), motions.cuda(), rel.cuda(), xyz.cuda(), quat=quat.cuda(), return_weights=False)
    assert float((x.cpu() - (xyz @ Rg.T + t)).abs().max()) < 2e-6
    qs = qg.expand(100000, 4)
    want = torch.stack([qs[:, 0] * quat[:, 0] - qs[:, 1] * quat[:, 1] - qs[:, 2] * quat[:, 2] - qs[:, 3] * quat[:, 3],
                        qs[:, 0] * quat[:, 1] + qs[:, 1] * quat[:, 0] + qs[:, 2] * quat[:, 3] - qs[:, 3] * quat[:, 2],
                        qs[:, 0] * quat[:, 2] - qs[:, 1] * quat[:, 3] + qs[:, 2] * quat[:, 0] + qs[:, 3] * quat[:, 1],
                        qs[:, 0] * quat[:, 3] + qs[:, 1] * quat[:, 2] - qs[:, 2] * quat[:, 1] + qs[:, 3] * quat[:, 0]], -1)
    assert float((q.cpu() - want).abs().max()) < 1e-5


def test_mat2quat_quat2mat_relations_helpers():
    from gs_dynamics_b200 import skinning as SK
    assert np.abs(SK.quat2mat(C(GOLD["m2q_quat_in"])).cpu().numpy() - GOLD["m2q_rot"]).max() < 1e-6
    assert np.abs(SK.mat2quat(C(GOLD["m2q_rot"])).cpu().numpy() - GOLD["m2q_quat_out"]).max() < 1e-6
    assert np.array_equal(SK.relations_to_matrix(C(GOLD["r2m_Rr"]), C(GOLD["r2m_Rs"])).cpu().numpy(), GOLD["r2m_rel"])
    with pytest.raises(ValueError):
        SK.interpolate_motions(torch.zeros(3, 3), torch.zeros(3, 3), torch.eye(3), torch.zeros(5, 3).cuda())


def _dm_config(nf):
    cfg = SO_GO.sloth_cfg(nf)
    return dict(train_config=dict(n_his=3, dist_thresh=0.002, out_dir=""), model_config=cfg,
                dataset_config=dict(datasets=[dict(max_nobj=60, adj_radius_range=[0.07, 0.09], fps_radius_range=[0.04, 0.05],
                                                    topk=6, connect_all=True)]))


from oracle import gnn_oracle as SO_GO  # noqa: E402


def test_dynamics_module_rollout_matches_oracle():
    """DynamicsModule.rollout (FPS -> edges -> GNN -> skinning, variable particle count per step, one skipped step) against
    the CPU restatement of dynamics_module.py:53-172; 5 steps, 4000 Gaussians.  Tolerance 2e-5 m on positions, 1e-4 on
    quaternions (five autoregressive fp32 steps)."""
    from gs_dynamics_b200 import gnn
    from gs_dynamics_b200.dynamics_module import DynamicsModule
    config = _dm_config(128)
    cfg = config["model_config"]
    sd = SO_GO.make_state_dict(cfg, 3, head_scale=1e-2)
    model = gnn.DynamicsPredictor(dict(cfg), torch.device("cuda")).cuda().eval()
    model.load_state_dict(sd)
    dm = DynamicsModule(config, model=model)
    starts = iter(range(1000))
    dm.start_idx_fn = lambda n: 0
    g = torch.Generator().manual_seed(11)
    n = 4000
    xyz_0 = torch.rand(n, 3, generator=g) * torch.tensor([0.3, 0.3, 0.08])
    quat_0 = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    rgb_0, opa_0 = torch.rand(n, 3, generator=g), torch.rand(n, 1, generator=g)
    inlier = torch.arange(0, n, 2)
    n_steps = 6
    eef = torch.tensor([0.15, 0.15, 0.1]) + torch.arange(n_steps)[:, None].float() * torch.tensor([0.004, 0.0, -0.001])
    eef[3] = eef[2] + 1e-4                                   # below dist_thresh -> step 3 is skipped (copied)
    eef = eef[:, None, :]
    xyz, rgb, quat, opa, bones, eef_out = dm.rollout(xyz_0.cuda(), rgb_0.cuda(), quat_0.cuda(), opa_0.cuda(), eef.cuda(), n_steps, inlier)
    o_xyz, o_quat, o_bones, o_eef = SO.rollout(sd, cfg, dict(n_his=3, dist_thresh=0.002, max_nobj=60, adj_thresh=0.08, fps_radius=0.045,
                                                             topk=6, connect_all=True), xyz_0, quat_0, eef, n_steps, inlier, lambda m: 0)
    assert xyz.shape == (n_steps, n, 3) and rgb.shape == (n_steps, n, 3) and opa.shape == (n_steps, n, 1) and not xyz.is_cuda
    assert torch.equal(xyz[3], xyz[2]) and torch.equal(quat[3], quat[2])
    assert float((xyz[1] - xyz[0]).abs().max()) > 1e-4      # the Gaussians do move
    assert float((bones - o_bones).abs().max()) < 2e-5
    assert float((xyz - o_xyz).abs().max()) < 2e-5
    assert float((quat - o_quat).abs().max()) < 1e-4
    assert torch.equal(eef_out, o_eef)


def test_gnn_rollout_with_attached_gaussians_graph_matches_eager():
    from gs_dynamics_b200 import gnn
    cfg = SO_GO.sloth_cfg(128)
    model = gnn.DynamicsPredictor(dict(cfg), torch.device("cuda")).cuda().eval()
    model.load_state_dict(SO_GO.make_state_dict(cfg, 0, head_scale=1e-2))
    gi = SO_GO.make_graph_inputs(300, 2, "sloth")
    g = torch.Generator().manual_seed(1)
    xyz = (torch.rand(5000, 3, generator=g) * torch.tensor([0.5, 0.5, 0.1])).cuda()
    quat = torch.nn.functional.normalize(torch.randn(5000, 4, generator=g), dim=-1).cuda()
    outs = []
    for use_graph in (False, True):
        ro = gnn.GnnRollout(model, gi["state"][0, :, :300].cuda(), gi["state"][0, :, 300:].cuda(), 0.075, 8, True, use_graph=use_graph)
        ro.attach_gaussians(xyz, quat)
        for _ in range(4):
            ro.step(torch.tensor([0.005, 0.0, 0.0], device="cuda"))
        outs.append((ro.gs_xyz.clone(), ro.gs_quat.clone()))
    assert float((outs[0][0] - xyz).abs().max()) > 1e-4
    assert torch.allclose(outs[0][0], outs[1][0], atol=1e-6) and torch.allclose(outs[0][1], outs[1][1], atol=1e-6)
