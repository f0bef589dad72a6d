"""ORACLE — test infrastructure only (tests/, __graft_entry__.smoke(), bench.py's CPU legs).  Never imported by the product.

CPU restatement of the reference's Gaussian skinning step, /root/reference/src/render/utils.py:52-243
(`quat2mat`, `mat2quat`, `relations_to_matrix`, `interpolate_motions`), in any float dtype (float64 for tolerance budgets,
float32 to mirror the reference).  Pinned by tests/golden/skinning_golden.npz, which tools/make_golden.py generates by importing
and running the reference's own functions on the CPU (tests/test_oracle_skinning.py).
"""
import numpy as np
import torch


def quat2mat(q):  # utils.py:52-66
    q = q / torch.sqrt((q * q).sum(-1, keepdim=True))
    r, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)


def mat2quat(rot):  # utils.py:68-109 (one matrix at a time; plain loops — small inputs only)
    q = torch.zeros((rot.shape[0], 4), dtype=rot.dtype)
    for n in range(rot.shape[0]):
        r = rot[n]
        t = max(float(r[0, 0] + r[1, 1] + r[2, 2]), -1.0)
        if t > -1:
            s = (r[0, 0] + r[1, 1] + r[2, 2] + 1).sqrt()
            q[n, 0] = 0.5 * s
            s = 0.5 / s
            q[n, 1], q[n, 2], q[n, 3] = (r[2, 1] - r[1, 2]) * s, (r[0, 2] - r[2, 0]) * s, (r[1, 0] - r[0, 1]) * s
        elif r[0, 0] >= r[1, 1] and r[0, 0] >= r[2, 2]:
            s = 0.5 / (1 + r[0, 0] - r[1, 1] - r[2, 2]).sqrt()
            q[n, 0], q[n, 1], q[n, 2], q[n, 3] = (r[2, 1] - r[1, 2]) * s, 0.5 * s, (r[1, 0] + r[0, 1]) * s, (r[2, 0] + r[0, 2]) * s
        elif r[1, 1] >= r[2, 2] and r[1, 1] > r[0, 0]:
            s = 0.5 / (1 + r[1, 1] - r[0, 0] - r[2, 2]).sqrt()
            q[n, 0], q[n, 1], q[n, 2], q[n, 3] = (r[0, 2] - r[2, 0]) * s, (r[2, 1] + r[1, 2]) * s, 0.5 * s, (r[0, 1] + r[1, 0]) * s
        else:
            s = 0.5 / (1 + r[2, 2] - r[0, 0] - r[1, 1]).sqrt()
            q[n, 0], q[n, 1], q[n, 2], q[n, 3] = (r[1, 0] - r[0, 1]) * s, (r[0, 2] + r[2, 0]) * s, (r[1, 2] + r[2, 1]) * s, 0.5 * s
    return q


def relations_to_matrix(Rr, Rs):  # utils.py:128-134
    rel = torch.zeros((Rr.shape[-1], Rs.shape[-1]), dtype=torch.int64)
    for j in range(Rr.shape[1]):
        assert Rr[0, j].sum() == 1 and Rs[0, j].sum() == 1
        rel[int(Rr[0, j].argmax()), int(Rs[0, j].argmax())] = 1
    return rel


def bone_rotation(F32):
    """utils.py:168-202 for one fp32 covariance F (3x3 tensor).  The rank decision and the determinant sign are taken on the
    fp32 matrix exactly as the reference does (torch.linalg.matrix_rank / det on CPU LAPACK); the factorisation itself runs in
    float64 so that the oracle is the accurate answer the tolerances are budgeted against."""
    rank = int(torch.linalg.matrix_rank(F32))
    F = F32.double()
    eye = torch.eye(3, dtype=torch.float64)
    if rank == 1:
        U, S, Vh = torch.linalg.svd(F)
        axis = U[:, 0]
        if axis[0] > 0:      # LAPACK returns the first left singular vector of a rank-1 matrix with x component <= 0
            axis = -axis     # (checked against torch.svd on 2 000 random fp32 cases); the reference inherits that sign
        x = torch.tensor([1., 0., 0.], dtype=torch.float64)
        perp = torch.linalg.cross(axis, x)
        if torch.norm(perp) < 1e-6:
            return eye
        perp = perp / torch.norm(perp)
        third = torch.linalg.cross(x, perp)
        third_after = torch.linalg.cross(axis, perp)
        X = torch.stack([x, perp, third], dim=1)
        Y = torch.stack([axis, perp, third_after], dim=1)
        return Y @ X.T
    U, S, Vh = torch.linalg.svd(F)
    V = Vh.T
    Sg = torch.eye(3, dtype=torch.float64)
    if float(torch.linalg.det(F32)) < 0:
        if rank == 3:
            return eye       # `S[cov_rank, cov_rank] = -1` is out of range for rank 3 -> caught by the bare except -> identity
        Sg[rank, rank] = -1
    R = U @ Sg @ V.T
    if abs(float(torch.linalg.det(R)) - 1) > 1e-3 and rank < 3:
        Sg[rank, rank] *= -1
        R = U @ Sg @ V.T
    return R


def interpolate_motions(bones, motions, relations, xyz, quat=None, weights=None, dtype=torch.float64):
    """utils.py:137-243.  Returns (xyz_transformed, rot_or_None, weights, bone_rotations)."""
    bones32, motions32 = bones.float(), motions.float()
    n_bones = bones.shape[0]
    R = torch.zeros((n_bones, 3, 3), dtype=torch.float64)
    for i in range(n_bones):
        adj = relations[i].nonzero().squeeze(1)
        if len(adj) == 0:
            R[i] = torch.eye(3, dtype=torch.float64)
            continue
        a_old = bones32[adj] - bones32[i]
        a_new = (bones32[adj] + motions32[adj]) - (bones32[i] + motions32[i])
        F32 = a_new.T @ a_old     # fp32, like the reference (W = identity)
        R[i] = bone_rotation(F32)
    R = R.to(dtype)
    bones_d, motions_d, xyz_d = bones.to(dtype), motions.to(dtype), xyz.to(dtype)
    if weights is None:
        dist = (xyz_d[:, None, :] - bones_d[None, :, :]).norm(dim=-1).clamp(min=1e-4)
        w = 1 / dist
        w = w / w.sum(dim=1, keepdim=True)
    else:
        w = weights.to(dtype)
    y = torch.einsum('bij,pbj->pbi', R, xyz_d[:, None, :] - bones_d[None]) + motions_d[None] + bones_d[None]
    xyz_t = (y * w[:, :, None]).sum(1)
    rot = None
    if quat is not None:
        bq = torch.nn.functional.normalize(mat2quat(R), dim=-1)
        qs = torch.nn.functional.normalize((bq[None] * w[:, :, None]).sum(1), dim=-1)
        q2 = quat.to(dtype)
        rot = torch.stack([qs[:, 0] * q2[:, 0] - qs[:, 1] * q2[:, 1] - qs[:, 2] * q2[:, 2] - qs[:, 3] * q2[:, 3],
                           qs[:, 0] * q2[:, 1] + qs[:, 1] * q2[:, 0] + qs[:, 2] * q2[:, 3] - qs[:, 3] * q2[:, 2],
                           qs[:, 0] * q2[:, 2] - qs[:, 1] * q2[:, 3] + qs[:, 2] * q2[:, 0] + qs[:, 3] * q2[:, 1],
                           qs[:, 0] * q2[:, 3] + qs[:, 1] * q2[:, 2] - qs[:, 2] * q2[:, 1] + qs[:, 3] * q2[:, 0]], -1)
    return xyz_t, rot, w, R


def make_skinning_inputs(n_bones, n_particles, seed, special=True):
    """Seeded scene: bones in a 0.3 x 0.3 x 0.1 m slab, radius graph (incl. self edges like the rollout's top-k graph),
    cm-scale motions = a global rotation + noise.  special=True appends bones that exercise the reference's branches:
    an isolated bone, a bone whose only neighbour is itself + one other (rank 1), a coplanar neighbourhood (rank 2)."""
    rng = np.random.default_rng(seed)
    bones = rng.uniform([0, 0, 0], [0.3, 0.3, 0.1], size=(n_bones, 3))
    if special:
        bones[-1] = [2.0, 2.0, 2.0]                       # isolated (only its self edge -> F = 0 -> rank 0)
        bones[-2] = [1.0, 1.0, 1.0]; bones[-3] = [1.02, 1.01, 0.99]        # a pair: rank 1
        bones[-4] = [-1.0, 0.0, 0.5]; bones[-5] = [-1.03, 0.0, 0.52]; bones[-6] = [-0.98, 0.0, 0.47]; bones[-7] = [-1.01, 0.0, 0.55]  # coplanar (y = 0)
    ang = 0.15
    Rg = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    motions = (bones - bones.mean(0)) @ Rg.T + bones.mean(0) - bones + rng.normal(0, 0.003, size=(n_bones, 3))
    if special:
        motions[-7:-3, 1] = 0.0                           # keep the coplanar group coplanar after the motion
    d = np.linalg.norm(bones[:, None] - bones[None], axis=-1)
    rel = (d < 0.08).astype(np.int64)
    xyz = rng.uniform([-0.02, -0.02, -0.02], [0.32, 0.32, 0.12], size=(n_particles, 3))
    k = min(3, n_particles)
    xyz[:k] = bones[:k]                                   # particles sitting exactly on a bone (clamp at 1e-4)
    quat = rng.normal(size=(n_particles, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    t = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(bones=t(bones), motions=t(motions), relations=torch.tensor(rel), xyz=t(xyz), quat=t(quat))


def rollout(sd, cfg, dm, xyz_0, quat_0, eef_xyz, n_steps, inlier_idx_all, start_idx_fn, dtype=torch.float32):
    """CPU restatement of DynamicsModule.rollout, /root/reference/src/render/dynamics_module.py:53-172, from the oracle pieces
    (oracle.gnn_oracle: fps, fps_radius, construct_edges, one_hot_edges, forward; interpolate_motions above).
    dm: dict(n_his, dist_thresh, max_nobj, adj_thresh, fps_radius, topk, connect_all).  Returns (xyz, quat, xyz_bones, eef)."""
    from oracle import gnn_oracle as GO
    n_his = dm['n_his']

    def downsample(x):
        idx1 = GO.fps(x[None], min(dm['max_nobj'], x.shape[0]), start_idx=0)[0]
        sub = x[idx1]
        idx2 = GO.fps_radius(sub, dm['fps_radius'], start_idx_fn(sub.shape[0]))
        idx = idx1[idx2]
        return x[idx], idx

    all_pos = xyz_0
    fps_all_idx = GO.fps(xyz_0[inlier_idx_all][None], min(1000, len(inlier_idx_all)), start_idx=0)[0]
    fps_all_pos = all_pos[inlier_idx_all][fps_all_idx]
    hist = fps_all_pos[None].repeat(n_his, 1, 1)
    eef_hist = eef_xyz[0][None].repeat(n_his, 1, 1)
    eef_pos = eef_xyz[0]
    p0, _ = downsample(fps_all_pos)
    quat = quat_0[None].repeat(n_steps, 1, 1)
    xyz = xyz_0[None].repeat(n_steps, 1, 1)
    bones_out = torch.zeros(n_steps, dm['max_nobj'], 3)
    eef = eef_xyz[0][None].repeat(n_steps, 1, 1)
    bones_out[0, :p0.shape[0]] = p0
    for i in range(1, n_steps):
        if torch.norm(eef_xyz[i] - eef_pos) < dm['dist_thresh']:
            quat[i], xyz[i], bones_out[i], eef[i] = quat[i - 1], xyz[i - 1], bones_out[i - 1], eef[i - 1]
            continue
        eef_delta = eef_xyz[i] - eef_pos
        particle_pos, fps_idx = downsample(fps_all_pos)
        nobj = particle_pos.shape[0]
        states = torch.zeros((1, n_his, nobj + 1, 3))
        states[:, :, :nobj] = hist[:, fps_idx]
        states[:, :, nobj:] = eef_hist
        action = torch.zeros((1, nobj + 1, 3))
        action[:, nobj:] = eef_delta
        attrs = torch.zeros((1, nobj + 1, 2))
        attrs[:, :nobj, 0] = 1.
        attrs[:, nobj:, 1] = 1.
        mask = torch.ones(nobj + 1, dtype=torch.bool)
        tool = torch.zeros(nobj + 1, dtype=torch.bool)
        tool[nobj] = True
        recv, send = GO.construct_edges(states[0, -1], dm['adj_thresh'], mask, tool, dm['topk'], dm['connect_all'])
        Rr, Rs = GO.one_hot_edges(recv, send, nobj + 1)
        pred, _ = GO.forward(sd, cfg, states, attrs, Rr[None], Rs[None], torch.ones((1, nobj, 1)), action)
        eef_hist = torch.cat([eef_hist[1:], eef_xyz[i][None]], 0)
        eef_pos = eef_xyz[i]
        rel = relations_to_matrix(Rr[None], Rs[None])[:nobj, :nobj]
        x, q, _, _ = interpolate_motions(particle_pos, pred[0] - particle_pos, rel, all_pos, quat=quat[i - 1], dtype=dtype)
        all_pos = x.float()
        fps_all_pos = all_pos[inlier_idx_all][fps_all_idx]
        hist = torch.cat([hist[1:], fps_all_pos[None]], 0)
        quat[i], xyz[i] = q.float(), all_pos
        bones_out[i, :nobj] = pred[0]
        eef[i] = eef_pos
    return xyz, quat, bones_out, eef
