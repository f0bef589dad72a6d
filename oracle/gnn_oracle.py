"""oracle/gnn_oracle.py — TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT CODE.

CPU restatement (PyTorch, dense one-hot formulation exactly like the reference) of
  DynamicsPredictor.forward              /root/reference/src/gnn/model.py:112-246
  construct_edges_from_states            /root/reference/src/data/dataset.py:88-147
  dgl.geometry.farthest_point_sampler    (DGL is un-vendored and unpinned: requirements.txt:4, README.md:23; restated as
                                          batched FPS: idx[0]=start, dist=+inf, dist=min(dist,|x-x[idx]|^2), argmax first max)
  fps_rad_idx_torch                      /root/reference/src/data/utils.py:50-65
Pinned against the reference itself by tests/golden/gnn_golden.npz (tools/make_golden.py imports /root/reference/src/gnn/model.py
and src/data/dataset.py).  The model is a pure function of a state_dict so that no nn.Module code is duplicated.
Only tests/, smoke() and bench.py's cpu legs may import this module.
"""
import torch
import torch.nn.functional as F

# seeded input / weight builders are plain data generators shared with bench.py; they live in the product package
from gs_dynamics_b200.workloads import (model_dims, make_state_dict, sloth_cfg, rope_cfg, make_graph_inputs,  # noqa: F401
                                        make_training_batch)


def _mlp3(sd, name, x):
    for i in (0, 2, 4):
        x = F.relu(F.linear(x, sd[f"{name}.model.{i}.weight"], sd[f"{name}.model.{i}.bias"]))
    return x


def forward(sd, cfg, state, attrs, Rr, Rs, p_instance, action=None):
    """Dense one-hot forward, op for op as model.py:112-246. Returns (pred_pos, pred_motion)."""
    n_his = cfg['n_his']
    B, N = attrs.shape[:2]
    n_p, n_inst = p_instance.shape[1], p_instance.shape[2]
    n_s = N - n_p
    sdim = state.shape[3]
    Rr_t = Rr.transpose(1, 2).contiguous()
    state_t = state.transpose(1, 2).contiguous().view(B, N, n_his * sdim)
    p_inputs = attrs
    if cfg['state_dim'] == 3:
        p_inputs = torch.cat([p_inputs, state_t], 2)
    elif cfg['state_dim'] == 1:
        p_inputs = torch.cat([attrs, state_t.view(B, N, n_his, sdim)[:, :, :, 2]], 2)
    if cfg.get('motion_dim', 0) > 0:
        xyz = state_t.view(B, N, n_his, sdim)
        p_inputs = torch.cat([p_inputs, (xyz[:, :, 1:] - xyz[:, :, :-1]).reshape(B, N, (n_his - 1) * 3)], 2)
    if cfg['action_dim'] > 0:
        p_inputs = torch.cat([p_inputs, action], 2)
    g = torch.cat([p_instance, torch.zeros(B, n_s, n_inst, dtype=state.dtype)], 1)
    rel_inputs = torch.cat([Rr.bmm(attrs), Rs.bmm(attrs), torch.sum(torch.abs(Rr.bmm(g) - Rs.bmm(g)), 2, keepdim=True),
                            Rr.bmm(state_t) - Rs.bmm(state_t)], 2)
    particle_encode = _mlp3(sd, "particle_encoder", p_inputs)
    relation_encode = _mlp3(sd, "relation_encoder", rel_inputs)
    effect = particle_encode
    for _ in range(cfg['pstep']):
        er, es = Rr.bmm(effect), Rs.bmm(effect)
        rel = F.relu(F.linear(torch.cat([relation_encode, er, es], 2), sd["relation_propagator.linear.weight"],
                              sd["relation_propagator.linear.bias"]))
        agg = Rr_t.bmm(rel)
        effect = F.relu(F.linear(torch.cat([particle_encode, agg], 2), sd["particle_propagator.linear.weight"],
                                 sd["particle_propagator.linear.bias"]) + effect)
    x = effect[:, :n_p].contiguous()
    x = F.relu(F.linear(x, sd["non_rigid_predictor.linear_0.weight"], sd["non_rigid_predictor.linear_0.bias"]))
    x = F.relu(F.linear(x, sd["non_rigid_predictor.linear_1.weight"], sd["non_rigid_predictor.linear_1.bias"]))
    motion = F.linear(x, sd["non_rigid_predictor.linear_2.weight"], sd["non_rigid_predictor.linear_2.bias"])
    return state[:, -1, :n_p] + torch.clamp(motion, max=100.0, min=-100.0), motion


def construct_edges(states, adj_thresh, mask, tool_mask, topk=10, connect_all=False):
    """Dense restatement of dataset.py:88-147. Returns (receivers, senders) int64 in adj.nonzero() order."""
    N = states.shape[0]
    diff = states[:, None, :] - states[None, :, :]
    dis = torch.sum(diff ** 2, -1)
    m12 = mask[:, None] & mask[None, :]
    dis[~m12] = 1e10
    t12 = tool_mask[:, None] & tool_mask[None, :]
    dis[t12] = 1e10
    adj = ((dis - adj_thresh * adj_thresh) < 0).float()
    k = min(N, topk)
    n_tool = int(tool_mask.sum())
    dis_obj = dis[:-n_tool, :-n_tool] if n_tool > 0 else dis
    idx = torch.topk(dis_obj, k=k, dim=-1, largest=False)[1]
    tk = torch.zeros_like(dis_obj)
    tk.scatter_(-1, idx, 1)
    if n_tool > 0:
        adj[:-n_tool, :-n_tool] = adj[:-n_tool, :-n_tool] * tk
    else:
        adj = adj * tk
    if connect_all:
        adj[tool_mask[:, None] & mask[None, :]] = 1.
        adj[tool_mask[None, :] & mask[:, None]] = 1.
        adj[t12] = 0.
    rels = adj.nonzero()
    return rels[:, 0], rels[:, 1]


def one_hot_edges(recv, send, N):
    E = recv.shape[0]
    Rr, Rs = torch.zeros(E, N), torch.zeros(E, N)
    ar = torch.arange(E)
    Rr[ar, recv] = 1
    Rs[ar, send] = 1
    return Rr, Rs


def fps(pos, npoints, start_idx=0):
    """pos [B,N,3] -> int64 [B,npoints]."""
    B, N, _ = pos.shape
    out = torch.zeros(B, npoints, dtype=torch.int64)
    for b in range(B):
        dist = torch.full((N,), float("inf"))
        cur = int(start_idx if isinstance(start_idx, int) else start_idx[b])
        for i in range(npoints):
            out[b, i] = cur
            d = ((pos[b] - pos[b, cur]) ** 2).sum(-1)
            dist = torch.minimum(dist, d)
            cur = int(torch.argmax(dist))
    return out


def fps_radius(pcd, radius, start_idx):
    idx = [int(start_idx)]
    dist = torch.norm(pcd - pcd[idx[0]], dim=1)
    while dist.max() > radius:
        a = int(dist.argmax())
        idx.append(a)
        dist = torch.minimum(dist, torch.norm(pcd - pcd[a], dim=1))
    return torch.tensor(idx)




# ---------------------------------------------------------------------------------------------------------------------
# GNN training: losses and the n_future-step unroll of /root/reference/src/train.py, dense one-hots
# ---------------------------------------------------------------------------------------------------------------------
def length_loss(pred, state, Rr, Rs):
    """/root/reference/src/train.py:66-83"""
    n_p = pred.shape[1]
    pos = state[:, 0, :n_p].detach()
    Rr, Rs = Rr[:, :, :n_p], Rs[:, :, :n_p]
    pos_diff = Rr.bmm(pos) - Rs.bmm(pos)
    pred_diff = Rr.bmm(pred) - Rs.bmm(pred)
    return F.mse_loss(torch.norm(pred_diff, dim=-1), torch.norm(pos_diff, dim=-1))


def local_rigid_loss(pred, state, Rr, Rs):
    """/root/reference/src/train.py:85-102"""
    n_p = pred.shape[1]
    pos = state[:, 0, :n_p].detach()
    Rr, Rs = Rr[:, :, :n_p], Rs[:, :, :n_p]
    diff_r = torch.norm(Rr.bmm(pred) - Rr.bmm(pos), dim=-1)
    diff_s = torch.norm(Rs.bmm(pred) - Rs.bmm(pos), dim=-1)
    return F.mse_loss(diff_r, diff_s)


def batch_edges(batch, adj_thresh, topk, connect_all):
    """Per-element construct_edges on the newest frame, one-hots padded with zero rows to a common n_rel
    (pad_torch, /root/reference/src/data/dataset.py:229-238)."""
    B, _, N, _ = batch["state"].shape
    rs = [construct_edges(batch["state"][b, -1], adj_thresh, batch["state_mask"][b], batch["eef_mask"][b], topk, connect_all)
          for b in range(B)]
    n_rel = max(r.numel() for r, _ in rs)
    Rr = torch.zeros(B, n_rel, N, dtype=batch["state"].dtype)
    Rs = torch.zeros(B, n_rel, N, dtype=batch["state"].dtype)
    for b, (r, s) in enumerate(rs):
        ar = torch.arange(r.numel())
        Rr[b, ar, r] = 1
        Rs[b, ar, s] = 1
    return Rr, Rs


def unrolled_loss(sd, cfg, batch, Rr, Rs, n_future, w_mse=1.0, w_len=0.01):
    """/root/reference/src/train.py:183-211 with loss_funcs = [(mse_loss, w_mse), (length_loss, w_len)].
    Returns (loss_sum, [[mse_i, len_i] per step])."""
    state, action = batch["state"], batch["action"]
    loss_sum, parts = 0, []
    for fi in range(n_future):
        gt = batch["state_future"][:, fi].clone()
        pred, _ = forward(sd, cfg, state, batch["attrs"], Rr, Rs, batch["p_instance"], action)
        pred_p = pred[:, :gt.shape[1], :3].clone()
        l_mse = w_mse * F.mse_loss(pred_p, gt)
        l_len = w_len * length_loss(pred_p, state, Rr, Rs)
        loss_sum = loss_sum + l_mse + l_len
        parts.append([l_mse, l_len])
        if fi < n_future - 1:
            next_state = batch["tool_future"][:, fi].clone().unsqueeze(1)
            next_state[:, -1, :pred_p.shape[1]] = pred_p
            state = torch.cat([state[:, 1:], next_state], dim=1)
            action = batch["action_future"][:, fi].clone()
    return loss_sum, parts
