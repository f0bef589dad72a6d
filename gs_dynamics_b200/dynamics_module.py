"""`DynamicsModule` — host-side mirror of /root/reference/src/render/dynamics_module.py:15-172 on the B200 kernels:
FPS (`gsd_fps`), edge build (`gsd_gnn_build_edges`), the GNN forward, and the Gaussian skinning step (`gsd_skin_*`).

Same rollout semantics as the reference (`rollout` keeps its argument list and return values); what changes is where the work
runs: nothing is copied to the host inside the loop (the reference moves every step's result with `.cpu()`,
dynamics_module.py:165-170), edges stay index lists, and the skinning step never materialises [n_particles, n_bones] tensors.
The one host round trip per step that remains is the particle count of the radius-terminated FPS, which sets the graph size
(the reference's own loop syncs once per picked particle)."""
import os

import numpy as np
import torch

from .gnn import (DynamicsPredictor, construct_edges_index, farthest_point_sampler, fps_rad_idx_torch)
from .skinning import interpolate_motions


class DynamicsModule:
    def __init__(self, config=None, epoch='latest', device='cuda', model=None):
        """`DynamicsModule(config, epoch, device)` as in the reference (loads `<out_dir>/checkpoints/{latest,model_N}.pth`), or
        `DynamicsModule(config, model=...)` with an already built `DynamicsPredictor` (tests, benchmarks)."""
        self.device = torch.device(device)
        train_config = config['train_config']
        model_config = config['model_config']
        if model is None:
            name = 'latest.pth' if epoch == 'latest' else 'model_{}.pth'.format(epoch)
            model = self.load_model(train_config, model_config, os.path.join(train_config['out_dir'], 'checkpoints', name), self.device)
        self.model = model
        self.n_his = train_config['n_his']
        self.dist_thresh = train_config['dist_thresh']
        dataset_config = config['dataset_config']['datasets'][0]
        self.max_nobj = dataset_config['max_nobj']
        self.adj_thresh = (dataset_config['adj_radius_range'][0] + dataset_config['adj_radius_range'][1]) / 2
        self.fps_radius = (dataset_config['fps_radius_range'][0] + dataset_config['fps_radius_range'][1]) / 2
        self.topk = dataset_config['topk']
        self.connect_all = dataset_config['connect_all']
        self.start_idx_fn = None   # callable(n) -> first index of the radius FPS; None = np.random.randint like data/utils.py:54

    def load_model(self, train_config, model_config, checkpoint_dir, device):
        model_config['n_his'] = train_config['n_his']
        model = DynamicsPredictor(model_config, device)
        model.to(device)
        model.eval()
        model.load_state_dict(torch.load(checkpoint_dir, map_location=device))
        return model

    def downsample_vertices(self, xyz):  # (n, 3)
        """dynamics_module.py:44-51: FPS to max_nobj from index 0, then radius-terminated FPS; returns (points, indices)."""
        n = xyz.shape[0]
        idx1 = farthest_point_sampler(xyz[None], min(self.max_nobj, n), start_idx=0)[0]
        sub = xyz[idx1]
        start = self.start_idx_fn(sub.shape[0]) if self.start_idx_fn is not None else int(np.random.randint(sub.shape[0]))
        _, idx2 = fps_rad_idx_torch(sub, self.fps_radius, start_idx=start)
        fps_idx = idx1[idx2]
        return xyz[fps_idx], fps_idx

    @torch.no_grad()
    def rollout(self, xyz_0, rgb_0, quat_0, opa_0, eef_xyz, n_steps, inlier_idx_all):
        """dynamics_module.py:53-172.  xyz_0 [n,3], rgb_0 [n,3], quat_0 [n,4], opa_0 [n,1], eef_xyz [n_steps,1,3];
        returns (xyz, rgb, quat, opa, xyz_bones, eef) as CPU tensors with a leading n_steps axis, like the reference."""
        model, device = self.model, self.device
        n_his = model.model_config['n_his']
        xyz_0, quat_0 = xyz_0.to(device).float(), quat_0.to(device).float()
        eef_xyz = eef_xyz.to(device).float()
        eef_host = eef_xyz.detach().cpu()                       # the skip test below runs on the host copy: no sync per step
        inlier = torch.as_tensor(inlier_idx_all, device=device)

        all_pos = xyz_0
        fps_all_idx = farthest_point_sampler(xyz_0[inlier][None], min(1000, inlier.numel()), start_idx=0)[0]
        fps_all_pos = all_pos[inlier][fps_all_idx]
        fps_all_pos_history = fps_all_pos[None].repeat(n_his, 1, 1)
        eef_pos_history = eef_xyz[0][None].repeat(n_his, 1, 1)  # (n_his, 1, 3)
        eef_pos_host = eef_host[0]
        particle_pos_0, _ = self.downsample_vertices(fps_all_pos)

        quat = quat_0[None].repeat(n_steps, 1, 1)
        xyz = xyz_0[None].repeat(n_steps, 1, 1)
        xyz_bones = torch.zeros(n_steps, self.max_nobj, 3, device=device)
        eef = eef_xyz[0][None].repeat(n_steps, 1, 1)
        xyz_bones[0, :particle_pos_0.shape[0]] = particle_pos_0

        for i in range(1, n_steps):
            if float(torch.norm(eef_host[i] - eef_pos_host)) < self.dist_thresh:
                quat[i], xyz[i], xyz_bones[i], eef[i] = quat[i - 1], xyz[i - 1], xyz_bones[i - 1], eef[i - 1]
                continue
            eef_pos_this_step = eef_xyz[i]
            eef_delta = eef_pos_this_step - eef_pos_history[-1]

            particle_pos, fps_idx = self.downsample_vertices(fps_all_pos)
            particle_pos_history = fps_all_pos_history[:, fps_idx]
            nobj = particle_pos.shape[0]

            states = torch.zeros((1, n_his, nobj + 1, 3), device=device)
            states[:, :, :nobj] = particle_pos_history
            states[:, :, nobj:] = eef_pos_history
            states_delta = torch.zeros((1, nobj + 1, 3), device=device)
            states_delta[:, nobj:] = eef_delta
            attrs = torch.zeros((1, nobj + 1, 2), dtype=torch.float32, device=device)
            attrs[:, :nobj, 0] = 1.
            attrs[:, nobj:, 1] = 1.
            p_instance = torch.ones((1, nobj, 1), dtype=torch.float32, device=device)
            state_mask = torch.ones((1, nobj + 1), dtype=torch.bool, device=device)
            eef_mask = torch.zeros((1, nobj + 1), dtype=torch.bool, device=device)
            eef_mask[:, nobj] = True

            edges = construct_edges_index(states[:, -1], self.adj_thresh, state_mask, eef_mask, topk=self.topk,
                                          connect_all=self.connect_all, n_tool=1)
            pred_state, _ = model(states, attrs, edges, None, p_instance, action=states_delta)  # (1, nobj, 3)

            eef_pos_history = torch.cat([eef_pos_history[1:], eef_pos_this_step[None]], dim=0)
            eef_pos_host = eef_host[i]

            # skin all Gaussians from the particles (bones); the tool node of `edges` is ignored (relations[:nobj, :nobj])
            all_pos, all_rot, _ = interpolate_motions(bones=particle_pos, motions=pred_state[0] - particle_pos, relations=edges,
                                                      xyz=all_pos, quat=quat[i - 1], return_weights=False)
            fps_all_pos = all_pos[inlier][fps_all_idx]
            fps_all_pos_history = torch.cat([fps_all_pos_history[1:], fps_all_pos[None]], dim=0)

            quat[i], xyz[i] = all_rot, all_pos
            xyz_bones[i, :nobj] = pred_state[0]
            eef[i] = eef_pos_this_step

        rgb = rgb_0.detach().cpu()[None].repeat(n_steps, 1, 1)
        opa = opa_0.detach().cpu()[None].repeat(n_steps, 1, 1)
        return xyz.cpu(), rgb, quat.cpu(), opa, xyz_bones.cpu(), eef.cpu()
