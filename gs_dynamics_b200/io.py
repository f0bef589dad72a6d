"""On-disk formats either side of the tracking path and the episode driver of
/root/reference/src/tracking/train_gs.py, so that the tracker is a drop-in for the rest of the reference pipeline:

  in : `<data>/<seq>/<metadata.json>`   {w, h, k[t][c] 3x3, w2c[t][c] 4x4, fn[t][c], cam_id[t][c]}   (utils/metadata.py:95-105)
       `<data>/<seq>/<init_pt_cld.npz>`  ["data"] float64 [n,7] = xyz, rgb, seg                          (utils/init_pcd.py:106-131)
       images `fn`, masks `<cam dir>/seg/seg_{frame:06}.png`                                              (train_utils.py:10-19)
  out: `<output>/<exp>/<seq>/params.npz`  stacked per-frame means3D / rgb_colors / unnorm_rotations + the frame-0 rest
       (helpers.py:132-148), consumed by preprocess.py:199-206 and render/dynamics_module.py:177-184
       `.splat` export of real_world/gs/convert.py:23-51 (32 bytes per Gaussian)
"""
import json
import os
from random import randint

import numpy as np
import torch

from . import tracking as TR


def map_to_segmentation_path(img_path):
    """train_utils.py:10-19: `cam/<dir>/<name>_<n>.png` -> `cam/seg/seg_{n:06}.png`."""
    directory, filename = img_path.rsplit('/', 1)
    directory = directory.rsplit('/', 1)[0]
    number = int(filename.split('_')[-1].split('.')[0])
    return f'{directory}/seg/seg_{number:06}.png'


def map_to_depth_path(img_path):
    """train_utils.py:21-29."""
    directory, filename = img_path.rsplit('/', 1)
    number = int(filename.split('_')[1].split('.')[0])
    return f'{directory}/depth/depth_{number:06}.png'


def load_metadata(seq, metadata_path, data_root="./data"):
    return json.load(open(os.path.join(data_root, seq, metadata_path), 'r'))


def get_custom_dataset(t, md, seq, data_root="./data", device="cuda"):
    """train_utils.py:32-78: per camera {'cam', 'im' [3,H,W] in [0,1], 'seg' (seg, 0, 1-seg) [3,H,W], 'id'}; near = 1, far = 100."""
    from PIL import Image
    dataset = []
    for c in range(len(md['fn'][t])):
        w, h = md['w'], md['h']
        cam = TR.setup_camera(w, h, md['k'][t][c], md['w2c'][t][c], near=1.0, far=100, device=device)
        fn = md['fn'][t][c]
        im = np.array(Image.open(os.path.join(data_root, seq, fn)))
        im = torch.tensor(im).float().to(device).permute(2, 0, 1)[:3].contiguous() / 255
        seg = np.array(Image.open(os.path.join(data_root, seq, map_to_segmentation_path(fn)))).astype(np.float32)
        seg = torch.tensor(seg).float().to(device)
        seg_col = torch.stack((seg, torch.zeros_like(seg), 1 - seg))
        dataset.append({'cam': cam, 'im': im, 'seg': seg_col, 'id': c})
    return dataset


def get_batch(todo_dataset, dataset):
    """train_utils.py:81-85.  The reference refills by REBINDING its local (`todo_dataset = dataset.copy()`), so the caller's
    list stays empty and every call draws one camera uniformly WITH replacement using one randint(0, len - 1): reproduced
    exactly (same draws for the same random.seed), the caller's list is left untouched."""
    todo = todo_dataset if todo_dataset else list(dataset)
    return todo.pop(randint(0, len(todo) - 1))


def initialize_params(seq, md, init_pt_cld_path, data_root="./data", device="cuda"):
    """train_utils.py:88-149 from the files of a sequence."""
    init_pt_cld = np.load(os.path.join(data_root, seq, init_pt_cld_path))["data"]
    cam_centers = np.linalg.inv(np.asarray(md['w2c'][0], np.float64))[:, :3, 3]
    return TR.initialize_params_from_point_cloud(init_pt_cld, cam_centers, device=device)


def params2cpu(params, is_initial_timestep):
    """helpers.py:122-129."""
    keep = None if is_initial_timestep else ('means3D', 'rgb_colors', 'unnorm_rotations')
    return {k: v.detach().cpu().contiguous().numpy() for k, v in params.items() if keep is None or k in keep}


def save_params(output_params, seq, exp, out_root="./output"):
    """helpers.py:132-140: keys present after frame 0 are stacked over frames, the rest saved from frame 0."""
    to_save = {}
    for k in output_params[0].keys():
        if len(output_params) > 1 and k in output_params[1].keys():
            to_save[k] = np.stack([p[k] for p in output_params])
        else:
            to_save[k] = output_params[0][k]
    os.makedirs(os.path.join(out_root, exp, seq), exist_ok=True)
    path = os.path.join(out_root, exp, seq, "params")
    np.savez(path, **to_save)
    return path + ".npz"


def load_scene(params_path, frame=0, device="cuda"):
    """What the consumers of params.npz read (render/dynamics_module.py:177-190, preprocess.py:199-206): activated Gaussians
    of one frame: xyz [n,3], rgb [n,3], unit quaternions [n,4], opacities [n,1], scales [n,3]."""
    p = {k: torch.tensor(v).to(device).float() for k, v in dict(np.load(params_path)).items()}
    pick = lambda v, d: v[frame] if v.dim() == d + 1 else v
    return dict(xyz=pick(p['means3D'], 2), rgb=pick(p['rgb_colors'], 2),
                quat=torch.nn.functional.normalize(pick(p['unnorm_rotations'], 2)), opa=torch.sigmoid(p['logit_opacities']),
                scales=torch.exp(p['log_scales']))


def save_to_splat(pts, colors, scales, quats, opacities, output_file):
    """real_world/gs/convert.py:23-51, vectorised: centre the cloud, rotate by -90 deg about x, 32 bytes per Gaussian =
    position f32[3] | scale f32[3] | rgba u8[4] | quaternion u8[4] ((q/|q|)*128+128)."""
    pts = np.asarray(pts)
    pts = pts - np.mean(pts, axis=0)
    n = pts.shape[0]
    rot_inv = np.linalg.inv(np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float32))
    pos = (pts.astype(np.float32) @ rot_inv.T.astype(np.float32)).astype(np.float32)
    w = np.sqrt(1 + rot_inv[0, 0] + rot_inv[1, 1] + rot_inv[2, 2]) / 2
    qx = np.array([w, (rot_inv[2, 1] - rot_inv[1, 2]) / (4 * w), (rot_inv[0, 2] - rot_inv[2, 0]) / (4 * w),
                   (rot_inv[1, 0] - rot_inv[0, 1]) / (4 * w)], dtype=np.float32)
    q = np.asarray(quats, dtype=np.float32)
    w1, x1, y1, z1 = qx
    w2, x2, y2, z2 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot = np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                    w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1).astype(np.float32)
    color = np.concatenate([np.asarray(colors)[:, :3], np.asarray(opacities)[:, :1]], -1)
    rec = np.zeros(n, dtype=[('pos', '<f4', 3), ('scale', '<f4', 3), ('rgba', 'u1', 4), ('rot', 'u1', 4)])
    rec['pos'] = pos
    rec['scale'] = np.asarray(scales, dtype=np.float32)[:, :3]
    rec['rgba'] = (color * 255).clip(0, 255).astype(np.uint8)
    rec['rot'] = ((rot / np.linalg.norm(rot, axis=1, keepdims=True)) * 128 + 128).clip(0, 255).astype(np.uint8)
    with open(output_file, "wb") as f:
        f.write(rec.tobytes())


def train(seq, exp, remove_threshold=0.005, remove_thresh_5k=0.25, weight_soft_col_cons=0.01, weight_im=50.0, weight_seg=200.0,
          weight_rigid=200.0, weight_bg=200.0, weight_iso=1000.0, weight_rot=4.0, num_knn=20, scale_scene_radius=0.05,
          metadata_path="train_meta.json", init_pt_cld_path="init_pt_cld.npz", data_root="./data", out_root="./output",
          iters_first=10000, iters_next=2000, device="cuda"):
    """train_gs.py:10-46.  Frame 0: the reference's loop (get_loss / backward / densify / Adam) on the B200 kernels; every later
    frame: `FusedTrackingStep` (one CUDA graph per camera).  Writes params.npz like the reference and returns its path."""
    md = load_metadata(seq, metadata_path, data_root)
    num_timesteps = len(md['fn'])
    params, variables = initialize_params(seq, md, init_pt_cld_path, data_root, device)
    optimizer = TR.initialize_optimizer(params, variables)
    loss_kwargs = dict(weight_soft_col_cons=weight_soft_col_cons, weight_im=weight_im, weight_seg=weight_seg, weight_rigid=weight_rigid,
                       weight_bg=weight_bg, weight_iso=weight_iso, weight_rot=weight_rot)
    output_params, path = [], None
    step = None
    for t in range(num_timesteps):
        dataset = get_custom_dataset(t, md, seq, data_root, device)
        todo_dataset = []
        if t == 0:
            for i in range(iters_first):
                curr_data = get_batch(todo_dataset, dataset)
                loss, variables = TR.get_loss(params, curr_data, variables, True, fused=False, **loss_kwargs)
                loss.backward()
                with torch.no_grad():
                    params, variables, num_pts = TR.densify(params, variables, optimizer, i, remove_threshold, remove_thresh_5k, scale_scene_radius)
                    optimizer.step()
                    optimizer.zero_grad(set_to_none=True)
            os.makedirs(os.path.join(out_root, exp, seq), exist_ok=True)
            with open(os.path.join(out_root, exp, seq, "num_pts.txt"), 'w') as f:
                f.write(f"Number of points: {params['means3D'].shape[0]}\n")
        else:
            params, variables = TR.initialize_per_timestep(params, variables, optimizer)
            if step is None:   # graphs captured once, reused for every later frame (only the targets change)
                step = TR.FusedTrackingStep(params, variables, optimizer, dataset, loss_kwargs=loss_kwargs)
                step.prepare()
            else:
                step.set_targets(dataset)
            draws = [get_batch(todo_dataset, dataset) for _ in range(iters_next)]
            TR.run_frame(step, [next(j for j, d in enumerate(dataset) if d is cd) for cd in draws])
            variables = step.variables
        output_params.append(params2cpu(params, t == 0))
        if t == 0:
            variables = TR.initialize_post_first_timestep(params, variables, optimizer, num_knn)
        if (t % 5 == 0 and t > 0) or t == num_timesteps - 1:
            path = save_params(output_params, seq, exp, out_root)
    return path
