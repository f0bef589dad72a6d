"""oracle/tracking_cpu.py — TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT CODE.

The reference's tracking iteration (get_loss + backward + Adam, /root/reference/src/tracking/train_utils.py:167-246,
train_gs.py:25-39) on the HOST cores: the rasterizer is the C restatement (oracle/raster_oracle.c, OpenMP), every other
op is CPU PyTorch exactly as the reference composes it (two separate renders, conv2d SSIM, boolean-mask indexing,
torch.optim.Adam).  The reference has no CPU path of its own (its rasterizer is CUDA-only and un-vendored), so this port
is what bench.py times as `cpu_baseline` / `--impl reference` (kind = "port").
"""
import numpy as np
import torch

from . import raster_c
from . import tracking_oracle as T


class _RasterCPU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, colors, scales, rotations, cam, bg):
        args = (means3D.detach(), colors.detach(), opacities.detach(), scales.detach(), rotations.detach(), cam["viewmatrix"],
                cam["projmatrix"], bg, cam["tanfovx"], cam["tanfovy"], cam["image_height"], cam["image_width"])
        out = raster_c.forward(*args)
        ctx.args = args
        ctx.opac_shape = opacities.shape
        return torch.from_numpy(out["color"]), torch.from_numpy(out["radii"].copy()), torch.from_numpy(out["depth"])

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth):
        g = raster_c.backward(*ctx.args, g_color.contiguous())
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        return (t(g["means3D"]), t(g["means2D"]), t(g["opacities"]).reshape(ctx.opac_shape), t(g["colors"]), t(g["scales"]),
                t(g["rotations"]), None, None)


def render(cam, rendervar, bg):
    return _RasterCPU.apply(rendervar["means3D"], rendervar["means2D"], rendervar["opacities"], rendervar["colors_precomp"],
                            rendervar["scales"], rendervar["rotations"], cam, bg)


def params2rendervar(params):
    return {"means3D": params["means3D"], "colors_precomp": params["rgb_colors"],
            "rotations": torch.nn.functional.normalize(params["unnorm_rotations"]),
            "opacities": torch.sigmoid(params["logit_opacities"]), "scales": torch.exp(params["log_scales"]),
            "means2D": torch.zeros_like(params["means3D"], requires_grad=True) + 0}


def get_loss(params, data, variables, is_initial_timestep, weight_im=50.0, weight_seg=200.0, weight_rigid=200.0,
             weight_bg=200.0, weight_iso=1000.0, weight_rot=4.0):
    bg = torch.zeros(3)
    rv = params2rendervar(params)
    im, radius, _ = render(data["cam"], rv, bg)
    cid = data["id"]
    im = torch.exp(params["cam_m"][cid])[:, None, None] * im + params["cam_c"][cid][:, None, None]
    loss = weight_im * T.photometric(im, data["im"])
    srv = params2rendervar(params)
    srv["colors_precomp"] = params["seg_colors"]
    seg, _, _ = render(data["cam"], srv, bg)
    loss = loss + weight_seg * T.photometric(seg, data["seg"])
    if not is_initial_timestep:
        is_fg = (params["seg_colors"][:, 0] > 0.5).detach()
        L = T.prior_losses(rv["means3D"], rv["rotations"], is_fg, variables["prev_inv_rot_fg"], variables["neighbor_indices"],
                           variables["neighbor_weight"], variables["neighbor_dist"], variables["prev_offset"],
                           variables["init_bg_pts"], variables["init_bg_rot"])
        loss = loss + weight_rigid * L["rigid"] + weight_rot * L["rot"] + weight_iso * L["iso"] + 2.0 * L["floor"] + weight_bg * L["bg"]
    seen = radius > 0
    variables["max_2D_radius"][seen] = torch.max(radius[seen].float(), variables["max_2D_radius"][seen])
    variables["seen"] = seen
    return loss, variables


def make_optimizer(params, scene_radius):
    lrs = {"means3D": 0.00016 * scene_radius, "rgb_colors": 0.0, "seg_colors": 0.0, "unnorm_rotations": 0.001,
           "logit_opacities": 0.0, "log_scales": 0.0, "cam_m": 0.0, "cam_c": 0.0}  # after t = 0 (train_utils.py:370-373)
    groups = [{"params": [v], "name": k, "lr": lrs[k]} for k, v in params.items()]
    return torch.optim.Adam(groups, lr=0.0, eps=1e-15)


def iteration(params, data, variables, opt):
    loss, variables = get_loss(params, data, variables, False)
    loss.backward()
    with torch.no_grad():
        opt.step()
        opt.zero_grad(set_to_none=True)
    return float(loss.detach())
