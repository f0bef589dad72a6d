"""baseline/upstream_structure — BENCHMARK BASELINE, NOT PRODUCT CODE (nothing under gs_dynamics_b200/ imports this).

What the north-star calls "the reference's own path timed in the same run ... reference CUDA rasterizer on 1 GPU": the
reference's tracking iteration exactly as /root/reference/src/tracking composes it —

    get_loss (train_utils.py:167-246): params2rendervar, TWO rasterizer passes (RGB, seg), exp(cam_m) * im + cam_c,
        0.8 * L1 + 0.2 * (1 - SSIM) with the 11x11 window REBUILT ON THE HOST every call (external.py:101-111) and five grouped
        conv2d per call, boolean-mask indexing of the foreground points, quat_mult / build_rotation / weighted_l2 priors
        as ~15 materialised [G, 20, 3|4] temporaries, radius bookkeeping with boolean indexing
    loss.backward(); torch.optim.Adam(eps=1e-15).step(); zero_grad(set_to_none=True)            (train_gs.py:31-39)

— in eager PyTorch on the GPU, on top of `raster_upstream.cu`, a clearly-labelled STRUCTURAL RE-CREATION of the un-vendored
upstream rasterizer (see that file's header).  If the real `diff_gaussian_rasterization` extension is importable on the box,
`rasterizer_module(prefer_real=True)` returns it instead and the line says so.
"""
import ctypes as C
import math
import os
import subprocess

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgsu_upstream.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "raster_upstream.cu")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                               "-shared", "-o", LIB, src, "-cudart", "static"])
    return LIB


class _Args(C.Structure):
    _fields_ = [("P", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("scale_modifier", C.c_float)] + [(k, C.c_void_p) for k in (
                    "means3D", "colors", "opacities", "scales", "rotations", "viewmatrix", "projmatrix", "bg", "out_color", "out_depth",
                    "radii", "geom", "binning", "img")]


class _Grads(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("dL_dpix", "dL_dmean2D", "dL_dconic", "dL_dopacity", "dL_dcolor", "dL_dmean3D", "dL_dcov3D",
                                          "dL_dscale", "dL_drot")]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError("baseline library not built: run __graft_entry__.build()")
        l = C.CDLL(LIB)
        l.gsu_buffer_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_size_t)]
        l.gsu_buffer_bytes.restype = None
        l.gsu_forward_count.argtypes = [C.POINTER(_Args), C.c_void_p]
        l.gsu_forward_count.restype = C.c_int64
        l.gsu_forward_render.argtypes = [C.POINTER(_Args), C.c_int64, C.c_void_p]
        l.gsu_forward_render.restype = C.c_int64
        l.gsu_backward.argtypes = [C.POINTER(_Args), C.POINTER(_Grads), C.c_int64, C.c_void_p]
        l.gsu_backward.restype = C.c_int64
        _lib = l
    return _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Rasterize(torch.autograd.Function):
    """The upstream autograd function's contract: (color, radii, depth); geometry / binning / image buffers allocated through
    torch once num_rendered is known (upstream's resize callbacks), kept on ctx for the backward."""

    @staticmethod
    def forward(ctx, means3D, means2D, opacities, colors, scales, rotations, rs):
        l = lib()
        P, H, W = means3D.shape[0], int(rs.image_height), int(rs.image_width)
        dev = means3D.device
        t = [x.contiguous().float() for x in (means3D, colors, opacities, scales, rotations)]
        view, proj, bg = rs.viewmatrix.contiguous().reshape(-1), rs.projmatrix.contiguous().reshape(-1), rs.bg.contiguous()
        sz = (C.c_size_t * 3)()
        l.gsu_buffer_bytes(P, W, H, 0, sz)
        geom = torch.empty(sz[0], dtype=torch.uint8, device=dev)
        img = torch.empty(sz[2], dtype=torch.uint8, device=dev)
        color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros(P, dtype=torch.int32, device=dev)
        a = _Args()
        a.P, a.W, a.H, a.tanfovx, a.tanfovy, a.scale_modifier = P, W, H, float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier)
        a.means3D, a.colors, a.opacities, a.scales, a.rotations = [x.data_ptr() for x in t]
        a.viewmatrix, a.projmatrix, a.bg = view.data_ptr(), proj.data_ptr(), bg.data_ptr()
        a.out_color, a.out_depth, a.radii = color.data_ptr(), depth.data_ptr(), radii.data_ptr()
        a.geom, a.img = geom.data_ptr(), img.data_ptr()
        R = l.gsu_forward_count(C.byref(a), _stream())        # blocks on the D2H copy of num_rendered, like upstream
        if R < 0:
            raise RuntimeError("gsu_forward_count failed: %d" % R)
        l.gsu_buffer_bytes(P, W, H, R, sz)
        binning = torch.empty(sz[1], dtype=torch.uint8, device=dev)
        a.binning = binning.data_ptr()
        rc = l.gsu_forward_render(C.byref(a), R, _stream())
        if rc < 0:
            raise RuntimeError("gsu_forward_render failed: %d" % rc)
        ctx.a, ctx.R, ctx.keep = a, R, (t, view, proj, bg, geom, binning, img, color, depth, radii)
        ctx.opac_shape = opacities.shape
        ctx.mark_non_differentiable(radii, depth)
        return color, radii, depth

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth):
        a, P = ctx.a, ctx.a.P
        dev = g_color.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)   # upstream zero-fills every gradient buffer per call
        g = dict(m2d=z(P, 3), conic=z(P, 4), opac=z(P, 1), col=z(P, 3), m3d=z(P, 3), cov=z(P, 6), sc=z(P, 3), rot=z(P, 4))
        gc = g_color.contiguous().float()
        gr = _Grads()
        gr.dL_dpix, gr.dL_dmean2D, gr.dL_dconic, gr.dL_dopacity, gr.dL_dcolor = (gc.data_ptr(), g["m2d"].data_ptr(), g["conic"].data_ptr(),
                                                                                 g["opac"].data_ptr(), g["col"].data_ptr())
        gr.dL_dmean3D, gr.dL_dcov3D, gr.dL_dscale, gr.dL_drot = g["m3d"].data_ptr(), g["cov"].data_ptr(), g["sc"].data_ptr(), g["rot"].data_ptr()
        rc = lib().gsu_backward(C.byref(a), C.byref(gr), ctx.R, _stream())
        if rc < 0:
            raise RuntimeError("gsu_backward failed: %d" % rc)
        ctx.keep = None
        return g["m3d"], g["m2d"], g["opac"].reshape(ctx.opac_shape), g["col"], g["sc"], g["rot"], None


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        return _Rasterize.apply(means3D, means2D, opacities, colors_precomp, scales, rotations, self.raster_settings)


def rasterizer_module(prefer_real=True):
    """(GaussianRasterizer class, label): the real upstream extension when it is installed on this box, else the re-creation."""
    if prefer_real:
        try:
            import importlib
            m = importlib.import_module("diff_gaussian_rasterization")
            if hasattr(m, "_C") and "gs_dynamics_b200" not in (getattr(m, "__file__", "") or ""):
                return m.GaussianRasterizer, "upstream diff_gaussian_rasterization (installed on this box)"
        except Exception:
            pass
    return GaussianRasterizer, "structural re-creation (baseline/upstream_structure/raster_upstream.cu)"


# ----------------------------------------------------------------------------------------------------------------------
# the reference's eager iteration (train_utils.py:167-246, helpers.py:36-94, external.py:25-135, train_gs.py:31-39)
# ----------------------------------------------------------------------------------------------------------------------
def params2rendervar(params):
    return {'means3D': params['means3D'], 'colors_precomp': params['rgb_colors'],
            'rotations': F.normalize(params['unnorm_rotations']), 'opacities': torch.sigmoid(params['logit_opacities']),
            'scales': torch.exp(params['log_scales']),
            'means2D': torch.zeros_like(params['means3D'], requires_grad=True, device="cuda") + 0}


def _window(size, channel):
    g = torch.Tensor([math.exp(-(x - size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(size)])   # host, every call
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, size, size).contiguous()


def calc_ssim(a, b, size=11):
    ch = a.size(-3)
    w = _window(size, ch).cuda(a.get_device()).type_as(a)       # H2D copy every call, as external.py:109
    conv = lambda t: F.conv2d(t, w, padding=size // 2, groups=ch)
    mu1, mu2 = conv(a), conv(b)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1, s2, s12 = conv(a * a) - mu1_sq, conv(b * b) - mu2_sq, conv(a * b) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))).mean()


def _quat_mult(q1, q2):
    w1, x1, y1, z1 = q1.T
    w2, x2, y2, z2 = q2.T
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2]).T


def _build_rotation(q):
    n = torch.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    q = q / n[:, None]
    rot = torch.zeros((q.size(0), 3, 3), device='cuda')
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot[:, 0, 0] = 1 - 2 * (y * y + z * z); rot[:, 0, 1] = 2 * (x * y - r * z); rot[:, 0, 2] = 2 * (x * z + r * y)   # 9 strided writes
    rot[:, 1, 0] = 2 * (x * y + r * z); rot[:, 1, 1] = 1 - 2 * (x * x + z * z); rot[:, 1, 2] = 2 * (y * z - r * x)
    rot[:, 2, 0] = 2 * (x * z - r * y); rot[:, 2, 1] = 2 * (y * z + r * x); rot[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return rot


def get_loss(Renderer, params, curr_data, variables, is_initial_timestep, weight_im=50.0, weight_seg=200.0, weight_rigid=200.0,
             weight_bg=200.0, weight_iso=1000.0, weight_rot=4.0):
    losses = {}
    rendervar = params2rendervar(params)
    rendervar['means2D'].retain_grad()
    im, radius, _ = Renderer(raster_settings=curr_data['cam'])(**rendervar)
    cid = curr_data['id']
    im = torch.exp(params['cam_m'][cid])[:, None, None] * im + params['cam_c'][cid][:, None, None]
    losses['im'] = 0.8 * torch.abs(im - curr_data['im']).mean() + 0.2 * (1.0 - calc_ssim(im, curr_data['im']))
    variables['means2D'] = rendervar['means2D']
    segrendervar = params2rendervar(params)
    segrendervar['colors_precomp'] = params['seg_colors']
    seg, _, _ = Renderer(raster_settings=curr_data['cam'])(**segrendervar)
    losses['seg'] = 0.8 * torch.abs(seg - curr_data['seg']).mean() + 0.2 * (1.0 - calc_ssim(seg, curr_data['seg']))
    if not is_initial_timestep:
        is_fg = (params['seg_colors'][:, 0] > 0.5).detach()
        fg_pts = rendervar['means3D'][is_fg]
        fg_rot = rendervar['rotations'][is_fg]
        rel_rot = _quat_mult(fg_rot, variables["prev_inv_rot_fg"])
        rot = _build_rotation(rel_rot)
        neighbor_pts = fg_pts[variables["neighbor_indices"]]
        curr_offset = neighbor_pts - fg_pts[:, None]
        off_prev = (rot.transpose(2, 1)[:, None] @ curr_offset[:, :, :, None]).squeeze(-1)
        w = variables["neighbor_weight"]
        losses['rigid'] = torch.sqrt(((off_prev - variables["prev_offset"]) ** 2).sum(-1) * w + 1e-20).mean()
        losses['rot'] = torch.sqrt(((rel_rot[variables["neighbor_indices"]] - rel_rot[:, None]) ** 2).sum(-1) * w + 1e-20).mean()
        mag = torch.sqrt((curr_offset ** 2).sum(-1) + 1e-20)
        losses['iso'] = torch.sqrt(((mag - variables["neighbor_dist"]) ** 2) * w + 1e-20).mean()
        losses['floor'] = torch.clamp(fg_pts[:, 1], min=0).mean()
        bg_pts = rendervar['means3D'][~is_fg]
        bg_rot = rendervar['rotations'][~is_fg]
        if bg_pts.shape[0] > 0:   # the benchmark scene has no background points: mean() of an empty tensor would be NaN
            losses['bg'] = torch.abs(bg_pts - variables["init_bg_pts"]).sum(-1).mean() + torch.abs(bg_rot - variables["init_bg_rot"]).sum(-1).mean()
    wts = {'im': weight_im, 'seg': weight_seg, 'rigid': weight_rigid, 'iso': weight_iso, 'rot': weight_rot, 'floor': 2.0, 'bg': weight_bg}
    loss = sum([wts[k] * v for k, v in losses.items()])
    seen = radius > 0
    variables['max_2D_radius'][seen] = torch.max(radius[seen].float(), variables['max_2D_radius'][seen])
    variables['seen'] = seen
    return loss, variables


def make_optimizer(params, scene_radius):
    """initialize_optimizer (train_utils.py:152-164) with the learning rates of the steady state (after
    initialize_post_first_timestep, train_utils.py:370-373)."""
    lrs = {'means3D': 0.00016 * scene_radius, 'rgb_colors': 0.0, 'seg_colors': 0.0, 'unnorm_rotations': 0.001,
           'logit_opacities': 0.0, 'log_scales': 0.0, 'cam_m': 0.0, 'cam_c': 0.0}
    return torch.optim.Adam([{'params': [v], 'name': k, 'lr': lrs[k]} for k, v in params.items()], lr=0.0, eps=1e-15)


def iteration(Renderer, params, data, variables, optimizer):
    loss, variables = get_loss(Renderer, params, data, variables, False)
    loss.backward()
    with torch.no_grad():
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
    return loss.detach()
