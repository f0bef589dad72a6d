"""Host side of the tracking iteration (path A): the reference's tracking/helpers.py + train_utils.py surface, backed
by the sm_100a kernels of libgsd_b200.so.

Mirrors (same names, argument meaning, results):
  setup_camera, params2rendervar            /root/reference/src/tracking/helpers.py:10-45
  l1_loss_v1, calc_ssim, calc_psnr          helpers.py:71-72, external.py:45-47,101-135
  get_loss                                  /root/reference/src/tracking/train_utils.py:167-246
  initialize_optimizer                      train_utils.py:152-164   (-> FusedAdam, one launch for all groups)
  initialize_per_timestep                   train_utils.py:331-351   (in place: storage stays valid for CUDA graphs)
  initialize_post_first_timestep            train_utils.py:354-374
Differences by design: no boolean-mask indexing / host syncs inside get_loss (index lists are built once), the RGB and
seg renders of one iteration run as ONE 6-channel rasterizer pass when the colour render's means2D gradient is not
needed (t > 0), and `TrackingStep` captures loss + backward + Adam in a CUDA graph per camera.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib, _nvtx
from . import rasterizer as R
from .rasterizer import GaussianRasterizationSettings as Camera
from .rasterizer import GaussianRasterizer as Renderer

FLOOR_WEIGHT = 2.0  # hard-coded in the reference (train_utils.py:237)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ----------------------------------------------------------------------------------------------------
# cameras / activations
# ----------------------------------------------------------------------------------------------------
def setup_camera(w, h, k, w2c, near=0.01, far=100, bg=(0, 0, 0), device="cuda"):
    fx, fy, cx, cy = k[0][0], k[1][1], k[0][2], k[1][2]
    w2c = torch.tensor(np.asarray(w2c), dtype=torch.float32, device=device)
    cam_center = torch.inverse(w2c)[:3, 3]
    w2c = w2c.unsqueeze(0).transpose(1, 2)
    opengl_proj = torch.tensor([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                                [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                                [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                                [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32, device=device).unsqueeze(0).transpose(1, 2)
    full_proj = w2c.bmm(opengl_proj)
    return Camera(image_height=h, image_width=w, tanfovx=w / (2 * fx), tanfovy=h / (2 * fy),
                  bg=torch.tensor(bg, dtype=torch.float32, device=device), scale_modifier=1.0,
                  viewmatrix=w2c.contiguous(), projmatrix=full_proj.contiguous(), sh_degree=0, campos=cam_center,
                  prefiltered=False)


def params2rendervar(params):
    return {
        'means3D': params['means3D'],
        'colors_precomp': params['rgb_colors'],
        'rotations': torch.nn.functional.normalize(params['unnorm_rotations']),
        'opacities': torch.sigmoid(params['logit_opacities']),
        'scales': torch.exp(params['log_scales']),
        'means2D': torch.zeros_like(params['means3D'], requires_grad=True) + 0,
    }


# ----------------------------------------------------------------------------------------------------
# photometric loss  (fused L1 + SSIM)
# ----------------------------------------------------------------------------------------------------
def _ph_desc(x, y, n_sets, w_l1, w_ssim, set_weight, ws, affine=None, y_stats=None):
    d = _lib.GsdPhotometric()
    d.C, d.H, d.W, d.n_sets = x.shape[0], x.shape[1], x.shape[2], n_sets
    d.x, d.y = x.data_ptr(), y.data_ptr()
    d.affine_log_scale = affine[0].data_ptr() if affine is not None else None
    d.affine_shift = affine[1].data_ptr() if affine is not None else None
    d.w_l1, d.w_ssim = float(w_l1), float(w_ssim)
    d.set_weight[0], d.set_weight[1] = float(set_weight[0]), float(set_weight[1])
    d.ws = ws.data_ptr()
    d.y_mu = y_stats[0].data_ptr() if y_stats is not None else None
    d.y_s22 = y_stats[1].data_ptr() if y_stats is not None else None
    return d


def target_stats(y):
    """Window statistics (conv(y), conv(y*y)) of a target image: constant while the target does not change."""
    y = R._f32c(y, "y")
    mu, s22 = torch.empty_like(y), torch.empty_like(y)
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().gsd_photometric_target_stats(y.shape[0], y.shape[1], y.shape[2], y.data_ptr(), mu.data_ptr(),
                                                           s22.data_ptr(), _stream()), "gsd_photometric_target_stats")
    return mu, s22


def _ph_workspace(x):
    nbytes = C.c_size_t()
    _lib.check(_lib.lib().gsd_photometric_workspace_bytes(x.shape[0], x.shape[1], x.shape[2], C.byref(nbytes)),
               "gsd_photometric_workspace_bytes")
    return torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, w_l1, w_ssim):
        x = R._f32c(x, "x")
        y = R._f32c(y, "y")
        if x.shape != y.shape or x.dim() != 3 or x.shape[0] > 6:
            raise ValueError("photometric loss expects two [C<=6,H,W] tensors")
        with torch.cuda.device(x.device):
            ws = _ph_workspace(x)
            out = torch.empty(4, dtype=torch.float32, device=x.device)
            d = _ph_desc(x, y, 1, w_l1, w_ssim, (1.0, 1.0), ws)
            _lib.check(_lib.lib().gsd_photometric_forward(C.byref(d), out.data_ptr(), _stream()), "gsd_photometric_forward")
        ctx.save_for_backward(x, y, ws)
        ctx.w = (w_l1, w_ssim)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g_loss, g_out):
        x, y, ws = ctx.saved_tensors
        g_loss = g_loss.contiguous().float()
        grad = torch.empty_like(x)
        with torch.cuda.device(x.device):
            d = _ph_desc(x, y, 1, ctx.w[0], ctx.w[1], (1.0, 1.0), ws)
            _lib.check(_lib.lib().gsd_photometric_backward(C.byref(d), g_loss.data_ptr(), grad.data_ptr(), _stream()),
                       "gsd_photometric_backward")
        return grad, None, None, None


def photometric_loss(x, y, w_l1=0.8, w_ssim=0.2, return_parts=False):
    """0.8*l1_loss_v1(x,y) + 0.2*(1-calc_ssim(x,y)) of train_utils.py:185,195 in two kernels."""
    loss, parts = _Photometric.apply(x, y, float(w_l1), float(w_ssim))
    return (loss, parts) if return_parts else loss


def l1_loss_v1(x, y):
    return photometric_loss(x, y, 1.0, 0.0)


def calc_ssim(img1, img2):
    return photometric_loss(img1, img2, 0.0, 1.0, return_parts=True)[1][2]


def calc_psnr(img1, img2):
    mse = ((img1 - img2) ** 2).view(img1.shape[0], -1).mean(1, keepdim=True)
    return 20 * torch.log10(1.0 / torch.sqrt(mse))


# ----------------------------------------------------------------------------------------------------
# physical priors (rigid / rot / iso / floor / bg)
# ----------------------------------------------------------------------------------------------------
def _track_priors_launch(means3D, rotations, variables, weights, unnormalized=False):
    """One launch of gsd_track_losses_fwd_bwd: returns (losses[6] = rigid, rot, iso, floor, bg, weighted total; dL/dmeans3D;
    dL/drotations).  Plain function: the autograd wrapper below and the fused step both call it (no shared state).
    unnormalized: `rotations` is params['unnorm_rotations']; the kernels normalise on the fly (gradient w.r.t. the normalised)."""
    lib = _lib.lib()
    x = R._f32c(means3D, "means3D")
    q = R._f32c(rotations, "rotations")
    G = x.shape[0]
    v = variables
    t = _lib.GsdTrackLosses()
    fg, bg = v.get("fg_index"), v.get("bg_index")
    Gf = int(v["prev_inv_rot_fg"].shape[0])
    K = int(v["neighbor_indices_i32"].shape[1]) if Gf > 0 else 0
    Gb = int(bg.shape[0]) if bg is not None else 0
    t.G, t.Gf, t.K, t.Gb = G, Gf, K, Gb
    t.rotations_unnormalized = 1 if unnormalized else 0
    ptr = lambda tt: tt.data_ptr() if tt is not None and tt.numel() > 0 else None
    t.means3D, t.rotations = x.data_ptr(), q.data_ptr()
    t.fg_index = ptr(fg)
    t.prev_inv_rot = ptr(v["prev_inv_rot_fg"])
    t.neighbor_indices, t.neighbor_weight = ptr(v["neighbor_indices_i32"]), ptr(v["neighbor_weight"])
    t.neighbor_dist, t.prev_offset = ptr(v["neighbor_dist"]), ptr(v["prev_offset"])
    t.in_ptr, t.in_edge = ptr(v["in_ptr"]), ptr(v["in_edge"])
    er = v.get("edge_records")
    if (er is not None and er.get("src") is v["prev_offset"] and er.get("ver") == v["prev_offset"]._version
            and er.get("src_inv") is v["prev_inv_rot_fg"] and er.get("ver_inv") == v["prev_inv_rot_fg"]._version
            and er["layout"]["src"] is v["neighbor_indices_i32"]):
        # packed tables: the foreground set re-numbered along a Morton curve (see _priors_layout)
        lay = er["layout"]
        t.edge_records = er["data"].data_ptr()
        t.fg_index, t.prev_inv_rot = ptr(lay["fg_index"]), ptr(lay["prev_inv"])
        t.in_ptr, t.in_edge = ptr(lay["in_ptr"]), ptr(lay["in_edge"])
    t.bg_index, t.init_bg_pts, t.init_bg_rot = ptr(bg), ptr(v.get("init_bg_pts")), ptr(v.get("init_bg_rot"))
    t.w_rigid, t.w_rot, t.w_iso, t.w_floor, t.w_bg = [float(w) for w in weights]
    nbytes = C.c_size_t()
    _lib.check(lib.gsd_track_losses_workspace_bytes(Gf, Gb, C.byref(nbytes)), "gsd_track_losses_workspace_bytes")
    with torch.cuda.device(x.device):
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        losses = torch.empty(6, dtype=torch.float32, device=x.device)
        gx = torch.empty_like(x)
        gq = torch.empty_like(q)
        t.ws, t.losses, t.grad_means3D, t.grad_rotations = ws.data_ptr(), losses.data_ptr(), gx.data_ptr(), gq.data_ptr()
        with _nvtx.range("gsd.track_priors"):
            _lib.check(lib.gsd_track_losses_fwd_bwd(C.byref(t), _stream()), "gsd_track_losses_fwd_bwd")
    return losses, gx, gq


class _TrackPriors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, rotations, variables, weights):
        losses, gx, gq = _track_priors_launch(means3D, rotations, variables, weights)
        ctx.save_for_backward(gx, gq)
        ctx.mark_non_differentiable(losses)
        return losses[5], losses

    @staticmethod
    def backward(ctx, g_total, g_losses):
        gx, gq = ctx.saved_tensors
        return gx * g_total, gq * g_total, None, None


def _morton_order(pts):
    """Permutation that sorts points along a 30-bit Morton curve of their bounding box (10 bits per axis)."""
    lo, hi = pts.min(0).values, pts.max(0).values
    q = ((pts - lo) / (hi - lo).clamp_min(1e-12) * 1023.0).long().clamp_(0, 1023)

    def spread(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return torch.argsort(code, stable=True)


def _priors_layout(v, positions):
    """Static part of the packed priors tables (the kNN graph never changes after frame 0, train_utils.py:354-368): the foreground
    points re-numbered along a Morton curve.  The priors kernel is bound by the gathers of its neighbours' records; in curve
    order a warp's points share most of their neighbours, so those gathers hit L1 / the same L2 sectors (the Gaussians
    themselves keep the caller's order: the kernel reads and writes them through fg_index)."""
    nbr = v["neighbor_indices_i32"]
    lay = v.get("priors_layout")
    if lay is not None and lay["src"] is nbr and lay["ver"] == nbr._version:
        return lay
    Gf, K = nbr.shape
    dev = nbr.device
    fg = v.get("fg_index")
    fg_l = torch.arange(Gf, device=dev) if fg is None else fg.long()
    if positions is None:
        positions = v.get("prev_pts")
    if Gf > 0 and positions is not None and torch.is_tensor(positions) and positions.is_cuda and positions.shape[0] > int(fg_l.max()):
        perm = _morton_order(positions.detach()[fg_l].float())
    else:
        perm = torch.arange(Gf, device=dev)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(Gf, device=dev)
    nbr_p = inv[nbr.long()[perm]].to(torch.int32).contiguous()          # row f' = old row perm[f'], entries renamed
    in_ptr, in_edge = build_in_edges(nbr_p)
    lay = dict(src=nbr, ver=nbr._version, perm=perm, nbr=nbr_p, w=v["neighbor_weight"][perm].contiguous(),
               d=v["neighbor_dist"][perm].contiguous(), fg_index=fg_l[perm].to(torch.int32).contiguous(), in_ptr=in_ptr, in_edge=in_edge,
               prev_offset=torch.empty((Gf, K, 3), dtype=torch.float32, device=dev),
               prev_inv=torch.empty((Gf, 4), dtype=torch.float32, device=dev),
               data=torch.empty((Gf * K, 8), dtype=torch.float32, device=dev))
    v["priors_layout"] = lay
    return lay


def pack_edge_records(variables, positions=None):
    """Packs the per-edge tables into 32-byte records for the priors kernel, in the Morton order of `positions` ([G,3] means;
    default variables['prev_pts']; neither: caller's order).  Call after prev_offset / prev_inv_rot_fg changed (once per
    timestep); a stale pack is detected (tensor identity + version) and ignored.  All buffers are reused across calls (a captured
    CUDA graph keeps reading them on the following frames)."""
    v = variables
    Gf, K = v["neighbor_indices_i32"].shape
    lay = _priors_layout(v, positions)
    torch.index_select(v["prev_offset"], 0, lay["perm"], out=lay["prev_offset"])
    torch.index_select(v["prev_inv_rot_fg"], 0, lay["perm"], out=lay["prev_inv"])
    out = lay["data"]
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().gsd_track_pack_edges(Gf, K, lay["nbr"].data_ptr(), lay["w"].data_ptr(), lay["d"].data_ptr(),
                                                   lay["prev_offset"].data_ptr(), out.data_ptr(), _stream()), "gsd_track_pack_edges")
    v["edge_records"] = dict(data=out, src=v["prev_offset"], ver=v["prev_offset"]._version, src_inv=v["prev_inv_rot_fg"],
                             ver_inv=v["prev_inv_rot_fg"]._version, layout=lay)
    return v


def track_prior_losses(means3D, rotations, variables, weight_rigid, weight_rot, weight_iso, weight_bg,
                       weight_floor=FLOOR_WEIGHT):
    """Returns (weighted total, [rigid, rot, iso, floor, bg, total]) — train_utils.py:198-240."""
    return _TrackPriors.apply(means3D, rotations, variables, (weight_rigid, weight_rot, weight_iso, weight_floor, weight_bg))


# ----------------------------------------------------------------------------------------------------
# fused two-set rasterization (RGB + seg in one pass)
# ----------------------------------------------------------------------------------------------------
class _RasterizeTwoSets(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, colors0, colors1, scales, rotations, settings, bg1, capacity, sticky):
        color, radii, depth, state = R.raster_forward(settings, means3D, opacities, colors0, scales, rotations,
                                                       colors1=colors1, bg1=bg1, capacity=capacity, sticky=sticky)
        ctx.state = state
        ctx.opac_shape = opacities.shape
        ctx.need_m2d = means2D is not None and means2D.requires_grad
        ctx.mark_non_differentiable(radii, depth)
        ctx.status = state.status
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth):
        g = R.raster_backward(ctx.state, grad_color, need_means2D=ctx.need_m2d)
        ctx.state = None
        return (g["means3D"], g["means2D"], g["opacities"].reshape(ctx.opac_shape), g["colors0"], g["colors1"], g["scales"],
                g["rotations"], None, None, None, None)


def render_two_sets(settings, rendervar, colors1, bg1=None, capacity=None, sticky=None):
    """One 6-channel pass: channels 0-2 use rendervar['colors_precomp'], 3-5 use colors1 (same geometry)."""
    return _RasterizeTwoSets.apply(rendervar['means3D'], rendervar.get('means2D'), rendervar['opacities'],
                                   rendervar['colors_precomp'], colors1, rendervar['scales'], rendervar['rotations'],
                                   settings, bg1, capacity, sticky)


# ----------------------------------------------------------------------------------------------------
# get_loss
# ----------------------------------------------------------------------------------------------------
def update_seen(radius, variables):
    G = radius.shape[0]
    seen = torch.empty(G, dtype=torch.uint8, device=radius.device)
    with torch.cuda.device(radius.device):
        _lib.check(_lib.lib().gsd_track_update_radii(G, radius.data_ptr(), variables['max_2D_radius'].data_ptr(),
                                                     seen.data_ptr(), _stream()), "gsd_track_update_radii")
    variables['seen'] = seen.view(torch.bool)
    return variables


def get_loss(params, curr_data, variables, is_initial_timestep, weight_soft_col_cons=0.01, weight_im=50.0,
             weight_seg=200.0, weight_rigid=200.0, weight_bg=200.0, weight_iso=1000.0, weight_rot=4.0, fused=None,
             capacity=None, sticky=None):
    """Same contract as the reference's get_loss (train_utils.py:167-246): returns (loss, variables) and updates
    variables['means2D' | 'max_2D_radius' | 'seen'].  `fused` (default: t > 0) renders RGB+seg in one pass; in that mode
    variables['means2D'].grad holds the gradient of BOTH renders (only the t = 0 densifier reads it, so t = 0 defaults to
    the reference's two separate passes)."""
    losses = {}
    if fused is None:
        fused = not is_initial_timestep
    rendervar = params2rendervar(params)
    rendervar['means2D'].retain_grad()
    curr_id = curr_data['id']
    cam = curr_data['cam']
    if fused:
        out, radius, _ = render_two_sets(cam, rendervar, params['seg_colors'], capacity=capacity, sticky=sticky)
        im, seg = out[:3], out[3:]
    else:
        im, radius, _ = Renderer(raster_settings=cam)(**rendervar)
    im = torch.exp(params['cam_m'][curr_id])[:, None, None] * im + params['cam_c'][curr_id][:, None, None]
    losses['im'] = photometric_loss(im, curr_data['im'])
    variables['means2D'] = rendervar['means2D']
    if not fused:
        segrendervar = params2rendervar(params)
        segrendervar['colors_precomp'] = params['seg_colors']
        seg, _, _ = Renderer(raster_settings=cam)(**segrendervar)
    losses['seg'] = photometric_loss(seg, curr_data['seg'])
    loss = weight_im * losses['im'] + weight_seg * losses['seg']
    if not is_initial_timestep:
        prior, parts = track_prior_losses(rendervar['means3D'], rendervar['rotations'], variables, weight_rigid, weight_rot,
                                          weight_iso, weight_bg)
        variables['prior_losses'] = parts  # rigid, rot, iso, floor, bg, weighted total
        loss = loss + prior  # soft_col_cons is identically 0.0 in the reference (train_utils.py:230-232)
    variables = update_seen(radius, variables)
    return loss, variables


# ----------------------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------------------
class FusedAdam:
    """torch.optim.Adam(param_groups, lr=0.0, eps=1e-15) of initialize_optimizer (train_utils.py:152-164) as one kernel
    launch over every group.  Exposes param_groups / state like torch optimizers so the reference's per-timestep state
    surgery translates directly; step counters live on the device."""

    def __init__(self, param_groups, lr=0.0, betas=(0.9, 0.999), eps=1e-15):
        self.param_groups = []
        for g in param_groups:
            g = dict(g)
            g.setdefault('lr', lr)
            self.param_groups.append(g)
        self.betas, self.eps = betas, eps
        self.state = {}
        for g in self.param_groups:
            for p in g['params']:
                self.state[p] = dict(step=torch.zeros((), dtype=torch.float32, device=p.device),
                                     exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
        if sum(len(g['params']) for g in self.param_groups) > _lib.GSD_ADAM_MAX_TENSORS:
            raise ValueError("too many tensors for one fused launch")

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g['params']:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        a = _lib.GsdAdam()
        a.beta1, a.beta2, a.eps = self.betas[0], self.betas[1], self.eps
        n = 0
        dev = None
        for g in self.param_groups:
            for p in g['params']:
                if p.grad is None or not p.requires_grad:
                    continue
                if g['lr'] == 0.0:
                    # lr 0: parameter frozen; moments would be updated by torch but never used again unless lr changes.
                    # Keep exact torch semantics (moments advance) so a later lr change behaves identically.
                    pass
                st = self.state[p]
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                a.param[n], a.grad[n] = p.data_ptr(), grad.data_ptr()
                a.exp_avg[n], a.exp_avg_sq[n], a.step[n] = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), st['step'].data_ptr()
                a.lr[n], a.numel[n] = float(g['lr']), p.numel()
                dev = p.device
                n += 1
        a.n_tensors = n
        if n:
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().gsd_adam_step(C.byref(a), _stream()), "gsd_adam_step")


def initialize_optimizer(params, variables):
    lrs = {'means3D': 0.00016 * variables['scene_radius'], 'rgb_colors': 0.0, 'seg_colors': 0.0, 'unnorm_rotations': 0.001,
           'logit_opacities': 0.05, 'log_scales': 0.001, 'cam_m': 1e-4, 'cam_c': 1e-4}
    param_groups = [{'params': [v], 'name': k, 'lr': lrs[k]} for k, v in params.items()]
    return FusedAdam(param_groups, lr=0.0, eps=1e-15)


# ----------------------------------------------------------------------------------------------------
# per-timestep state
# ----------------------------------------------------------------------------------------------------
def knn(pts, num_knn, device="cuda"):
    """o3d_knn of the reference (helpers.py:97-115): squared distances + indices of the num_knn nearest OTHER points.
    The reference walks an Open3D KD-tree on the CPU with one Python iteration per point; here gsd_knn does the exact search
    on the device in float64.  numpy in -> (float64 [n,k], int64 [n,k]) numpy out; CUDA tensor in -> CUDA tensors out."""
    is_np = not torch.is_tensor(pts)
    p = torch.as_tensor(np.ascontiguousarray(pts, np.float32), device=device) if is_np else pts.detach().float().contiguous()
    if not p.is_cuda:
        raise ValueError("knn needs a CUDA device (no CPU path)")
    n = p.shape[0]
    sq = torch.empty((n, num_knn), dtype=torch.float64, device=p.device)
    idx = torch.empty((n, num_knn), dtype=torch.int32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().gsd_knn(n, num_knn, p.data_ptr(), sq.data_ptr(), idx.data_ptr(), _stream()), "gsd_knn")
    if is_np:
        return sq.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    return sq, idx.long()


def build_in_edges(neighbor_indices):
    """Transposed adjacency of the static kNN graph: for every point the ids (i*K+k) of the edges that name it."""
    Gf, K = neighbor_indices.shape
    flat = neighbor_indices.reshape(-1).long()
    order = torch.argsort(flat, stable=True)
    counts = torch.bincount(flat, minlength=Gf)
    in_ptr = torch.zeros(Gf + 1, dtype=torch.int32, device=flat.device)
    in_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return in_ptr.contiguous(), order.to(torch.int32).contiguous()


def initialize_post_first_timestep(params, variables, optimizer, num_knn=20):
    is_fg = params['seg_colors'][:, 0] > 0.5
    fg_index = torch.nonzero(is_fg).reshape(-1)
    bg_index = torch.nonzero(~is_fg).reshape(-1)
    init_fg_pts = params['means3D'][fg_index]
    neighbor_sq_dist, neighbor_indices = knn(init_fg_pts, num_knn)         # on the device, float64 like Open3D
    variables["neighbor_indices"] = neighbor_indices.contiguous()
    variables["neighbor_indices_i32"] = variables["neighbor_indices"].to(torch.int32).contiguous()
    variables["neighbor_weight"] = torch.exp(-2000 * neighbor_sq_dist).float().contiguous()
    variables["neighbor_dist"] = torch.sqrt(neighbor_sq_dist).float().contiguous()
    variables["in_ptr"], variables["in_edge"] = build_in_edges(variables["neighbor_indices_i32"])
    all_fg = bool(fg_index.numel() == is_fg.numel())
    variables["fg_index"] = None if all_fg else fg_index.to(torch.int32).contiguous()
    variables["bg_index"] = bg_index.to(torch.int32).contiguous()
    variables["init_bg_pts"] = params['means3D'][bg_index].detach().clone()
    variables["init_bg_rot"] = torch.nn.functional.normalize(params['unnorm_rotations'][bg_index]).detach().clone()
    variables["prev_pts"] = params['means3D'].detach().clone()
    variables["prev_rot"] = torch.nn.functional.normalize(params['unnorm_rotations']).detach().clone()
    for group in optimizer.param_groups:
        if group["name"] in ['logit_opacities', 'log_scales', 'cam_m', 'cam_c', 'rgb_colors']:
            group['lr'] = 0.0
    return variables


@torch.no_grad()
def initialize_per_timestep(params, variables, optimizer):
    """Constant-velocity warm start + Adam moment reset (train_utils.py:331-351), done IN PLACE."""
    pts = params['means3D']
    rot = torch.nn.functional.normalize(params['unnorm_rotations'])
    new_pts = pts + (pts - variables["prev_pts"])
    new_rot = torch.nn.functional.normalize(rot + (rot - variables["prev_rot"]))
    fg = variables.get("fg_index")
    rot_fg = rot if fg is None else rot[fg.long()]
    pts_fg = pts if fg is None else pts[fg.long()]
    prev_inv = rot_fg.clone()
    prev_inv[:, 1:] = -1 * prev_inv[:, 1:]
    prev_offset = pts_fg[variables["neighbor_indices"]] - pts_fg[:, None]

    def assign(key, value):
        if key in variables and isinstance(variables[key], torch.Tensor) and variables[key].shape == value.shape:
            variables[key].copy_(value)
        else:
            variables[key] = value.detach().clone().contiguous()
    assign('prev_inv_rot_fg', prev_inv)
    assign('prev_offset', prev_offset)
    assign('prev_col', params['rgb_colors'])
    assign('prev_pts', pts)
    assign('prev_rot', rot)
    pack_edge_records(variables)
    for k, v in (('means3D', new_pts), ('unnorm_rotations', new_rot)):
        p = params[k]
        p.data.copy_(v)
        st = optimizer.state[p]
        st['exp_avg'].zero_()
        st['exp_avg_sq'].zero_()
    return params, variables


# ----------------------------------------------------------------------------------------------------
# CUDA-graph fast path: one replay = get_loss + backward + Adam for one camera
# ----------------------------------------------------------------------------------------------------
class CapacityOverflow(RuntimeError):
    """Raised when a replayed iteration produced more tile instances than the captured buffers hold."""


class TrackingStep:
    """Captures the steady-state (t > 0) iteration of train_gs.py:25-39 per camera. Parameters, Adam state, neighbour
    tables and targets are static device tensors; `step(cam_id)` replays the graph and returns the device loss scalar.

    The rasterizer buffers of a captured graph have a fixed instance capacity (probe x capacity_margin) where the reference
    sizes them exactly from num_rendered on every call (train_utils.py:178).  Every forward call therefore records its
    instance count in `self.sticky` (device int32[2]: high-water R, number of overflowed calls); `check_capacity()` reads it
    with ONE 8-byte copy — call it once per frame.  train_frames / io.train redo a frame whose replays overflowed."""

    def __init__(self, params, variables, optimizer, dataset, is_initial_timestep=False, loss_kwargs=None,
                 capacity_margin=1.5, use_graph=True):
        self.params, self.variables, self.optimizer, self.dataset = params, variables, optimizer, dataset
        if is_initial_timestep and use_graph:
            # the first-frame loss renders RGB and seg separately through the autograd rasterizer, which reads the instance
            # count back to size its buffers (the one sync upstream also performs): not capturable, and densification
            # changes G between iterations anyway
            raise ValueError("the first-frame (t = 0) iteration cannot be captured in a CUDA graph: pass use_graph=False")
        self.is_initial = is_initial_timestep
        self.kw = dict(loss_kwargs or {})
        self.margin = capacity_margin
        self.use_graph = use_graph
        self.graphs, self.losses, self.capacity = {}, {}, {}
        self.sticky = torch.zeros(2, dtype=torch.int32, device=params['means3D'].device) if 'means3D' in params else None

    def _iteration(self, data, capacity):
        loss, self.variables = get_loss(self.params, data, self.variables, self.is_initial, capacity=capacity,
                                        sticky=self.sticky if capacity is not None else None, **self.kw)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _probe_capacity(self, data):
        rv = params2rendervar(self.params)
        with torch.no_grad():
            _, _, _, st = R.raster_forward(data['cam'], rv['means3D'], rv['opacities'], rv['colors_precomp'], rv['scales'],
                                           rv['rotations'])
        return max(1024, int(int(st.status[0].item()) * self.margin))

    # ---- optimisation state that an iteration modifies: saved / restored IN PLACE (captured pointers stay valid)
    def snapshot(self):
        snap = {'max_2D_radius': self.variables['max_2D_radius'].clone(), 'opt': []}
        for g in self.optimizer.param_groups:
            for p in g['params']:
                st = self.optimizer.state[p]
                snap['opt'].append((p, p.detach().clone(), st['exp_avg'].clone(), st['exp_avg_sq'].clone(), st['step'].clone()))
        return snap

    @torch.no_grad()
    def restore(self, snap):
        self.variables['max_2D_radius'].copy_(snap['max_2D_radius'])
        for p, val, m, v, step in snap['opt']:
            st = self.optimizer.state[p]
            p.data.copy_(val)
            st['exp_avg'].copy_(m); st['exp_avg_sq'].copy_(v); st['step'].copy_(step)

    def prepare(self, cam_ids=None):
        """Capacity probe, one eager warm-up iteration per camera on a side stream, capture.  The warm-up iterations are undone
        (parameters, Adam moments, step counters, max_2D_radius restored), so a frame is exactly the iterations the caller
        asks for (train_gs.py:25: 2 000), not 2 000 + n_cams."""
        cam_ids = list(range(len(self.dataset))) if cam_ids is None else list(cam_ids)
        for c in cam_ids:
            self.capacity[c] = self._probe_capacity(self.dataset[c])
        if not self.use_graph:
            return
        snap = self.snapshot()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for c in cam_ids:
                self.optimizer.zero_grad(set_to_none=True)
                self._iteration(self.dataset[c], self.capacity[c])
        torch.cuda.current_stream().wait_stream(s)
        self.restore(snap)
        for c in cam_ids:
            g = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(g, stream=getattr(self, 'capture_stream', None)):
                self.losses[c] = self._iteration(self.dataset[c], self.capacity[c])
            self.graphs[c] = g
        self.sticky.zero_()

    def check_capacity(self, raise_on_overflow=True):
        """One 8-byte D2H read: (high-water instance count, overflowed forward calls) since the last call; clears both."""
        hw, n_over = [int(x) for x in self.sticky.tolist()]
        self.sticky.zero_()
        if n_over and raise_on_overflow:
            raise CapacityOverflow("%d rasterizer call(s) exceeded the captured capacity (max R = %d, capacities %s)"
                                   % (n_over, hw, sorted(self.capacity.values())))
        return hw, n_over

    def grow(self, factor=1.5):
        """Re-probe at the current state with a larger margin and re-capture (after an overflow)."""
        self.margin *= factor
        self.graphs, self.losses = {}, {}
        self.prepare(list(self.capacity.keys()) or None)

    def step(self, cam_id):
        with _nvtx.range("gsd.tracking_iteration"):
            if self.use_graph:
                self.graphs[cam_id].replay()
                return self.losses[cam_id]
            self.optimizer.zero_grad(set_to_none=True)
            return self._iteration(self.dataset[cam_id], self.capacity.get(cam_id))


class FusedTrackingStep(TrackingStep):
    """Steady-state (t > 0) iteration without any PyTorch op in the loop: after the first frame only means3D and
    unnorm_rotations have a non-zero learning rate (train_utils.py:370-373), so opacities / scales / colours are constants
    and the whole iteration is

        render branch:  rasterize (RGB+seg in one pass; the preprocess kernel normalises the rotations) -> photometric (both sets,
                        affine colour fix fused) -> blend backward -> per-Gaussian backward with normalize-backward + gradient
                        sum + Adam applied in registers (the two live groups)
        side branch:    priors (rigid / rot / iso / floor / bg + gradients) | blend backward's prefix pass | scalar loss reduction

    = 17 launches of this library.  Same loss and parameter trajectory as get_loss + backward + FusedAdam.step (frozen groups
    are skipped: their values cannot change; their unused Adam moments are not advanced).  Requires every group other than
    means3D / unnorm_rotations to have lr == 0; use TrackingStep otherwise."""

    def __init__(self, params, variables, optimizer, dataset, loss_kwargs=None, capacity_margin=1.5, use_graph=True):
        super().__init__(params, variables, optimizer, dataset, False, loss_kwargs, capacity_margin, use_graph)
        live = {g['name'] for g in optimizer.param_groups if g['lr'] != 0.0}
        if not live <= {'means3D', 'unnorm_rotations'}:
            raise ValueError("FusedTrackingStep needs lr == 0 for every group except means3D / unnorm_rotations; live: %s" % live)
        kw = dict(weight_im=50.0, weight_seg=200.0, weight_rigid=200.0, weight_bg=200.0, weight_iso=1000.0, weight_rot=4.0)
        kw.update({k: v for k, v in self.kw.items() if k in kw})
        self.w = kw
        with torch.no_grad():
            self.opac = torch.sigmoid(params['logit_opacities']).reshape(-1).contiguous()
            self.scales = torch.exp(params['log_scales']).contiguous()
            self.rgb = params['rgb_colors'].detach().contiguous()
            self.seg = params['seg_colors'].detach().contiguous()
            self.targets = [torch.cat([d['im'], d['seg']], 0).contiguous() for d in dataset]
            self.rot = torch.empty_like(params['unnorm_rotations'])
            self.tstats = [target_stats(tg) for tg in self.targets]
        self.outputs = {}
        self._u8_stage = {}
        self.side = torch.cuda.Stream(device=params['means3D'].device)
        # the render branch is captured on a high-priority stream: when the side branch's CTAs fill the machine, the block scheduler
        # hands freed slots to the render branch first (tools/graph_timeline.py: preprocess starts 12 us earlier; -3 us per iteration)
        self.capture_stream = torch.cuda.Stream(device=params['means3D'].device, priority=-1)
        self.prefix_on_side = True   # blend backward's prefix pass on the side branch, beside the photometric kernels
        self.fuse_update = True      # rasterizer backward + update in one per-Gaussian kernel (False: gsd_raster_backward, gsd_track_update)
        self.priors_fork = 'start'   # where the priors branch forks off the render branch: 'start' | 'after_forward'
        self.block_counter = torch.zeros(1, dtype=torch.int32, device=params['means3D'].device)   # self-resetting (gsd_track_update)
        self.lr = {g['name']: float(g['lr']) for g in optimizer.param_groups}

    def set_targets(self, dataset):
        """Next frame of the episode: same cameras, new images (train_utils.py:10-29 loads them per frame)."""
        for c, d in enumerate(dataset):
            self.set_target(c, d['im'], d['seg'])

    def set_target(self, cam_id, im, seg):
        """New target images for a camera (e.g. the next frame's, from pinned host memory): copies them into the static
        buffers the captured graph reads and refreshes their window statistics."""
        tg = self.targets[cam_id]
        tg[:3].copy_(im, non_blocking=True)
        tg[3:].copy_(seg, non_blocking=True)
        mu, s22 = self.tstats[cam_id]
        with torch.cuda.device(tg.device):
            _lib.check(_lib.lib().gsd_photometric_target_stats(6, tg.shape[1], tg.shape[2], tg.data_ptr(), mu.data_ptr(),
                                                               s22.data_ptr(), _stream()), "gsd_photometric_target_stats")

    def set_target_u8(self, cam_id, im_hwc_u8, seg_u8):
        """Like set_target, from the dataset's bytes: im_hwc_u8 [H,W,3|4] uint8 (a PIL image as np.array gives it) and seg_u8 [H,W]
        uint8, e.g. pinned host tensors.  One sixth of set_target's host-to-device traffic; the float planes are built on the
        device exactly as train_utils.py:66-75 builds them on the host (im / 255; seg, 0, 1 - seg)."""
        tg = self.targets[cam_id]
        H, W = tg.shape[1], tg.shape[2]
        if im_hwc_u8.dtype != torch.uint8 or seg_u8.dtype != torch.uint8 or im_hwc_u8.dim() != 3 or im_hwc_u8.shape[:2] != (H, W) \
                or im_hwc_u8.shape[2] not in (3, 4) or seg_u8.shape != (H, W):
            raise ValueError("set_target_u8 expects uint8 im [H,W,3|4] and seg [H,W] of the camera's size")
        st = self._u8_stage.get(cam_id)
        if st is None or st[0].shape != im_hwc_u8.shape:
            st = (torch.empty(im_hwc_u8.shape, dtype=torch.uint8, device=tg.device), torch.empty((H, W), dtype=torch.uint8, device=tg.device))
            self._u8_stage[cam_id] = st
        st[0].copy_(im_hwc_u8, non_blocking=True)
        st[1].copy_(seg_u8, non_blocking=True)
        mu, s22 = self.tstats[cam_id]
        with torch.cuda.device(tg.device):
            _lib.check(_lib.lib().gsd_track_unpack_target_u8(H, W, im_hwc_u8.shape[2], st[0].data_ptr(), st[1].data_ptr(), tg.data_ptr(),
                                                             _stream()), "gsd_track_unpack_target_u8")
            _lib.check(_lib.lib().gsd_photometric_target_stats(6, H, W, tg.data_ptr(), mu.data_ptr(), s22.data_ptr(), _stream()),
                       "gsd_photometric_target_stats")

    @torch.no_grad()
    def _iteration(self, data, capacity):
        lib = _lib.lib()
        P, V = self.params, self.variables
        x, uq = P['means3D'], P['unnorm_rotations']
        G = x.shape[0]
        cid = data['id']
        idx = next((i for i, d_ in enumerate(self.dataset) if d_ is data), None)
        tgt = self.targets[idx] if idx is not None else torch.cat([data['im'], data['seg']], 0).contiguous()
        tst = self.tstats[idx] if idx is not None else None
        with torch.cuda.device(x.device):
            st = _stream()
            # F.normalize(unnorm_rotations) (params2rendervar) happens inside the consumers: the rasterizer's preprocess kernel
            # writes self.rot for the backward, the priors' kernels normalise on the fly — no launch in front of the fork
            # the physical priors depend only on the parameters, not on the render: they run on a side stream concurrently
            # with binning / sorting / blending (fork-join, captured as parallel branches of the CUDA graph)
            main = torch.cuda.current_stream()

            def launch_priors():
                fork = torch.cuda.Event()
                fork.record(main)
                with torch.cuda.stream(self.side):
                    self.side.wait_event(fork)
                    parts, gx_p, gq_p = _track_priors_launch(x.detach(), uq.detach(), V, (self.w['weight_rigid'], self.w['weight_rot'],
                                                             self.w['weight_iso'], FLOOR_WEIGHT, self.w['weight_bg']), unnormalized=True)
                    join = torch.cuda.Event()
                    join.record(self.side)
                for tns in (parts, gx_p, gq_p):
                    tns.record_stream(main)
                return parts, gx_p, gq_p, join

            if self.priors_fork == 'start':
                parts, gx_p, gq_p, join = launch_priors()
            color, radii, _, state = R.raster_forward(data['cam'], x.detach(), self.opac, self.rgb, self.scales, self.rot,
                                                      colors1=self.seg, capacity=capacity,
                                                      sticky=self.sticky if capacity is not None else None, unnorm_rotations=uq.detach())
            if self.priors_fork == 'after_forward':
                parts, gx_p, gq_p, join = launch_priors()
            # the blend backward's per-chunk prefix pass needs only the forward's state: side branch, beside the photometric kernels
            prefix_ev = None
            if self.prefix_on_side:
                fwd_done = torch.cuda.Event()
                fwd_done.record(main)
                with torch.cuda.stream(self.side):
                    self.side.wait_event(fwd_done)
                    R.raster_backward_prepare(state)
                    prefix_ev = torch.cuda.Event()
                    prefix_ev.record(self.side)
            prior = parts[5]
            ws = _ph_workspace(color)
            ph = torch.empty(8, dtype=torch.float32, device=x.device)
            d = _ph_desc(color, tgt, 2, 0.8, 0.2, (self.w['weight_im'], self.w['weight_seg']), ws,
                         affine=(P['cam_m'][cid], P['cam_c'][cid]), y_stats=tst)
            with _nvtx.range("gsd.photometric"):
                _lib.check(lib.gsd_photometric_stats(C.byref(d), st), "gsd_photometric_stats")
            # the scalar reduction (and the addition of the prior losses: ph[7] is the iteration's loss) is not needed by the
            # gradient pass: it joins the side branch instead of sitting on the critical path
            stats_done = torch.cuda.Event()
            stats_done.record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(stats_done)
                _lib.check(lib.gsd_photometric_reduce(C.byref(d), prior.data_ptr(), ph.data_ptr(), C.c_void_p(self.side.cuda_stream)),
                           "gsd_photometric_reduce")
                loss_done = torch.cuda.Event()
                loss_done.record(self.side)
            for tns in (ph, ws, color, tgt):
                tns.record_stream(self.side)
            dL = torch.empty_like(color)
            _lib.check(lib.gsd_photometric_backward(C.byref(d), None, dL.data_ptr(), st), "gsd_photometric_backward")
            main.wait_event(join)
            u = _lib.GsdTrackUpdate()
            u.G = G
            u.beta1, u.beta2, u.eps = self.optimizer.betas[0], self.optimizer.betas[1], self.optimizer.eps
            u.lr_means, u.lr_rot = self.lr['means3D'], self.lr['unnorm_rotations']
            sm, sr = self.optimizer.state[x], self.optimizer.state[uq]
            u.means3D, u.unnorm_rotations = x.data_ptr(), uq.data_ptr()
            u.g_means_b, u.g_rot_b = gx_p.data_ptr(), gq_p.data_ptr()
            u.m_means, u.v_means = sm['exp_avg'].data_ptr(), sm['exp_avg_sq'].data_ptr()
            u.m_rot, u.v_rot = sr['exp_avg'].data_ptr(), sr['exp_avg_sq'].data_ptr()
            u.step_means, u.step_rot = sm['step'].data_ptr(), sr['step'].data_ptr()
            # bookkeeping of get_loss (seen / max_2D_radius, train_utils.py:243-245) and the step advance ride in the same launch
            seen = torch.empty(G, dtype=torch.uint8, device=x.device)
            u.radii, u.max_2D_radius, u.seen = radii.data_ptr(), V['max_2D_radius'].data_ptr(), seen.data_ptr()
            u.block_counter = self.block_counter.data_ptr()
            if prefix_ev is not None:
                main.wait_event(prefix_ev)
            if self.fuse_update:
                # the rasterizer's per-Gaussian backward applies the update in registers: no gradient arrays, one launch less
                R.raster_backward(state, dL, fused_update=u, prefix_done=prefix_ev is not None)
            else:
                g = R.raster_backward(state, dL, need_means2D=False, geom_only=True, prefix_done=prefix_ev is not None)
                u.g_means_a, u.g_rot_a = g['means3D'].data_ptr(), g['rotations'].data_ptr()
                _lib.check(lib.gsd_track_update(C.byref(u), st), "gsd_track_update")
            main.wait_event(loss_done)
            # per-camera result buffers (every captured graph writes its own): step() re-points `variables` at the ones of the
            # camera it replayed
            self.outputs[idx if idx is not None else cid] = {'seen': seen.view(torch.bool), 'prior_losses': parts, 'photometric_losses': ph[:7]}
            V.update(self.outputs[idx if idx is not None else cid])
        return ph[7]

    def step(self, cam_id):
        loss = super().step(cam_id)
        if self.use_graph:
            self.variables.update(self.outputs[cam_id])
        return loss


# ----------------------------------------------------------------------------------------------------
# t = 0 path: densification statistics and parameter/optimizer surgery (A9, A12: /root/reference/src/tracking/external.py:138-299).
# Host-side PyTorch logic on top of FusedAdam's state; the kernels above do the per-iteration work.  Variable G is
# incompatible with CUDA-graph capture, so t = 0 runs eagerly through get_loss(..., fused=False).
# ----------------------------------------------------------------------------------------------------
PER_POINT_KEYS = ('means3D', 'rgb_colors', 'seg_colors', 'unnorm_rotations', 'logit_opacities', 'log_scales')


def _group(optimizer, name):
    for g in optimizer.param_groups:
        if g['name'] == name:
            return g
    raise KeyError(name)


def _replace_param(optimizer, params, name, value, exp_avg=None, exp_avg_sq=None, keep_step=True):
    """Swap the tensor of a parameter group for `value`, carrying over / installing its Adam moments."""
    g = _group(optimizer, name)
    old = g['params'][0]
    st = optimizer.state.pop(old, None)
    new = torch.nn.Parameter(value.detach().clone().contiguous(), requires_grad=old.requires_grad)
    g['params'][0] = new
    params[name] = new
    step = st['step'] if (st is not None and keep_step) else torch.zeros((), dtype=torch.float32, device=new.device)
    optimizer.state[new] = dict(step=step,
                                exp_avg=torch.zeros_like(new) if exp_avg is None else exp_avg.contiguous(),
                                exp_avg_sq=torch.zeros_like(new) if exp_avg_sq is None else exp_avg_sq.contiguous())
    return new


def update_params_and_optimizer(new_params, params, optimizer):
    """external.py:145-157: new values, Adam moments reset to zero (step counter kept)."""
    for k, v in new_params.items():
        _replace_param(optimizer, params, k, v)
    return params


def cat_params_to_optimizer(new_params, params, optimizer):
    """external.py:160-174: append rows; the appended rows start with zero Adam moments."""
    for k, v in new_params.items():
        st = optimizer.state[_group(optimizer, k)['params'][0]]
        _replace_param(optimizer, params, k, torch.cat((params[k].detach(), v), 0),
                       torch.cat((st['exp_avg'], torch.zeros_like(v)), 0), torch.cat((st['exp_avg_sq'], torch.zeros_like(v)), 0))
    return params


def remove_points(to_remove, params, variables, optimizer):
    """external.py:177-222: drop rows (parameters, Adam moments and the densification statistics)."""
    keep = ~to_remove
    for k in [k for k in params.keys() if k not in ('cam_m', 'cam_c')]:
        st = optimizer.state[_group(optimizer, k)['params'][0]]
        _replace_param(optimizer, params, k, params[k].detach()[keep], st['exp_avg'][keep], st['exp_avg_sq'][keep])
    for k in ('means2D_gradient_accum', 'denom', 'max_2D_radius'):
        variables[k] = variables[k][keep]
    return params, variables


def accumulate_mean2d_gradient(variables):
    """external.py:138-142 without boolean-mask indexing (no host sync)."""
    seen = variables['seen']
    gnorm = torch.norm(variables['means2D'].grad[:, :2], dim=-1)
    variables['means2D_gradient_accum'] += torch.where(seen, gnorm, torch.zeros_like(gnorm))
    variables['denom'] += seen.to(variables['denom'].dtype)
    return variables


def _install_param(optimizer, params, name, value, exp_avg, exp_avg_sq):
    """Swap a parameter group's tensor and Adam moments for freshly built ones (no copies); the step counter is kept."""
    g = _group(optimizer, name)
    old = g['params'][0]
    st = optimizer.state.pop(old, None)
    new = torch.nn.Parameter(value, requires_grad=old.requires_grad)
    g['params'][0] = new
    params[name] = new
    step = st['step'] if st is not None else torch.zeros((), dtype=torch.float32, device=new.device)
    optimizer.state[new] = dict(step=step, exp_avg=exp_avg, exp_avg_sq=exp_avg_sq)


@torch.no_grad()
def densify(params, variables, optimizer, i, remove_thresh, remove_thresh_5k, scale_scene_radius, grad_thresh=0.0002, normal_sampler=None):
    """Clone / split / prune schedule of external.py:229-299 (t = 0 only): every 100 iterations for 500 <= i <= 5000 points
    with a large screen-space gradient are cloned (small) or split in two (large, sampled from the Gaussian, scale / 1.6);
    transparent (and, after 3000, oversized) points are pruned; opacities are reset to 0.01 every 3000 iterations.
    One round = two kernels (gsd_densify_plan / gsd_densify_apply: classification, stream compaction of all 18 per-point arrays
    and the Adam-state surgery) and ONE 16-byte read of the new point count, where the reference runs ~100 eager kernels with
    ~10 host syncs.  normal_sampler(stds) -> samples replaces torch.normal for the split copies (tests: reproducible noise)."""
    if i > 5000:
        return params, variables, params['means3D'].shape[0]
    variables = accumulate_mean2d_gradient(variables)
    do_round = i >= 500 and i % 100 == 0
    do_reset = i > 0 and i % 3000 == 0
    if not (do_round or do_reset):
        return params, variables, params['means3D'].shape[0]
    lib = _lib.lib()
    n = params['means3D'].shape[0]
    dev = params['means3D'].device
    src = [params[k].detach().contiguous() for k in PER_POINT_KEYS]
    states = [optimizer.state[_group(optimizer, k)['params'][0]] for k in PER_POINT_KEYS]
    widths = [t.shape[1] for t in src]
    with torch.cuda.device(dev):
        dst = torch.empty(4 * max(n, 1), dtype=torch.int32, device=dev)
        totals = torch.empty(4, dtype=torch.int32, device=dev)
        pl = _lib.GsdDensifyPlan()
        pl.n, pl.do_densify = n, int(do_round)
        pl.grad_thresh = float(grad_thresh)
        pl.clone_limit = float(scale_scene_radius * variables['scene_radius'])
        pl.prune_opacity = float(remove_thresh_5k if i == 5000 else remove_thresh)
        pl.prune_big = float(0.1 * variables['scene_radius']) if i >= 3000 else 0.0
        accum, denom = variables['means2D_gradient_accum'].contiguous(), variables['denom'].contiguous()
        pl.grad_accum, pl.denom = accum.data_ptr(), denom.data_ptr()
        pl.log_scales, pl.logit_opacities = src[5].data_ptr(), src[4].data_ptr()
        pl.dst, pl.totals = dst.data_ptr(), totals.data_ptr()
        _lib.check(lib.gsd_densify_plan(C.byref(pl), _stream()), "gsd_densify_plan")
        k0, k1, k2, ns = [int(v) for v in totals.tolist()]          # the one host sync of the round
        n_out = k0 + k1 + 2 * k2
        scaled = 0
        if ns > 0 and normal_sampler is not None:
            split = dst[3 * n:4 * n] >= 0
            samples = normal_sampler(torch.exp(src[5])[split].repeat(2, 1)).contiguous()
            scaled = 1
        else:
            samples = torch.randn((max(2 * ns, 1), 3), dtype=torch.float32, device=dev)
        new_p = [torch.empty((n_out, w), dtype=torch.float32, device=dev) for w in widths]
        new_m = [torch.empty((n_out, w), dtype=torch.float32, device=dev) for w in widths]
        new_v = [torch.empty((n_out, w), dtype=torch.float32, device=dev) for w in widths]
        ap = _lib.GsdDensifyApply()
        ap.n, ap.samples_scaled, ap.reset_opacity = n, scaled, int(do_reset)
        ap.dst, ap.totals, ap.samples = dst.data_ptr(), totals.data_ptr(), samples.data_ptr()
        keep = []
        for t in range(6):
            m, v = states[t]['exp_avg'].contiguous(), states[t]['exp_avg_sq'].contiguous()
            keep += [m, v]
            ap.p_src[t], ap.m_src[t], ap.v_src[t] = src[t].data_ptr(), m.data_ptr(), v.data_ptr()
            ap.p_dst[t], ap.m_dst[t], ap.v_dst[t] = new_p[t].data_ptr(), new_m[t].data_ptr(), new_v[t].data_ptr()
            ap.width[t] = widths[t]
        _lib.check(lib.gsd_densify_apply(C.byref(ap), _stream()), "gsd_densify_apply")
    for t, k in enumerate(PER_POINT_KEYS):
        _install_param(optimizer, params, k, new_p[t], new_m[t], new_v[t])
    if do_round:   # the statistics restart after every round (external.py:275-277; pruning zeros stays zeros)
        for k in ('means2D_gradient_accum', 'denom', 'max_2D_radius'):
            variables[k] = torch.zeros(n_out, device=dev)
    return params, variables, n_out


def initialize_params_from_point_cloud(init_pt_cld, cam_centers, device="cuda", max_cams=50):
    """initialize_params of train_utils.py:89-149 from an in-memory [n,7] array (xyz, rgb, seg) instead of the .npz path."""
    pts = np.asarray(init_pt_cld, np.float64)
    seg = pts[:, 6]
    sq, _ = knn(pts[:, :3], 3)
    mean3 = sq.mean(-1).clip(min=1e-7)
    raw = {'means3D': pts[:, :3], 'rgb_colors': pts[:, 3:6], 'seg_colors': np.stack((seg, np.zeros_like(seg), 1 - seg), -1),
           'unnorm_rotations': np.tile([1, 0, 0, 0], (seg.shape[0], 1)), 'logit_opacities': np.zeros((seg.shape[0], 1)),
           'log_scales': np.tile(np.log(np.sqrt(mean3))[..., None], (1, 3)), 'cam_m': np.zeros((max_cams, 3)),
           'cam_c': np.zeros((max_cams, 3))}
    params = {k: torch.nn.Parameter(torch.tensor(v, dtype=torch.float32, device=device).contiguous()) for k, v in raw.items()}
    params['rgb_colors'].requires_grad = False
    cam_centers = np.asarray(cam_centers, np.float64)
    scene_radius = 1.1 * np.max(np.linalg.norm(cam_centers - np.mean(cam_centers, 0)[None], axis=-1))
    n = params['means3D'].shape[0]
    z = lambda: torch.zeros(n, device=device)
    variables = {'max_2D_radius': z(), 'scene_radius': float(scene_radius), 'means2D_gradient_accum': z(), 'denom': z()}
    return params, variables


def run_frame(step, cam_sequence, max_retries=3):
    """One tracked frame (train_gs.py:25-39): step.step(c) for every c of cam_sequence — exactly len(cam_sequence) optimiser
    iterations — then ONE capacity check.  If any replay overflowed the captured rasterizer buffers the frame is redone from
    its start state with larger buffers (re-probe + re-capture) instead of keeping iterations that dropped instances."""
    cam_sequence = list(cam_sequence)
    for attempt in range(max_retries + 1):
        snap = step.snapshot()
        for c in cam_sequence:
            step.step(c)
        hw, n_over = step.check_capacity(raise_on_overflow=False)
        if n_over == 0:
            return hw
        if attempt == max_retries:
            raise CapacityOverflow("frame still overflows after %d re-captures (max R = %d)" % (max_retries, hw))
        step.restore(snap)
        step.grow(max(1.5, 1.25 * hw / max(1, min(step.capacity.values()))))
    return hw


def train_frames(params, variables, optimizer, datasets, iters_first=10000, iters_next=2000, num_knn=20, densify_args=None,
                 loss_kwargs=None, seed=0):
    """The episode loop of train_gs.py:10-46 on in-memory per-frame datasets (lists of {'cam','im','seg','id'}): 10 000
    iterations with densification at t = 0, then ONE FusedTrackingStep whose per-camera CUDA graphs are captured at the first
    tracked frame and reused for every later frame (only the target images change; all other buffers are updated in place).
    Cameras are drawn like the reference's get_batch (train_utils.py:81-85: uniform WITH replacement — its refill rebinds a
    local, so the caller's todo list stays empty).  Returns per-frame snapshots of (means3D, rgb_colors, unnorm_rotations)
    like params2cpu (helpers.py:141-147)."""
    import random
    rnd = random.Random(seed)
    out = []
    step = None
    for t, dataset in enumerate(datasets):
        if t == 0:
            for i in range(iters_first):
                data = dataset[rnd.randint(0, len(dataset) - 1)]
                loss, variables = get_loss(params, data, variables, True, fused=False, **(loss_kwargs or {}))
                loss.backward()
                with torch.no_grad():
                    if densify_args is not None:
                        params, variables, _ = densify(params, variables, optimizer, i, *densify_args)
                    optimizer.step()
                    optimizer.zero_grad(set_to_none=True)
            variables = initialize_post_first_timestep(params, variables, optimizer, num_knn)
        else:
            params, variables = initialize_per_timestep(params, variables, optimizer)
            if step is None:
                step = FusedTrackingStep(params, variables, optimizer, dataset, loss_kwargs=loss_kwargs)
                step.prepare()
            else:
                step.set_targets(dataset)
            run_frame(step, [rnd.randint(0, len(dataset) - 1) for _ in range(iters_next)])
            variables = step.variables
        out.append({k: params[k].detach().cpu().numpy().copy() for k in ('means3D', 'rgb_colors', 'unnorm_rotations')})
    return params, variables, out
