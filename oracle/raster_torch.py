"""
oracle/raster_torch.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Dense float64 (or float32) PyTorch restatement of the 3D-Gaussian rasterizer-with-depth that the
reference drives through ``GaussianRasterizer(raster_settings)(means3D, means2D, opacities,
colors_precomp, scales, rotations)`` (/root/reference/src/tracking/train_utils.py:178,192;
camera conventions /root/reference/src/tracking/helpers.py:10-33).

It is written independently of oracle/raster_oracle.c: every pixel is blended against every
Gaussian as dense [P, G] tensors and all gradients come from autograd, so it checks the hand-derived
backward of the C restatement (and of the CUDA kernels).  The rasterizer source
(JonathonLuiten/diff-gaussian-rasterization-w-depth, unpinned, /root/reference/README.md:26-35)
is not under /root/reference; its algorithm is restated from SURVEY.md §2.1.  PARITY UNPINNED.

Conventions reproduced so autograd matches the upstream hand-written backward:
  * alpha = min(0.99, o*G) passes the gradient straight through the clamp;
  * the 1.3*tanfov clamp of the view-space x/y used in the Jacobian has zero gradient when active and
    no dependence on t_z;
  * radii / tile rectangles / "done" thresholding / depth output are non-differentiable.

Only tests/ may import this module.  Sizes must stay small (P*G dense).
"""
import torch

TILE = 16


def quat_to_rot(q):
    r, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)
    return R.reshape(q.shape[:-1] + (3, 3))


def rasterize_dense(means3D, means2D, opacities, colors, scales, rotations, viewmatrix, projmatrix,
                    tanfovx, tanfovy, H, W, bg, scale_modifier=1.0, dtype=torch.float64):
    """Returns dict(color[CH,H,W], depth[1,H,W], radii[G], final_T[H,W], n_contrib[H,W], R)."""
    f = lambda t: t.to(dtype)
    means3D, opac, colors, scales, rot = f(means3D), f(opacities).reshape(-1), f(colors), f(scales), f(rotations)
    V = f(viewmatrix).reshape(4, 4)   # = w2c^T
    Pm = f(projmatrix).reshape(4, 4)  # = (P w2c)^T
    bg = f(bg)
    G = means3D.shape[0]
    ones = torch.ones(G, 1, dtype=dtype)
    hom = torch.cat([means3D, ones], 1)
    p_view = hom @ V[:, :3]                     # [G,3]
    p_hom = hom @ Pm                            # [G,4]
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    if means2D is not None:
        ndc = ndc + f(means2D)[:, :2]           # grad sink, numerically zero
    valid = p_view[:, 2] > 0.2

    # 3D covariance
    R = quat_to_rot(rot)
    s = scale_modifier * scales
    Sigma = R @ torch.diag_embed(s * s) @ R.transpose(1, 2)

    # EWA projection
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    tz = p_view[:, 2]
    tz_safe = torch.where(valid, tz, torch.ones_like(tz))
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = p_view[:, 0] / tz_safe, p_view[:, 1] / tz_safe
    xin = (txtz >= -limx) & (txtz <= limx)
    yin = (tytz >= -limy) & (tytz <= limy)
    tx = torch.where(xin, p_view[:, 0], (txtz.clamp(-limx, limx) * tz_safe).detach())
    ty = torch.where(yin, p_view[:, 1], (tytz.clamp(-limy, limy) * tz_safe).detach())
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz_safe, zero, -(fx * tx) / (tz_safe * tz_safe),
                     zero, fy / tz_safe, -(fy * ty) / (tz_safe * tz_safe)], -1).reshape(G, 2, 3)
    Rw = V[:3, :3].t()                           # w2c rotation
    M = J @ Rw
    cov2 = M @ Sigma @ M.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    valid = valid & (det != 0)
    det_safe = torch.where(valid, det, torch.ones_like(det))
    conA, conB, conC = c / det_safe, -b / det_safe, a / det_safe
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    pxd, pyd = px.detach(), py.detach()
    minx = torch.trunc((pxd - radius) / TILE).clamp(0, gx)
    miny = torch.trunc((pyd - radius) / TILE).clamp(0, gy)
    maxx = torch.trunc((pxd + radius + TILE - 1) / TILE).clamp(0, gx)
    maxy = torch.trunc((pyd + radius + TILE - 1) / TILE).clamp(0, gy)
    tiles = (maxx - minx) * (maxy - miny)
    valid = valid & (tiles > 0)
    radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)

    # order: depth bits of the fp32 depth (what the reference sorts on), ties by index (stable)
    depth32 = p_view[:, 2].detach().to(torch.float32)
    order = torch.argsort(depth32, stable=True)
    order = order[valid[order]]
    n = order.numel()

    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    pixx, pixy = xs.reshape(-1), ys.reshape(-1)            # [P]
    P = pixx.numel()
    if n == 0:
        col = bg[:, None].expand(-1, P).reshape(-1, H, W).clone()
        return dict(color=col, depth=torch.zeros(1, H, W, dtype=dtype), radii=radii,
                    final_T=torch.ones(H, W, dtype=dtype), n_contrib=torch.zeros(H, W, dtype=torch.int32), R=0)
    o = order
    dx = px[o][None, :] - pixx[:, None]                    # [P,n]
    dy = py[o][None, :] - pixy[:, None]
    power = -0.5 * (conA[o][None] * dx * dx + conC[o][None] * dy * dy) - conB[o][None] * dx * dy
    Gv = torch.exp(torch.clamp(power, max=0.0))
    a_raw = opac[o][None] * Gv
    alpha = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()   # straight-through clamp
    tpx = torch.div(pixx, TILE, rounding_mode="floor")
    tpy = torch.div(pixy, TILE, rounding_mode="floor")
    in_rect = (tpx[:, None] >= minx[o][None]) & (tpx[:, None] < maxx[o][None]) & \
              (tpy[:, None] >= miny[o][None]) & (tpy[:, None] < maxy[o][None])
    keep = in_rect & (power <= 0) & (alpha.detach() >= 1.0 / 255.0)
    alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
    one_m = 1.0 - alpha
    T_incl = torch.cumprod(one_m, 1)
    T_excl = torch.cat([torch.ones(P, 1, dtype=dtype), T_incl[:, :-1]], 1)
    # "done": first kept Gaussian whose inclusion would push T below 1e-4 stops the pixel (excluded)
    stop = keep & (T_incl.detach() < 1e-4)
    stopped = torch.cumsum(stop.to(torch.int32), 1) > 0
    live = keep & ~stopped
    w = torch.where(live, alpha * T_excl, torch.zeros_like(alpha))   # [P,n]
    col = w @ colors[o]                                              # [P,CH]
    dep = w @ depth32[o].to(dtype)[:, None]
    final_T = torch.where(live, one_m, torch.ones_like(one_m)).prod(1)
    color = (col + final_T[:, None] * bg[None]).t().reshape(-1, H, W)
    # n_contrib: 1-based position, inside the pixel's own TILE list, of the last live Gaussian
    tile_member = in_rect
    pos_in_tile = torch.cumsum(tile_member.to(torch.int64), 1)
    last = torch.where(live, pos_in_tile, torch.zeros_like(pos_in_tile)).max(1).values
    return dict(color=color, depth=dep.reshape(1, H, W), radii=radii, final_T=final_T.reshape(H, W),
                n_contrib=last.reshape(H, W).to(torch.int32), R=int(tiles[valid].sum().item()))


def setup_camera_mats(w, h, k, w2c, near=0.01, far=100.0):
    """viewmatrix/projmatrix/tanfov exactly as /root/reference/src/tracking/helpers.py:10-33 builds them
    (CPU float32 tensors)."""
    fx, fy, cx, cy = k[0][0], k[1][1], k[0][2], k[1][2]
    w2c_t = torch.tensor(w2c, dtype=torch.float32)
    cam_center = torch.inverse(w2c_t)[:3, 3]
    view = w2c_t.unsqueeze(0).transpose(1, 2)
    opengl_proj = torch.tensor([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                                [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                                [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                                [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32).unsqueeze(0).transpose(1, 2)
    full_proj = view.bmm(opengl_proj)
    return dict(viewmatrix=view.contiguous(), projmatrix=full_proj.contiguous(), campos=cam_center,
                tanfovx=w / (2 * fx), tanfovy=h / (2 * fy), image_height=h, image_width=w)
