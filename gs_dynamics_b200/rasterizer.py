"""Drop-in for ``diff_gaussian_rasterization`` (JonathonLuiten/diff-gaussian-rasterization-w-depth), backed by the
sm_100a kernels of libgsd_b200.so through the C ABI of include/gsd.h.

Surface kept (what the reference constructs / calls):
  GaussianRasterizationSettings  — 11 fields in upstream order, built by keyword at
                                   /root/reference/src/tracking/helpers.py:20-32, src/render/renderer.py:37-49
  GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs=None, colors_precomp=None,
                                      scales=None, rotations=None, cov3D_precomp=None) -> (color, radii, depth)
                                 — called at /root/reference/src/tracking/train_utils.py:178,192,379,
                                   src/render/renderer.py:22, src/real_world/gs/trainer.py:61, src/demo.py:198,206
  GaussianRasterizer.markVisible(positions)

Not supported (never used by the reference): ``shs`` / ``sh_degree > 0`` / ``cov3D_precomp``.
There is no CPU path: tensors must be CUDA float32; a missing extension raises.
"""
import ctypes as C
from typing import NamedTuple

import torch

from . import _lib, _nvtx


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool


_ws_cache = {}


def _workspace_bytes(G, W, H, n_sets, capacity):
    key = (G, W, H, n_sets, capacity)
    r = _ws_cache.get(key)
    if r is None:
        out = (C.c_size_t * 4)()
        _lib.check(_lib.lib().gsd_raster_workspace_bytes(G, W, H, n_sets, capacity, out), "gsd_raster_workspace_bytes")
        r = tuple(int(x) for x in out)
        if len(_ws_cache) > 256:
            _ws_cache.clear()
        _ws_cache[key] = r
    return r


def _f32c(t, name, shape_tail=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (no CPU path)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32" % name)
    return t.contiguous()


class RasterState:
    """Buffers of one forward call that the matching backward needs."""
    __slots__ = ("desc", "keep", "capacity", "G", "W", "H", "n_sets", "status")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def raster_forward(settings, means3D, opacities, colors0, scales, rotations, colors1=None, bg1=None, capacity=None, sticky=None,
                   unnorm_rotations=None):
    """Runs the forward kernels. Returns (color [3*n_sets,H,W], radii [G] int32, depth [1,H,W], state).

    capacity=None reproduces upstream behaviour: the instance count R is read back once (one D2H sync, what
    upstream's num_rendered copy does) and buffers are sized exactly.  Passing a capacity >= R makes the call
    fully asynchronous; state.status[1] is set on the device if it was too small, and `sticky` (int32[2] CUDA tensor, zeroed
    by the caller) accumulates max R / number of overflowed calls across calls (read it once per frame, not per call).
    unnorm_rotations ([G,4]): `rotations` is then an OUTPUT buffer — the preprocess kernel writes F.normalize(unnorm_rotations)
    into it and renders with it (params2rendervar without a separate normalisation launch)."""
    lib = _lib.lib()
    means3D = _f32c(means3D, "means3D")
    dev = means3D.device
    G = means3D.shape[0]
    opacities = _f32c(opacities, "opacities").reshape(-1)
    colors0 = _f32c(colors0, "colors_precomp")
    scales = _f32c(scales, "scales")
    rotations = _f32c(rotations, "rotations")
    if means3D.shape != (G, 3) or colors0.shape != (G, 3) or scales.shape != (G, 3) or rotations.shape != (G, 4) \
            or opacities.shape[0] != G:
        raise ValueError("inconsistent input shapes")
    n_sets = 1
    if colors1 is not None:
        colors1 = _f32c(colors1, "colors1")
        if colors1.shape != (G, 3):
            raise ValueError("colors1 must be [G,3]")
        n_sets = 2
    H, W = int(settings.image_height), int(settings.image_width)
    view = _f32c(settings.viewmatrix, "viewmatrix").reshape(-1)
    proj = _f32c(settings.projmatrix, "projmatrix").reshape(-1)
    bg0 = _f32c(settings.bg, "bg").reshape(-1)
    if view.numel() != 16 or proj.numel() != 16 or bg0.numel() != 3:
        raise ValueError("viewmatrix/projmatrix must have 16 elements and bg 3")
    if bg1 is not None:
        bg1 = _f32c(bg1, "bg1").reshape(-1)

    if unnorm_rotations is not None:
        unnorm_rotations = _f32c(unnorm_rotations, "unnorm_rotations")
        if unnorm_rotations.shape != (G, 4):
            raise ValueError("unnorm_rotations must be [G,4]")
    with torch.cuda.device(dev):
        st = _stream()
        status = torch.empty(_lib.GSD_STATUS_WORDS, dtype=torch.int32, device=dev)   # zeroed by the library (memset node)
        radii = torch.empty(G, dtype=torch.int32, device=dev)
        d = _lib.GsdRasterFwd()
        d.G, d.W, d.H, d.n_sets = G, W, H, n_sets
        d.tanfovx, d.tanfovy, d.scale_modifier = float(settings.tanfovx), float(settings.tanfovy), float(settings.scale_modifier)
        d.viewmatrix, d.projmatrix, d.bg0 = view.data_ptr(), proj.data_ptr(), bg0.data_ptr()
        d.bg1 = bg1.data_ptr() if bg1 is not None else None
        d.means3D, d.opacities, d.scales, d.rotations = means3D.data_ptr(), opacities.data_ptr(), scales.data_ptr(), rotations.data_ptr()
        d.colors0 = colors0.data_ptr()
        d.unnorm_rotations = unnorm_rotations.data_ptr() if unnorm_rotations is not None else None
        d.colors1 = colors1.data_ptr() if colors1 is not None else None
        d.radii, d.status = radii.data_ptr(), status.data_ptr()
        if sticky is not None:
            if sticky.dtype != torch.int32 or not sticky.is_cuda or sticky.numel() < 2:
                raise ValueError("sticky must be a CUDA int32 tensor with 2 elements")
            d.sticky = sticky.data_ptr()

        if capacity is None:
            sz = _workspace_bytes(G, W, H, n_sets, 0)
            geom = torch.empty(sz[0], dtype=torch.uint8, device=dev)
            d.geom_ws = geom.data_ptr()
            d.capacity = 0
            _lib.check(lib.gsd_raster_count_instances(C.byref(d), st), "gsd_raster_count_instances")
            capacity = int(status[0].item())  # the one D2H sync upstream also performs
        capacity = max(int(capacity), 1)
        sz = _workspace_bytes(G, W, H, n_sets, capacity)
        geom = torch.empty(sz[0], dtype=torch.uint8, device=dev)
        binning = torch.empty(sz[1], dtype=torch.uint8, device=dev)
        image = torch.empty(sz[2], dtype=torch.uint8, device=dev)
        color = torch.empty((3 * n_sets, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        d.capacity = capacity
        d.geom_ws, d.binning_ws, d.image_ws = geom.data_ptr(), binning.data_ptr(), image.data_ptr()
        d.out_color, d.out_depth = color.data_ptr(), depth.data_ptr()
        with _nvtx.range("gsd.raster_forward"):
            _lib.check(lib.gsd_raster_forward(C.byref(d), st), "gsd_raster_forward")

    state = RasterState()
    state.desc = d
    state.keep = (means3D, opacities, colors0, colors1, scales, rotations, unnorm_rotations, view, proj, bg0, bg1, geom, binning, image,
                  color, depth, radii, status, sticky)
    state.capacity, state.G, state.W, state.H, state.n_sets, state.status = capacity, G, W, H, n_sets, status
    return color, radii, depth, state


def raster_backward_prepare(state):
    """The blend backward's per-chunk prefix pass (incoming transmittance / colour in front of every chunk).  It depends only on
    the forward's state, so it can run on a side stream while the image gradient is computed; pass prefix_done=True to
    raster_backward afterwards (the caller orders the two with events)."""
    b = _lib.GsdRasterBwd()
    b.fwd = state.desc
    with torch.cuda.device(state.status.device):
        _lib.check(_lib.lib().gsd_raster_backward_stage(C.byref(b), 4, _stream()), "gsd_raster_backward_stage")


def raster_backward(state, grad_color, need_means2D=True, geom_only=False, fused_update=None, prefix_done=False):
    """Runs the backward kernels. Returns dict of gradients (float32 CUDA tensors). geom_only: colours and opacities are
    frozen (steady-state tracking) — their gradients are not produced and the blend backward reduces 5 values per instance.
    fused_update (a filled _lib.GsdTrackUpdate; implies geom_only, no means2D): the per-Gaussian backward kernel applies the
    tracker's update itself (gsd_track_backward_update) and nothing is returned."""
    lib = _lib.lib()
    G, n_sets = state.G, state.n_sets
    dev = grad_color.device
    grad_color = _f32c(grad_color, "grad_color")
    if grad_color.shape != (3 * n_sets, state.H, state.W):
        raise ValueError("grad_color has wrong shape")
    with torch.cuda.device(dev):
        sz = _workspace_bytes(G, state.W, state.H, n_sets, state.capacity)
        partial = torch.empty(sz[3], dtype=torch.uint8, device=dev)
        b = _lib.GsdRasterBwd()
        b.fwd = state.desc
        b.dL_dcolor, b.partial_ws = grad_color.data_ptr(), partial.data_ptr()
        b.prefix_done = 1 if prefix_done else 0
        if fused_update is not None:
            with _nvtx.range("gsd.raster_backward_update"):
                _lib.check(lib.gsd_track_backward_update(C.byref(b), C.byref(fused_update), _stream()), "gsd_track_backward_update")
            return None
        g = dict(means3D=torch.empty((G, 3), dtype=torch.float32, device=dev),
                 means2D=torch.empty((G, 3), dtype=torch.float32, device=dev) if need_means2D else None,
                 colors0=torch.empty((G, 3), dtype=torch.float32, device=dev) if not geom_only else None,
                 colors1=torch.empty((G, 3), dtype=torch.float32, device=dev) if (n_sets == 2 and not geom_only) else None,
                 opacities=torch.empty((G, 1), dtype=torch.float32, device=dev) if not geom_only else None,
                 scales=torch.empty((G, 3), dtype=torch.float32, device=dev),
                 rotations=torch.empty((G, 4), dtype=torch.float32, device=dev))
        ptr = lambda t: t.data_ptr() if t is not None else None
        b.dL_dmeans3D, b.dL_dmeans2D = ptr(g["means3D"]), ptr(g["means2D"])
        b.dL_dcolors0, b.dL_dcolors1 = ptr(g["colors0"]), ptr(g["colors1"])
        b.dL_dopacities, b.dL_dscales, b.dL_drotations = ptr(g["opacities"]), ptr(g["scales"]), ptr(g["rotations"])
        with _nvtx.range("gsd.raster_backward"):
            _lib.check(lib.gsd_raster_backward(C.byref(b), _stream()), "gsd_raster_backward")
    return g


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, colors_precomp, scales, rotations, settings):
        color, radii, depth, state = raster_forward(settings, means3D, opacities, colors_precomp, scales, rotations)
        ctx.state = state
        ctx.opac_shape = opacities.shape
        ctx.mark_non_differentiable(radii, depth)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth):
        g = raster_backward(ctx.state, grad_color)
        ctx.state = None
        return (g["means3D"], g["means2D"], g["opacities"].reshape(ctx.opac_shape), g["colors0"], g["scales"],
                g["rotations"], None)


def rasterize_gaussians(means3D, means2D, opacities, colors_precomp, scales, rotations, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, opacities, colors_precomp, scales, rotations, raster_settings)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            positions = _f32c(positions, "positions")
            G = positions.shape[0]
            vis = torch.empty(G, dtype=torch.uint8, device=positions.device)
            view = _f32c(self.raster_settings.viewmatrix, "viewmatrix").reshape(-1)
            with torch.cuda.device(positions.device):
                _lib.check(_lib.lib().gsd_raster_mark_visible(G, positions.data_ptr(), view.data_ptr(), vis.data_ptr(),
                                                              _stream()), "gsd_raster_mark_visible")
            return vis.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is not None or cov3D_precomp is not None or int(rs.sh_degree) != 0:
            raise NotImplementedError("shs / cov3D_precomp / sh_degree>0 are not used by gs-dynamics and not implemented")
        return rasterize_gaussians(means3D, means2D, opacities, colors_precomp, scales, rotations, rs)
