"""ctypes front-end of oracle/raster_oracle.c — TEST INFRASTRUCTURE ONLY (see that file's header)."""
import ctypes as C
import numpy as np

from . import build as _build

_lib = None


class _In(C.Structure):
    _fields_ = [("G", C.c_int), ("W", C.c_int), ("H", C.c_int), ("CH", C.c_int),
                ("means3D", C.c_void_p), ("colors", C.c_void_p), ("opacities", C.c_void_p),
                ("scales", C.c_void_p), ("rotations", C.c_void_p), ("viewmatrix", C.c_void_p),
                ("projmatrix", C.c_void_p), ("bg", C.c_void_p),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.gsd_oracle_raster_forward.restype = C.c_int64
        _lib.gsd_oracle_raster_backward.restype = C.c_int64
        _lib.gsd_oracle_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(lib().gsd_oracle_num_threads())


def set_num_threads(n):
    """OpenMP team size of the C oracle (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    lib().gsd_oracle_set_num_threads(int(n))


def _f32(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _pack(means3D, colors, opacities, scales, rotations, viewmatrix, projmatrix, bg, tanfovx, tanfovy,
          H, W, scale_modifier):
    arrs = dict(means3D=_f32(means3D).reshape(-1, 3), colors=_f32(colors), opacities=_f32(opacities).reshape(-1),
                scales=_f32(scales).reshape(-1, 3), rotations=_f32(rotations).reshape(-1, 4),
                viewmatrix=_f32(viewmatrix).reshape(16), projmatrix=_f32(projmatrix).reshape(16), bg=_f32(bg).reshape(-1))
    G = arrs["means3D"].shape[0]
    arrs["colors"] = arrs["colors"].reshape(G, -1) if G > 0 else arrs["colors"].reshape(0, arrs["bg"].shape[0])
    CH = arrs["bg"].shape[0]
    assert arrs["colors"].shape[1] == CH
    s = _In()
    s.G, s.W, s.H, s.CH = G, W, H, CH
    for k in ("means3D", "colors", "opacities", "scales", "rotations", "viewmatrix", "projmatrix", "bg"):
        setattr(s, k, arrs[k].ctypes.data)
    s.tanfovx, s.tanfovy, s.scale_modifier = tanfovx, tanfovy, scale_modifier
    return s, arrs, G, CH


def forward(means3D, colors, opacities, scales, rotations, viewmatrix, projmatrix, bg, tanfovx, tanfovy,
            H, W, scale_modifier=1.0, debug=False):
    s, keep, G, CH = _pack(means3D, colors, opacities, scales, rotations, viewmatrix, projmatrix, bg,
                           tanfovx, tanfovy, H, W, scale_modifier)
    color = np.zeros((CH, H, W), np.float32)
    depth = np.zeros((1, H, W), np.float32)
    radii = np.zeros((max(G, 1),), np.int32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.int32)
    dbg = [None] * 4
    if debug:
        dbg = [np.zeros((max(G, 1), 2), np.float32), np.zeros((max(G, 1), 4), np.float32),
               np.zeros((max(G, 1),), np.float32), np.zeros((max(G, 1),), np.uint32)]
    p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)
    R = lib().gsd_oracle_raster_forward(C.byref(s), p(color), p(depth), p(radii), p(final_T), p(n_contrib),
                                        p(dbg[0]), p(dbg[1]), p(dbg[2]), p(dbg[3]))
    if R < 0:
        raise RuntimeError("oracle forward failed")
    out = dict(color=color, depth=depth, radii=radii[:G], final_T=final_T, n_contrib=n_contrib, R=int(R))
    if debug:
        out.update(xy=dbg[0][:G], conic_o=dbg[1][:G], gdepth=dbg[2][:G], tiles_touched=dbg[3][:G])
    return out


def backward(means3D, colors, opacities, scales, rotations, viewmatrix, projmatrix, bg, tanfovx, tanfovy,
             H, W, dL_dcolor, scale_modifier=1.0):
    s, keep, G, CH = _pack(means3D, colors, opacities, scales, rotations, viewmatrix, projmatrix, bg,
                           tanfovx, tanfovy, H, W, scale_modifier)
    dL = _f32(dL_dcolor).reshape(CH, H, W)
    n = max(G, 1)
    out = dict(means3D=np.zeros((n, 3), np.float32), means2D=np.zeros((n, 3), np.float32),
               colors=np.zeros((n, CH), np.float32), opacities=np.zeros((n,), np.float32),
               scales=np.zeros((n, 3), np.float32), rotations=np.zeros((n, 4), np.float32))
    p = lambda a: C.c_void_p(a.ctypes.data)
    R = lib().gsd_oracle_raster_backward(C.byref(s), p(dL), p(out["means3D"]), p(out["means2D"]), p(out["colors"]),
                                         p(out["opacities"]), p(out["scales"]), p(out["rotations"]))
    if R < 0:
        raise RuntimeError("oracle backward failed")
    return {k: v[:G] for k, v in out.items()}
