/*
 * gsd.h — C ABI of libgsd_b200.so: the B200-native (sm_100a) hot paths of robo-alex/gs-dynamics.
 *
 * Path A: differentiable 3D-Gaussian rasterizer with depth + the per-iteration tracking losses.
 *   Replaces the pybind11 module `diff_gaussian_rasterization._C` of
 *   JonathonLuiten/diff-gaussian-rasterization-w-depth (un-vendored; cloned per
 *   /root/reference/README.md:26-35) that the reference reaches through
 *   GaussianRasterizer(...)(...)  — /root/reference/src/tracking/train_utils.py:178,192,379;
 *   /root/reference/src/render/renderer.py:22; /root/reference/src/real_world/gs/trainer.py:61.
 * Path B: GNN particle-dynamics step (edge construction, gather, fused edge->node segment reduce, FPS).
 *   Replaces the dense one-hot torch.bmm formulation of /root/reference/src/gnn/model.py:112-246,
 *   /root/reference/src/data/dataset.py:88-216 and dgl.geometry.farthest_point_sampler
 *   (/root/reference/src/render/dynamics_module.py:46,65).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; all buffers are allocated by the
 *     caller (the library owns no device memory);
 *   - every entry point is asynchronous on the given stream (a cudaStream_t passed as void*), makes no
 *     host synchronisation, and returns 0 on success or a negative GsdStatus; gsd_last_error() returns a
 *     thread-local message;
 *   - all arithmetic fp32; indices int32 unless stated; matrices follow the reference's layout
 *     (viewmatrix = w2c^T, projmatrix = (P w2c)^T, both 16 contiguous floats in DEVICE memory, exactly the
 *     tensors the reference stores in GaussianRasterizationSettings).
 */
#ifndef GSD_H_
#define GSD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GSD_OK = 0,
    GSD_ERR_INVALID = -1,  /* bad argument */
    GSD_ERR_CUDA = -2,     /* a CUDA call failed; see gsd_last_error() */
    GSD_ERR_CAPACITY = -3, /* workspace too small */
    GSD_ERR_UNSUPPORTED = -4
} GsdStatus;

const char *gsd_last_error(void);
int gsd_version(void);

/* ------------------------------------------------------------------------------------------------
 * Path A.1 — rasterizer (replaces _C.rasterize_gaussians / _C.rasterize_gaussians_backward)
 * ------------------------------------------------------------------------------------------------ */

/* Device status block written by the forward pass (8 x int32):
 *   [0] n_instances R (sum of tiles touched)   [1] overflow flag (R > capacity; instances dropped)
 *   [2] n_visible Gaussians (radii > 0)        [3..7] reserved */
#define GSD_STATUS_WORDS 8

typedef struct {
    int32_t G;            /* number of Gaussians */
    int32_t W, H;         /* image size in pixels */
    int32_t n_sets;       /* 1: colors0 only (3 channels); 2: colors0+colors1 rendered in one pass (6 channels) */
    int64_t capacity;     /* max tile instances the binning / partial buffers can hold */
    float tanfovx, tanfovy, scale_modifier;
    /* inputs [device] */
    const float *viewmatrix; /* [16] = w2c^T (reference: raster_settings.viewmatrix) */
    const float *projmatrix; /* [16] = (P w2c)^T (reference: raster_settings.projmatrix) */
    const float *bg0;        /* [3] background of colour set 0 (reference: raster_settings.bg) */
    const float *bg1;        /* [3] background of colour set 1, or NULL (= zeros) */
    const float *means3D;   /* [G,3] */
    const float *opacities; /* [G]   */
    const float *scales;    /* [G,3] */
    const float *rotations; /* [G,4] (r,x,y,z) used as given */
    const float *colors0;   /* [G,3] */
    const float *colors1;   /* [G,3] or NULL */
    /* outputs [device] */
    float *out_color;  /* [3*n_sets,H,W] */
    float *out_depth;  /* [H,W] */
    int32_t *radii;    /* [G] */
    /* caller-allocated workspaces; sizes from gsd_raster_workspace_bytes(); geom/binning/image must be
     * kept unchanged until the matching backward call */
    void *geom_ws;
    void *binning_ws;
    void *image_ws;
    int32_t *status; /* [GSD_STATUS_WORDS] */
} GsdRasterFwd;

typedef struct {
    GsdRasterFwd fwd;           /* the same descriptor the forward call used */
    const float *dL_dcolor;     /* [3*n_sets,H,W] */
    void *partial_ws;           /* scratch, size out[3] of gsd_raster_workspace_bytes() */
    /* outputs [device], fully overwritten */
    float *dL_dmeans3D;   /* [G,3] */
    float *dL_dmeans2D;   /* [G,3] NDC-scaled screen-space gradient (xy), z = 0; may be NULL */
    float *dL_dcolors0;   /* [G,3]; may be NULL */
    float *dL_dcolors1;   /* [G,3]; may be NULL */
    float *dL_dopacities; /* [G] */
    float *dL_dscales;    /* [G,3] */
    float *dL_drotations; /* [G,4] */
} GsdRasterBwd;

/* out[0] geom_ws, out[1] binning_ws, out[2] image_ws, out[3] partial_ws (backward scratch) — bytes */
int gsd_raster_workspace_bytes(int32_t G, int32_t W, int32_t H, int32_t n_sets, int64_t capacity, size_t out[4]);

/* preprocess + prefix sum only: fills status[0] with the exact instance count R (what upstream copies
 * back to the host before binning). Needs geom_ws only. */
int gsd_raster_count_instances(const GsdRasterFwd *a, void *stream);

/* preprocess -> scan -> duplicate keys -> radix sort -> pack per-tile records -> blend */
int gsd_raster_forward(const GsdRasterFwd *a, void *stream);

/* blend backward (deterministic, atomic-free) -> cov2D/projection/cov3D backward */
int gsd_raster_backward(const GsdRasterBwd *a, void *stream);

/* replaces _C.mark_visible: visible[g] = (view-space z > 0.2) */
int gsd_raster_mark_visible(int32_t G, const float *means3D, const float *viewmatrix, uint8_t *visible, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GSD_H_ */
