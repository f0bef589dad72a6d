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
/* host-side count of kernels launched by this library since load: own kernels / CUB (library) kernels */
void gsd_launch_count(long long *own_kernels, long long *library_kernels);

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
    int32_t *status; /* [GSD_STATUS_WORDS], rewritten by every forward call */
    int32_t *sticky; /* [2] or NULL; never cleared by the library: [0] = max R over the forward calls so far, [1] = number of
                      * forward calls that overflowed `capacity`. Lets a caller that replays a captured graph thousands of
                      * times check ONCE per frame that no replay dropped instances (the reference sizes its buffers exactly
                      * from num_rendered on every call, /root/reference/src/tracking/train_utils.py:178) */
    const float *unnorm_rotations; /* [G,4] or NULL.  When set, `rotations` is an OUTPUT: the preprocess kernel computes
                      * rotations = F.normalize(unnorm_rotations) (params2rendervar, helpers.py:40) on the fly, uses it and writes
                      * it to `rotations` (a writable [G,4] buffer) for the backward — the tracker's iteration then needs no
                      * separate normalisation launch */
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
    float *dL_dopacities; /* [G]; may be NULL. dL_dcolors0 = dL_dcolors1 = dL_dopacities = NULL selects the geometry-only
                           * backward (colours / opacities frozen): 5 instead of 9-12 partials per instance */
    float *dL_dscales;    /* [G,3] */
    float *dL_drotations; /* [G,4] */
    int32_t prefix_done;  /* != 0: gsd_raster_backward_stage(a, 4, other_stream) already ran the per-chunk prefix pass on this forward
                           * state (and the caller ordered this call behind it) */
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

/* one stage of the backward alone (1 = blend backward, 2 = per-Gaussian backward; bench.py times stage 1 for the roofline;
 * 4 = only the blend backward's per-chunk prefix pass, which needs the forward's state but neither dL_dcolor nor partial_ws) */
int gsd_raster_backward_stage(const GsdRasterBwd *a, int32_t stage, void *stream);

/* replaces _C.mark_visible: visible[g] = (view-space z > 0.2) */
int gsd_raster_mark_visible(int32_t G, const float *means3D, const float *viewmatrix, uint8_t *visible, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path A.2 — per-iteration tracking losses (replace the eager PyTorch graph of get_loss,
 * /root/reference/src/tracking/train_utils.py:167-246)
 * ------------------------------------------------------------------------------------------------ */

/* loss_set = w_l1*mean|x-y| + w_ssim*(1 - mean SSIM_11x11(x,y))   (train_utils.py:185,195; external.py:101-135)
 * x, y: [C,H,W]; channels come in sets of 3 (n_sets = 2: RGB render + seg render of one iteration in one launch, C = 6;
 * n_sets = 1: any C <= 6). Optionally the first set is colour-corrected on the fly, x = exp(affine_log_scale[c]) * x +
 * affine_shift[c] (the cam_m / cam_c rows of train_utils.py:182; device pointers to 3 floats, or NULL).
 * ws keeps three partial-derivative maps between forward and backward.
 * loss_out: per set {loss, mean|x-y|, mean SSIM}, then sum_s set_weight[s]*loss_s  (3*n_sets + 1 floats). */
typedef struct {
    int32_t C, H, W, n_sets;
    const float *x, *y;
    const float *affine_log_scale, *affine_shift;
    float w_l1, w_ssim;
    float set_weight[2];
    void *ws;
    const float *y_mu, *y_s22; /* optional [C,H,W] window statistics of y from gsd_photometric_target_stats (or NULL) */
} GsdPhotometric;
int gsd_photometric_workspace_bytes(int32_t C, int32_t H, int32_t W, size_t *bytes);
int gsd_photometric_forward(const GsdPhotometric *p, float *loss_out, void *stream);
/* the two halves of gsd_photometric_forward: the statistics pass the gradient needs, and the scalar reduction, which only
 * reads ws and may run on another stream.  add (nullable device scalar): loss_out gets one more float, total + *add. */
int gsd_photometric_stats(const GsdPhotometric *p, void *stream);
int gsd_photometric_reduce(const GsdPhotometric *p, const float *add, float *loss_out, void *stream);
/* conv(y), conv(y*y) with the SSIM window: constant per target image, computed once per camera and frame */
int gsd_photometric_target_stats(int32_t C, int32_t H, int32_t W, const float *y, float *y_mu, float *y_s22, void *stream);
/* target planes from the dataset's bytes (train_utils.py:66-75): im_hwc [H,W,im_channels] uint8 (PIL layout, 3 or 4 channels),
 * seg [H,W] uint8 -> target6 [6,H,W] float = (im / 255 | seg, 0, 1 - seg): what the reference builds on the host before its upload */
int gsd_track_unpack_target_u8(int32_t H, int32_t W, int32_t im_channels, const uint8_t *im_hwc, const uint8_t *seg, float *target6,
                               void *stream);
/* grad = (gscale_ptr ? *gscale_ptr : 1) * set_weight[set] * d loss_set / d x_rendered (before the affine) */
int gsd_photometric_backward(const GsdPhotometric *p, const float *gscale_ptr, float *grad, void *stream);

/* rigid / rot / iso / floor / bg priors, forward + gradient in one call (train_utils.py:198-240).
 * The foreground set is fg_index (NULL = all G points, in order); neighbour tables index into that set.
 * in_ptr/in_edge: transposed adjacency (CSR over the neighbour id) of the static kNN graph, edge id = i*K+k. */
typedef struct {
    int32_t G, Gf, K, Gb;
    const float *means3D;           /* [G,3] */
    const float *rotations;         /* [G,4] normalised (what params2rendervar feeds the rasterizer) */
    const int32_t *fg_index;        /* [Gf] or NULL */
    const float *prev_inv_rot;      /* [Gf,4]  variables["prev_inv_rot_fg"] */
    const int32_t *neighbor_indices;/* [Gf,K]  variables["neighbor_indices"] (int32) */
    const float *neighbor_weight;   /* [Gf,K] */
    const float *neighbor_dist;     /* [Gf,K] */
    const float *prev_offset;       /* [Gf,K,3] */
    const int32_t *in_ptr;          /* [Gf+1] */
    const int32_t *in_edge;         /* [Gf*K] */
    const float *edge_records;      /* optional [Gf*K,8] from gsd_track_pack_edges (32-byte records: 4x fewer L1 sectors) */
    const int32_t *bg_index;        /* [Gb] */
    const float *init_bg_pts;       /* [Gb,3] */
    const float *init_bg_rot;       /* [Gb,4] */
    float w_rigid, w_rot, w_iso, w_floor, w_bg; /* loss weights (train_utils.py:236-240) */
    void *ws;                       /* gsd_track_losses_workspace_bytes() */
    float *losses;                  /* [6] rigid, rot, iso, floor, bg (unweighted means), weighted total */
    float *grad_means3D;            /* [G,3] d(weighted total)/d means3D, fully overwritten */
    float *grad_rotations;          /* [G,4] d(weighted total)/d rotations, fully overwritten */
    int32_t rotations_unnormalized; /* != 0: `rotations` holds params['unnorm_rotations']; F.normalize is applied on the fly (the
                                     * gradient is still w.r.t. the NORMALISED rotations, as gsd_track_update expects) */
} GsdTrackLosses;
int gsd_track_losses_workspace_bytes(int32_t Gf, int32_t Gb, size_t *bytes);
/* packs (neighbour id, weight, rest distance, previous offset) of every edge into one 32-byte record; call once per
 * timestep (initialize_per_timestep changes prev_offset, train_utils.py:331-351) */
int gsd_track_pack_edges(int32_t Gf, int32_t K, const int32_t *neighbor_indices, const float *neighbor_weight,
                         const float *neighbor_dist, const float *prev_offset, float *edge_records, void *stream);
int gsd_track_losses_fwd_bwd(const GsdTrackLosses *t, void *stream);

/* Multi-tensor Adam, one launch for every parameter group (torch.optim.Adam semantics, no weight decay /
 * amsgrad; eps inside the bias-corrected denominator) — initialize_optimizer, train_utils.py:152-164.
 * step[i] is a device-resident float counter per group, incremented by the kernel (CUDA-graph friendly). */
#define GSD_ADAM_MAX_TENSORS 16
typedef struct {
    int32_t n_tensors;
    float beta1, beta2, eps;
    float *param[GSD_ADAM_MAX_TENSORS];
    const float *grad[GSD_ADAM_MAX_TENSORS];
    float *exp_avg[GSD_ADAM_MAX_TENSORS];
    float *exp_avg_sq[GSD_ADAM_MAX_TENSORS];
    float *step[GSD_ADAM_MAX_TENSORS];
    float lr[GSD_ADAM_MAX_TENSORS];
    int64_t numel[GSD_ADAM_MAX_TENSORS];
} GsdAdam;
int gsd_adam_step(const GsdAdam *a, void *stream);

/* Steady-state (t > 0) fast path of the tracker, where only means3D and unnorm_rotations have a non-zero learning rate
 * (train_utils.py:370-373): rotations = F.normalize(unnorm_rotations) (helpers.py:40), and the fused update
 *   g_means = ga + gb;  g_unnorm = normalize_backward(ga_rot + gb_rot);  Adam step on both groups
 * (the two gradient sources are the rasterizer backward and the physical priors). Step counters are device floats,
 * read as (*step + 1) and advanced once per call: by the last CTA to finish when block_counter (a zero-initialised,
 * self-resetting device word owned by the caller) is given, else by a trailing 1-thread kernel.  radii / max_2D_radius /
 * seen (all or none; seen may be NULL) fold the bookkeeping of gsd_track_update_radii into the same launch. */
int gsd_track_normalize_rotations(int32_t G, const float *unnorm_rotations, float *rotations, void *stream);
typedef struct {
    int32_t G;
    float beta1, beta2, eps, lr_means, lr_rot;
    float *means3D, *unnorm_rotations;
    const float *g_means_a, *g_means_b;   /* [G,3] each; b may be NULL */
    const float *g_rot_a, *g_rot_b;       /* [G,4] each, w.r.t. the NORMALISED rotations; b may be NULL */
    float *m_means, *v_means, *m_rot, *v_rot;
    float *step_means, *step_rot;
    const int32_t *radii;                 /* optional [G] */
    float *max_2D_radius;                 /* optional [G] */
    uint8_t *seen;                        /* optional [G] */
    uint32_t *block_counter;              /* optional, see above */
} GsdTrackUpdate;
int gsd_track_update(const GsdTrackUpdate *u, void *stream);
/* gsd_raster_backward (geometry-only: dL_dcolors0/1, dL_dopacities, dL_dmeans2D must be NULL; dL_dmeans3D / dL_dscales /
 * dL_drotations are not written and may be NULL) with gsd_track_update applied inside its per-Gaussian kernel: replaces
 * loss.backward() + optimizer.step() of train_gs.py:31-39 for t > 0.  u->g_means_a / g_rot_a are ignored (the rasterizer's
 * gradients stay in registers); u->g_means_b / g_rot_b (the priors' gradients) are added; u->block_counter is required. */
int gsd_track_backward_update(const GsdRasterBwd *a, const GsdTrackUpdate *u, void *stream);

/* bookkeeping of get_loss (train_utils.py:243-245): seen = radii > 0; max_2D_radius = max(radii, max_2D_radius)[seen] */
int gsd_track_update_radii(int32_t G, const int32_t *radii, float *max_2D_radius, uint8_t *seen, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path B — GNN particle-dynamics step (replaces the one-hot torch.bmm formulation of
 * /root/reference/src/gnn/model.py:112-246 and the dense N x N edge construction of
 * /root/reference/src/data/dataset.py:88-216)
 * ------------------------------------------------------------------------------------------------ */

/* construct_edges_from_states[_batch]: adjacency = (d^2 < r^2) AND top-k nearest (object-object block), tool<->object
 * edges by radius or all (connect_all), no tool-tool edges; edges in adj.nonzero() order = grouped by receiver.
 * Tools must be the last n_tool nodes (dataset.py:120-124). Outputs are index lists, never one-hot matrices:
 * edges of batch element b occupy slots [b*capacity, b*capacity + n_edges[b]); unused slots hold -1. */
typedef struct {
    int32_t B, N, n_tool, topk, connect_all;
    int32_t capacity;            /* edge slots per batch element */
    const float *states;         /* [B,N,3] */
    const uint8_t *mask;         /* [B,N] valid particle */
    const uint8_t *tool_mask;    /* [B,N] */
    const float *adj_thresh;     /* [B] radius per element, or NULL to use adj_thresh_sq_scalar */
    float adj_thresh_sq_scalar;  /* float32(adj_thresh * adj_thresh) as the reference's scalar path computes it */
    void *ws;                    /* gsd_gnn_edges_workspace_bytes() */
    int32_t *row_ptr;            /* [B,N+1] CSR offsets by receiver */
    int32_t *n_edges;            /* [B] */
    int32_t *receivers;          /* [B,capacity] */
    int32_t *senders;            /* [B,capacity] */
} GsdGnnEdges;
int gsd_gnn_edges_workspace_bytes(int32_t B, int32_t N, size_t *bytes);
int gsd_gnn_build_edges(const GsdGnnEdges *g, void *stream);

/* rel_inputs[b,e,:] = [attrs[recv] | attrs[send] | sum_k |g[recv,k]-g[send,k]| | (state[h,recv]-state[h,send]) h=0..n_his-1]
 * (model.py:164-199); state is [B,n_his,N,3], g = p_instance padded with zeros for the n_s shape particles. */
int gsd_gnn_edge_inputs(int32_t B, int32_t N, int32_t capacity, int32_t n_his, int32_t attr_dim, int32_t n_instance,
                        int32_t n_p, const float *state, const float *attrs, const float *p_instance,
                        const int32_t *receivers, const int32_t *senders, float *rel_inputs, void *stream);

/* Rollout glue (replaces the host-side tensor plumbing of one autoregressive step: /root/reference/src/render/dynamics_module.py:
 * 104-111,127,145-158 and the particle-input concatenations of /root/reference/src/gnn/model.py:132-160).
 * pre: action rows of the tool nodes (n >= n_obj) <- eef_delta[b * delta_stride ..+3]; p_inputs [B*N, attr_dim + state_dim n_his +
 *      (motion ? 3 (n_his-1) : 0) + (has_action ? 3 : 0)] = [attrs | state over history (state_dim 3: xyz, 1: z, 0: none) |
 *      frame-to-frame motion | action];
 *      cur [B,N,3] = states[:, -1].   states: [B, n_his, N, 3].
 * post: pred [B,n_obj,3] = states[:, -1, :n_obj] + clamp(motion, +-clampv); tool nodes advance by eef_delta; history shifts by one. */
int gsd_gnn_rollout_pre(int32_t B, int32_t N, int32_t n_obj, int32_t n_his, int32_t attr_dim, int32_t state_dim, int32_t motion, int32_t has_action,
                        const float *states, const float *attrs, float *action, const float *eef_delta, int32_t delta_stride,
                        float *p_inputs, float *cur, void *stream);
int gsd_gnn_rollout_post(int32_t B, int32_t N, int32_t n_obj, int32_t n_his, float *states, const float *motion,
                         const float *eef_delta, int32_t delta_stride, float clampv, float *pred, void *stream);

/* agg[b*N+r, :] = sum over incoming edges e of ReLU(A[b*cap+e, :] + P[b*N+r, 0:F] + P[b*N+send(e), F:2F])
 * (relation propagator epilogue + Rr^T scatter-add of model.py:212-229). The last n_heavy rows of every element
 * (tool nodes) are split over several CTAs and summed in fixed order. F multiple of 128, <= 512.
 * ws: gsd_gnn_aggregate_workspace_bytes() bytes, ZERO-FILLED ONCE by the caller before its first use; the call leaves the
 * counters it holds at zero again, so the same workspace serves any number of calls issued in stream order (no memset per call). */
int gsd_gnn_aggregate_workspace_bytes(int32_t B, int32_t n_heavy, int32_t F, size_t *bytes);
int gsd_gnn_aggregate(int32_t B, int32_t N, int32_t capacity, int32_t F, int32_t n_heavy, const int32_t *row_ptr,
                      const int32_t *senders, const float *A, const float *P, void *ws, float *agg, void *stream);

/* Backward of gsd_gnn_aggregate / gsd_gnn_edge_inputs for GNN training (replaces what autograd derives from the one-hot
 * bmm gathers/scatters of /root/reference/src/gnn/model.py:164-229 under /root/reference/src/train.py:183-211).
 * col_ptr [B,N+1] / order [B,capacity]: the forward's edges grouped by SENDER (order = edge ids sorted by sender, stable).
 *   gA[e]     = g_agg[recv(e)] where A[e] + P[recv,0:F] + P[send,F:2F] > 0, else 0     [B*capacity, F]
 *   gP[n]     = [ sum_{e: recv=n} gA[e] | sum_{e: send=n} gA[e] ]                        [B*N, 2F]
 *   g_state[b,h,n,:] = sum_{e: recv=n} g_rel[b,e,off+3h:off+3h+3] - sum_{e: send=n} (same)  [B,n_his,N,3]
 * No atomics; fixed summation order. */
int gsd_gnn_aggregate_bwd_workspace_bytes(int32_t B, int32_t n_heavy, int32_t F, size_t *bytes);
int gsd_gnn_aggregate_bwd(int32_t B, int32_t N, int32_t capacity, int32_t F, int32_t n_heavy, const int32_t *row_ptr,
                          const int32_t *senders, const int32_t *col_ptr, const int32_t *order, const float *A, const float *P,
                          const float *g_agg, void *ws, float *gA, float *gP, void *stream);
int gsd_gnn_edge_inputs_bwd(int32_t B, int32_t N, int32_t capacity, int32_t n_his, int32_t width, int32_t offset,
                            const int32_t *row_ptr, const int32_t *col_ptr, const int32_t *order, const float *g_rel,
                            float *g_state, void *stream);

/* Operands for the error-compensated TF32 GEMMs of the dense layers (nn.Linear in src/gnn/model.py:16-108 computed as one
 * tensor-core GEMM over K = 3F).  t = relu ? max(x + add, 0) : x + add  (add may be NULL), hi = round_to_tf32(t), lo = t - hi.
 * out [rows, 3F]: activation layout (weight_layout = 0) [lo | hi | hi], weight layout (1) [hi | lo | hi].
 * full (nullable) [rows, F] receives t. */
int gsd_tf32_pack(int64_t rows, int32_t F, int32_t relu, int32_t weight_layout, const float *x, const float *add, float *full,
                  float *out, void *stream);

/* Exact k nearest neighbours of every point among the OTHER points (self excluded), brute force in float64.
 * Replaces o3d_knn, /root/reference/src/tracking/helpers.py:97-115 (Open3D KD-tree + Python loop over points on the CPU).
 * pts float32 [n,3]; sq_dist float64 [n,k] squared distances ascending; idx int32 [n,k].  k <= 64, k <= n-1. */
int gsd_knn(int32_t n, int32_t k, const float *pts, double *sq_dist, int32_t *idx, void *stream);

/* ---- linear-blend skinning of the Gaussians from the GNN particles -------------------------------
 * Replaces interpolate_motions, /root/reference/src/render/utils.py:129-243, called once per rollout step from
 * /root/reference/src/render/dynamics_module.py:150-156.
 *
 * gsd_skin_bone_transforms: per bone i, neighbours = columns j < n_bones of CSR row i (row_ptr [n_bones+1], cols); Procrustes
 * rotation R_i of the neighbourhood (utils.py:150-204, incl. the rank-1 construction and the identity fallbacks).
 * bone_tf [n_bones, 20] floats: R (9, row-major) | c = motion + b - R b (3) | normalised mat2quat(R) (4, wxyz) | b (3) | pad.
 * rot_out (nullable) [n_bones, 9] receives R. */
int gsd_skin_bone_transforms(int32_t n_bones, const float *bones, const float *motions, const int32_t *row_ptr,
                             const int32_t *cols, float *bone_tf, float *rot_out, void *stream);

/* gsd_skin_apply: per particle p, w_b = 1 / max(|xyz_p - b|, 1e-4) normalised over all bones (utils.py:206-213), or the rows
 * of weights_in [n_particles, n_bones] used as they are; xyz_out = sum_b w_b (R_b xyz_p + c_b); when quat is given,
 * quat_out = normalize(sum_b w_b q_b) (x) quat_p (utils.py:216-237).  weights_out (nullable) [n_particles, n_bones] receives the
 * dense normalised weights the reference returns. */
int gsd_skin_apply(int32_t n_particles, int32_t n_bones, const float *xyz, const float *quat, const float *bone_tf,
                   const float *weights_in, float *xyz_out, float *quat_out, float *weights_out, void *stream);

/* farthest point sampling, one CTA per batch element. radius <= 0: dgl.geometry.farthest_point_sampler(pos, npoints,
 * start_idx) (squared distances, first maximum). radius > 0: fps_rad_idx_torch (data/utils.py:50-65): stops when the
 * largest euclidean distance to the picked set is <= radius; count[b] = number of picks, unused outputs = -1. */
int gsd_fps(int32_t B, int32_t N, int32_t npoints, float radius, const float *pos, const int64_t *start_idx,
            int64_t *out_idx, int32_t *count, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path A.3 — first-frame densification (replaces densify / cat_params_to_optimizer / remove_points /
 * update_params_and_optimizer, /root/reference/src/tracking/external.py:145-299)
 * ------------------------------------------------------------------------------------------------ */

/* Classifies every point (keep / clone / split), evaluates the prune predicate of every row the round would create and assigns the
 * destination rows in the reference's final order [kept originals | kept clones | kept split copies A | kept split copies B].
 * dst [4][n]: row of the original, of the clone, of split copy A (copy B = + totals[2]), rank among all split candidates (-1: none).
 * totals [4]: kept originals, kept clones, kept split candidates, all split candidates -> n_out = t0 + t1 + 2 t2.
 * do_densify = 0 plans a plain copy (used for the opacity reset alone). */
typedef struct {
    int32_t n, do_densify;
    float grad_thresh;    /* 0.0002 */
    float clone_limit;    /* scale_scene_radius * scene_radius: clone at <=, split above */
    float prune_opacity;  /* remove_thresh (remove_thresh_5k at i == 5000) */
    float prune_big;      /* 0.1 * scene_radius for i >= 3000, <= 0: off */
    const float *grad_accum, *denom;       /* [n] means2D_gradient_accum, denom */
    const float *log_scales;               /* [n,3] */
    const float *logit_opacities;          /* [n] */
    int32_t *dst;                          /* [4][n] */
    int32_t *totals;                       /* [4] */
} GsdDensifyPlan;
int gsd_densify_plan(const GsdDensifyPlan *p, void *stream);

/* Writes the new arrays.  Tensor order: means3D(3) rgb_colors(3) seg_colors(3) unnorm_rotations(4) logit_opacities(1) log_scales(3);
 * p = parameter, m / v = Adam exp_avg / exp_avg_sq (appended rows get zeros, kept rows carry theirs).  samples: [2 * totals[3], 3]
 * normal draws for the split copies (row r: copy A of split rank r, row totals[3] + r: copy B); samples_scaled = 1 if they are already
 * multiplied by exp(log_scales).  reset_opacity: every output opacity = logit(0.01) with zero moments (external.py:293-295). */
typedef struct {
    int32_t n, samples_scaled, reset_opacity;
    const int32_t *dst, *totals;
    const float *samples;
    const float *p_src[6], *m_src[6], *v_src[6];
    float *p_dst[6], *m_dst[6], *v_dst[6];
    int32_t width[6];
} GsdDensifyApply;
int gsd_densify_apply(const GsdDensifyApply *a, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path B — dense layers on the tcgen05 tensor cores, fp32-accurate (replace nn.Linear of
 * /root/reference/src/gnn/model.py:16-22,36-47,58-67 as composed at model.py:202-241)
 * ------------------------------------------------------------------------------------------------ */

/* w_hi = tf32(w), w_lo = tf32(w - w_hi): the two TF32 operands of a weight matrix (once per weight update). */
int gsd_tf32_split(int64_t n, const float *w, float *w_hi, float *w_lo, void *stream);

/* out[M,N] = act(A[M,K] . W[N,K]^T + bias[N] + res1[M,N] + res2[M,N]); act = ReLU if relu else identity; bias / res1 / res2 may be
 * NULL; res1 / res2 / out share the row stride ldo, A has row stride lda (elements).  Error-compensated 3xTF32 (a_lo w_hi + a_hi w_lo
 * + a_hi w_hi, two fp32 accumulators in tensor memory): fp32-level accuracy.  Needs K % 32 == 0, N % 64 == 0, 16-byte aligned rows. */
int gsd_linear_tf32x3(int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *W_hi, const float *W_lo,
                      const float *bias, const float *res1, const float *res2, int32_t relu, float *out, int64_t ldo, void *stream);

/* the two layer shapes of the model that are not tensor-core work, in plain fp32: K <= 32 (first encoder layers; N % 4 == 0,
 * dense rows, optional ReLU) or N <= 8 (the 512 -> 3 motion head; K % 4 == 0, no activation). */
int gsd_linear_small(int64_t M, int32_t N, int32_t K, const float *x, int64_t ldx, const float *W, const float *bias, int32_t relu,
                     float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GSD_H_ */
