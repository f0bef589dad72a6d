// raster_api.cu — C-ABI entry points of the rasterizer (include/gsd.h, Path A.1).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

// launchers implemented in the other translation units
int gsd_launch_preprocess(int G, const GsdCam &cam, const GsdRasterFwd *a, const GsdGeomWs &g, int32_t *zero_fill, int n_zero, cudaStream_t st);
int gsd_launch_count(int G, const GsdGeomWs &g, int32_t *status, cudaStream_t st);
int gsd_launch_mark_visible(int G, const GsdCam &cam, const float *means3D, uint8_t *vis, cudaStream_t st);
int gsd_launch_binning(int G, const GsdCam &cam, const GsdRasterFwd *a, const GsdGeomWs &g, const GsdBinWs &b, int32_t *zero_flags, int n_flags, cudaStream_t st);
int gsd_launch_render_fwd(const GsdRenderParams &p, int tiles, int n_sets, cudaStream_t st);
int gsd_launch_render_bwd(const GsdRenderParams &p, int tiles, int n_sets, int which, cudaStream_t st);
int gsd_launch_preprocess_bwd(int G, const GsdCam &cam, const GsdRasterBwd *a, const GsdGeomWs &g, int geom_only, cudaStream_t st);
int gsd_launch_preprocess_bwd_update(int G, const GsdCam &cam, const GsdRasterBwd *a, const GsdGeomWs &g, const GsdTrackUpdate &u, cudaStream_t st);

#include <atomic>
static thread_local char g_err[512] = "";
static std::atomic<long long> g_own_launches{0}, g_lib_launches{0};
void gsd_count_launch(int own, int library) { g_own_launches += own; g_lib_launches += library; }
extern "C" void gsd_launch_count(long long *own_kernels, long long *library_kernels) {
    if (own_kernels) *own_kernels = g_own_launches.load();
    if (library_kernels) *library_kernels = g_lib_launches.load();
}

void gsd_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool gsd_pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("GSD_NO_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
    return on == 1;
}

extern "C" const char *gsd_last_error(void) { return g_err; }
extern "C" int gsd_version(void) { return 100; }

static int make_cam(const GsdRasterFwd *a, GsdCam *cam) {
    if (!a) { gsd_set_error("null descriptor"); return GSD_ERR_INVALID; }
    if (a->G < 0 || a->W <= 0 || a->H <= 0 || (a->n_sets != 1 && a->n_sets != 2)) {
        gsd_set_error("invalid sizes G=%d W=%d H=%d n_sets=%d", a->G, a->W, a->H, a->n_sets);
        return GSD_ERR_INVALID;
    }
    if (a->capacity < 0 || a->capacity > 0x7fffffffLL) { gsd_set_error("capacity out of range"); return GSD_ERR_INVALID; }
    if (!a->viewmatrix || !a->projmatrix || !a->bg0) { gsd_set_error("null camera pointer"); return GSD_ERR_INVALID; }
    cam->view = a->viewmatrix;
    cam->proj = a->projmatrix;
    cam->tanfovx = a->tanfovx;
    cam->tanfovy = a->tanfovy;
    cam->focal_x = a->W / (2.0f * a->tanfovx);
    cam->focal_y = a->H / (2.0f * a->tanfovy);
    cam->scale_modifier = a->scale_modifier;
    cam->W = a->W;
    cam->H = a->H;
    cam->gx = (a->W + GSD_TILE - 1) / GSD_TILE;
    cam->gy = (a->H + GSD_TILE - 1) / GSD_TILE;
    if (cam->gx > 0xffff || cam->gy > 0xffff) { gsd_set_error("image too large"); return GSD_ERR_INVALID; }
    return GSD_OK;
}

extern "C" int gsd_raster_workspace_bytes(int32_t G, int32_t W, int32_t H, int32_t n_sets, int64_t capacity, size_t out[4]) {
    if (G < 0 || W <= 0 || H <= 0 || capacity < 0 || !out) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (n_sets != 1 && n_sets != 2) { gsd_set_error("n_sets must be 1 or 2"); return GSD_ERR_INVALID; }
    GsdGeomWs g; GsdBinWs b; GsdImgWs im;
    int rc;
    if ((rc = gsd_carve_geom(G, nullptr, &g))) return rc;
    int tiles = ((W + GSD_TILE - 1) / GSD_TILE) * ((H + GSD_TILE - 1) / GSD_TILE);
    if ((rc = gsd_carve_bin(G, capacity, tiles, nullptr, &b))) return rc;
    if ((rc = gsd_carve_img(W, H, n_sets, b.max_items, nullptr, &im))) return rc;
    out[0] = g.total;
    out[1] = b.total;
    out[2] = im.total;
    out[3] = gsd_align_up((size_t)(capacity > 0 ? capacity : 1) * GSD_PART_FLOATS * 4);
    return GSD_OK;
}

extern "C" int gsd_raster_count_instances(const GsdRasterFwd *a, void *stream) {
    GsdCam cam;
    int rc;
    if ((rc = make_cam(a, &cam))) return rc;
    if (!a->geom_ws || !a->status || (a->G > 0 && !a->radii)) { gsd_set_error("null workspace"); return GSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    GsdGeomWs g;
    if ((rc = gsd_carve_geom(a->G, a->geom_ws, &g))) return rc;
    GSD_CUDA_CHECK(cudaMemsetAsync(a->status, 0, GSD_STATUS_WORDS * 4, st));
    if ((rc = gsd_launch_preprocess(a->G, cam, a, g, nullptr, 0, st))) return rc;
    return gsd_launch_count(a->G, g, a->status, st);
}

extern "C" int gsd_raster_forward(const GsdRasterFwd *a, void *stream) {
    GsdCam cam;
    int rc;
    if ((rc = make_cam(a, &cam))) return rc;
    if (!a->geom_ws || !a->binning_ws || !a->image_ws || !a->status || (a->G > 0 && !a->radii) || !a->out_color || !a->out_depth) {
        gsd_set_error("null workspace/output pointer");
        return GSD_ERR_INVALID;
    }
    if (a->G > 0 && (!a->means3D || !a->opacities || !a->scales || !a->rotations || !a->colors0 || (a->n_sets == 2 && !a->colors1))) {
        gsd_set_error("null input pointer");
        return GSD_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    GsdGeomWs g; GsdBinWs b; GsdImgWs im;
    const int tiles = cam.gx * cam.gy;
    if ((rc = gsd_carve_geom(a->G, a->geom_ws, &g))) return rc;
    if ((rc = gsd_carve_bin(a->G, a->capacity, tiles, a->binning_ws, &b))) return rc;
    if ((rc = gsd_carve_img(a->W, a->H, a->n_sets, b.max_items, a->image_ws, &im))) return rc;
    GSD_CUDA_CHECK(cudaMemsetAsync(a->status, 0, GSD_STATUS_WORDS * 4, st));
    if ((rc = gsd_launch_preprocess(a->G, cam, a, g, b.tile_total, tiles, st))) return rc;
    if ((rc = gsd_launch_binning(a->G, cam, a, g, b, im.chunk_flags, b.max_items * 8, st))) return rc;
    GsdRenderParams p;
    memset(&p, 0, sizeof(p));
    p.ranges = b.ranges;
    p.planes = b.records;
    p.plane_stride = a->capacity > 0 ? a->capacity : 1;
    p.W = a->W; p.H = a->H; p.gx = cam.gx; p.n_tiles = tiles;
    p.chunk_ptr = b.chunk_ptr; p.item_tile = b.item_tile; p.n_items = b.counters + 0; p.max_items = b.max_items;
    p.chunk_state = im.chunk_state; p.term_state = im.term_state;
    p.bg0 = a->bg0; p.bg1 = a->n_sets == 2 ? a->bg1 : nullptr;
    p.out_color = a->out_color;
    p.out_depth = a->out_depth;
    p.final_T = im.final_T;
    p.n_contrib = im.n_contrib;
    p.keys = b.keys; p.g_xy = g.xy; p.g_conic_o = g.conic_o; p.g_ext = g.ext; p.g_depth = g.depth; p.g_rect = g.rect;
    p.g_slot_base = g.slot_base; p.colors0 = a->colors0; p.colors1 = a->n_sets == 2 ? a->colors1 : nullptr;
    p.planes_w = b.records;
    p.exec_item = b.exec_item; p.chunk_flags = im.chunk_flags; p.replay_count = b.counters + 4;
    return gsd_launch_render_fwd(p, tiles, a->n_sets, st);
}

static int raster_backward_impl(const GsdRasterBwd *a, void *stream, int stages, const GsdTrackUpdate *fused = nullptr);
extern "C" int gsd_raster_backward(const GsdRasterBwd *a, void *stream) { return raster_backward_impl(a, stream, 3); }
// steady-state tracking: blend backward (geometry-only) + ONE per-Gaussian kernel that turns the partial records into the
// gradients of means3D / rotations and applies gsd_track_update to them in registers (no gradient arrays, one launch less)
extern "C" int gsd_track_backward_update(const GsdRasterBwd *a, const GsdTrackUpdate *u, void *stream) {
    if (!a || !u) { gsd_set_error("null descriptor"); return GSD_ERR_INVALID; }
    if (a->dL_dcolors0 || a->dL_dcolors1 || a->dL_dopacities || a->dL_dmeans2D) {
        gsd_set_error("gsd_track_backward_update is the geometry-only backward: colour / opacity / means2D gradient outputs must be NULL");
        return GSD_ERR_INVALID;
    }
    if (u->G != a->fwd.G) { gsd_set_error("GsdTrackUpdate.G differs from the rasterizer's G"); return GSD_ERR_INVALID; }
    if (u->G > 0 && (!u->means3D || !u->unnorm_rotations || !u->m_means || !u->v_means || !u->m_rot || !u->v_rot || !u->step_means ||
                     !u->step_rot || !u->block_counter)) {
        gsd_set_error("null pointer in GsdTrackUpdate (block_counter is required by the fused kernel)");
        return GSD_ERR_INVALID;
    }
    if (u->radii && !u->max_2D_radius) { gsd_set_error("radii given without max_2D_radius"); return GSD_ERR_INVALID; }
    return raster_backward_impl(a, stream, 3, u);
}
// stage 1: blend backward only (the dominant kernel, timed alone for the roofline); stage 2: per-Gaussian backward only
// stage 4: the per-chunk prefix pass of the blend backward alone — it needs only the forward's state (no dL_dcolor, no partial_ws),
// so a caller can run it on another stream while the image gradient is still being computed, and set prefix_done for the rest
extern "C" int gsd_raster_backward_stage(const GsdRasterBwd *a, int32_t stage, void *stream) {
    if (stage != 1 && stage != 2 && stage != 4) { gsd_set_error("stage must be 1, 2 or 4"); return GSD_ERR_INVALID; }
    return raster_backward_impl(a, stream, stage);
}
static int raster_backward_impl(const GsdRasterBwd *a, void *stream, int stages, const GsdTrackUpdate *fused) {
    if (!a) { gsd_set_error("null descriptor"); return GSD_ERR_INVALID; }
    const GsdRasterFwd *f = &a->fwd;
    GsdCam cam;
    int rc;
    if ((rc = make_cam(f, &cam))) return rc;
    if (!f->geom_ws || !f->binning_ws || !f->image_ws || (stages != 4 && (!a->partial_ws || !a->dL_dcolor)) ||
        ((stages & 2) && !fused && f->G > 0 && (!a->dL_dmeans3D || !a->dL_dscales || !a->dL_drotations))) {
        gsd_set_error("null workspace/output pointer");
        return GSD_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    GsdGeomWs g; GsdBinWs b; GsdImgWs im;
    const int tiles = cam.gx * cam.gy;
    if ((rc = gsd_carve_geom(f->G, f->geom_ws, &g))) return rc;
    if ((rc = gsd_carve_bin(f->G, f->capacity, tiles, f->binning_ws, &b))) return rc;
    if ((rc = gsd_carve_img(f->W, f->H, f->n_sets, b.max_items, f->image_ws, &im))) return rc;
    GsdRenderParams p;
    memset(&p, 0, sizeof(p));
    p.ranges = b.ranges;
    p.planes = b.records;
    p.plane_stride = f->capacity > 0 ? f->capacity : 1;
    p.W = f->W; p.H = f->H; p.gx = cam.gx; p.n_tiles = tiles;
    p.chunk_ptr = b.chunk_ptr; p.item_tile = b.item_tile; p.n_items = b.counters + 0; p.max_items = b.max_items;
    p.chunk_state = im.chunk_state; p.term_state = im.term_state;
    p.bg0 = f->bg0; p.bg1 = f->n_sets == 2 ? f->bg1 : nullptr;
    p.out_color = f->out_color;
    p.out_depth = f->out_depth;
    p.final_T = im.final_T;
    p.n_contrib = im.n_contrib;
    p.dL_dcolor = a->dL_dcolor;
    p.exec_item = b.exec_item;
    p.partials = (float *)a->partial_ws;
    // colours and opacities frozen (no output requested): geometry-only partials
    const int geom_only = (!a->dL_dcolors0 && !a->dL_dcolors1 && !a->dL_dopacities) ? 1 : 0;
    p.geom_only = geom_only;
    if (stages == 4) return (f->G > 0 && f->capacity > 0) ? gsd_launch_render_bwd(p, tiles, f->n_sets, 1, st) : GSD_OK;
    if ((stages & 1) && f->G > 0 && f->capacity > 0)
        if ((rc = gsd_launch_render_bwd(p, tiles, f->n_sets, a->prefix_done ? 2 : 3, st))) return rc;
    if ((stages & 2) && fused) return gsd_launch_preprocess_bwd_update(f->G, cam, a, g, *fused, st);
    if (stages & 2) return gsd_launch_preprocess_bwd(f->G, cam, a, g, geom_only, st);
    return GSD_OK;
}

extern "C" int gsd_raster_mark_visible(int32_t G, const float *means3D, const float *viewmatrix, uint8_t *visible, void *stream) {
    if (G < 0 || !viewmatrix || (G > 0 && (!means3D || !visible))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    GsdCam cam;
    memset(&cam, 0, sizeof(cam));
    cam.view = viewmatrix;
    return gsd_launch_mark_visible(G, cam, means3D, visible, (cudaStream_t)stream);
}
