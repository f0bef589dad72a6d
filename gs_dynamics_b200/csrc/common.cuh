// common.cuh — shared device helpers for libgsd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gsd.h"

#define GSD_TILE 16
#define GSD_REC_FLOATS 16  // one packed per-tile-instance record = 64 B
#define GSD_PART_FLOATS 16 // one backward partial-gradient record = 64 B

void gsd_set_error(const char *fmt, ...);
void gsd_count_launch(int own, int library); // host-side launch accounting (gsd_launch_count)

#define GSD_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            gsd_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GSD_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define GSD_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            gsd_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GSD_ERR_CUDA;                                                               \
        }                                                                                      \
        gsd_count_launch(1, 0);                                                                \
    } while (0)

static inline size_t gsd_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- camera constants passed by value to kernels -------------------------------------------------
struct GsdCam {
    const float *view; // device [16]
    const float *proj; // device [16]
    float tanfovx, tanfovy, focal_x, focal_y, scale_modifier;
    int W, H, gx, gy;
};

// ---- workspace carving (host) ----------------------------------------------------------------------
struct GsdGeomWs { // per-Gaussian state
    float2 *xy;        // pixel centre
    float4 *conic_o;   // conic a,b,c + opacity
    float2 *ext;       // conservative half extents of the alpha>=1/255 ellipse (pixels)
    float *depth;      // view-space z
    uint2 *rect;       // (minx | miny<<16, maxx | maxy<<16) in tiles, max exclusive
    uint32_t *tiles;   // tiles touched
    uint32_t *slot_base; // first partial-gradient slot of this Gaussian's instances (= its position in Gaussian order)
    uint32_t *block_sum; // [ceil(G/256)] instances per preprocess block
    uint32_t *block_base;// [ceil(G/256)] exclusive scan of block_sum
    size_t total;
};
#ifndef GSD_CHUNK
#define GSD_CHUNK 128      // records per blend work item (tile lists are split into chunks processed in parallel)
#endif
#ifndef GSD_BIN_BLOCK
#define GSD_BIN_BLOCK 1024 // Gaussians per binning block
#endif
struct GsdBinWs {
    int n_bb;            // binning blocks = ceil(G / GSD_BIN_BLOCK)
    int max_items;       // upper bound of blend work items = capacity / GSD_CHUNK + tiles
    int max_units;       // upper bound of sort units (1024-key segments) = capacity / 1024 + tiles
    int32_t *table;      // [tiles][n_bb] per (tile, binning block) instance counts -> exclusive scan (tile-major)
    int32_t *tile_total; // [tiles] instances per tile (zero-filled by the preprocess kernel, accumulated by the histogram pass)
    int32_t *tile_base;  // [tiles] exclusive scan of the totals (first slot of the tile's segment)
    uint2 *ranges;       // per tile [start,end) clipped to capacity
    int32_t *chunk_ptr;  // [tiles+1] exclusive scan of chunks per tile
    int4 *item_tile;     // [max_items] work item -> (tile, chunk index in the tile, first record, record count): one 16-byte read per CTA
    int32_t *counters;   // [8] 0: n_items, 1: sort units, 2: tiles with more than one sort unit, 3: "tile bases published" flag, 4: replay pairs of the blend forward
    int32_t *unit_tile;  // [max_units] tile of each sort unit
    int32_t *unit_seg;   // [max_units] segment index of each sort unit inside its tile
    int32_t *long_tile;  // [tiles] tiles whose list spans several sort units
    int32_t *exec_item;  // [max_items] blend work items in execution order: chunk index major (see gsd_blend_fwd_chunk_kernel)
    uint64_t *keys;      // [capacity] (depth bits << 32 | gaussian id), grouped by tile, unsorted inside a tile
    uint64_t *keys_tmp;  // [capacity] merge-sort ping-pong buffer for tile lists that do not fit shared memory
    float4 *records;     // 4 SoA planes of [capacity] float4: packed per-instance records sorted by (tile, depth, id)
    size_t total;
};
struct GsdImgWs {
    float *final_T;
    int32_t *n_contrib;
    float *chunk_state; // [max_items][5 + 3 n_sets][256]
    float *term_state;  // [tiles][4 + 3 n_sets][256]
    int32_t *chunk_flags; // [max_items][8] "chunk composite published" per (work item, 8x4-pixel rectangle); cleared every forward
    size_t total;
};
size_t gsd_chunk_state_floats(int n_sets, int max_items);
size_t gsd_term_state_floats(int n_sets, int tiles);

int gsd_carve_geom(int G, void *base, GsdGeomWs *ws);
int gsd_carve_bin(int G, int64_t capacity, int tiles, void *base, GsdBinWs *ws);
int gsd_carve_img(int W, int H, int n_sets, int max_items, void *base, GsdImgWs *ws);

struct GsdRenderParams {
    const uint2 *ranges;
    const float4 *planes; // 4 planes of [plane_stride] float4
    int64_t plane_stride;
    int W, H, gx, n_tiles;
    const int32_t *chunk_ptr;  // [tiles+1]
    const int4 *item_tile;     // [n_items] (tile, chunk, first record, record count)
    const int32_t *n_items;    // device scalar
    const int32_t *exec_item;  // [n_items] execution order of the forward chunk kernel
    int32_t *chunk_flags;      // [max_items][8] pass A: look-back flags; pass B -> C: list of (item * 8 + rectangle) replay pairs
    int32_t *replay_count;     // device scalar: entries of that list (cleared by the histogram pass of every forward)
    // forward A1 gathers the per-Gaussian data by sorted key and writes the record planes
    const uint64_t *keys; const float2 *g_xy; const float4 *g_conic_o; const float2 *g_ext; const float *g_depth;
    const uint2 *g_rect; const uint32_t *g_slot_base; const float *colors0; const float *colors1;
    float4 *planes_w;
    float *chunk_state;        // per work item: SoA fields x 256 pixels (see raster_render.cu)
    float *term_state;         // per tile: terminal record of each pixel
    int max_items;
    int geom_only;             // backward: only mean2D / conic partials (colours and opacities frozen)
    const float *bg0, *bg1; // device [3] each; bg1 may be null
    float *out_color; // [CH,H,W]
    float *out_depth; // [H,W]
    float *final_T;
    int32_t *n_contrib;
    // backward only
    const float *dL_dcolor;
    float *partials; // [capacity][GSD_PART_FLOATS]
};

#ifdef __CUDACC__
// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// Every kernel of the tracking iteration starts with gsd_pdl_wait() (griddepcontrol.wait: all memory operations of the
// preceding kernel are complete and visible — it is executed before ANY global-memory access, so the data dependences and
// write-after-read hazards between consecutive launches are exactly those of ordinary stream order) followed by
// gsd_pdl_launch() (the next kernel's CTAs may be made resident as SM slots free up).  Launched through gsd_launch() with
// the programmatic-stream-serialization attribute, a kernel's launch latency, CTA dispatch and prologue overlap the tail of
// its predecessor instead of following it; inside a captured CUDA graph these become programmatic edges.  GSD_NO_PDL=1 in
// the environment turns the attribute off (plain stream order).
// F.normalize of a quaternion (helpers.py:40; eps 1e-12), written with explicit roundings so that every kernel that normalises
// on the fly (preprocess, the priors' node records, the update's normalize-backward) gets the same bits whatever its -fmad setting
__device__ __forceinline__ float gsd_quat_norm(float4 v) {
    float s = __fmul_rn(v.x, v.x);
    s = __fmaf_rn(v.y, v.y, s);
    s = __fmaf_rn(v.z, v.z, s);
    s = __fmaf_rn(v.w, v.w, s);
    return fmaxf(__fsqrt_rn(s), 1e-12f);
}
__device__ __forceinline__ float4 gsd_quat_normalize(float4 v) {
    const float n = gsd_quat_norm(v);
    return make_float4(__fdiv_rn(v.x, n), __fdiv_rn(v.y, n), __fdiv_rn(v.z, n), __fdiv_rn(v.w, n));
}

__device__ __forceinline__ void gsd_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef GSD_PDL_EARLY
#define GSD_PDL_EARLY 0
#endif
__device__ __forceinline__ void gsd_pdl_launch() {
#if GSD_PDL_EARLY
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// unconditional early trigger: used by the kernels of the GNN step (one or few waves of long-lived CTAs), where letting the next
// kernel's CTAs start their prologue on SMs as they free up hides ~2-3 us of launch latency + set-up per kernel of the chain
__device__ __forceinline__ void gsd_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool gsd_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline void gsd_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = gsd_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);   // errors are picked up by GSD_LAUNCH_CHECK (cudaGetLastError)
}

// ---- mbarrier / bulk async copy (TMA 1-D) -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on the mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// The one place the blend exponent is evaluated: forward and backward must produce bit-identical
// alpha so that the per-pixel contributor decisions agree. Explicit rn intrinsics stop the compiler from
// contracting the two kernels differently.
__device__ __forceinline__ float gsd_power(float A, float B, float C, float dx, float dy) {
    float a1 = __fmul_rn(A, dx);
    float c1 = __fmul_rn(C, dy);
    float s = __fmaf_rn(c1, dy, __fmul_rn(a1, dx));
    float b1 = __fmul_rn(B, dx);
    return __fmaf_rn(-0.5f, s, -__fmul_rn(b1, dy));
}
// exp(power) as __expf computes it (ex2.approx of power * log2(e)) but with the flush-to-zero form of the instruction: the
// range fix-up of the non-ftz expansion (3 more instructions per evaluation) only matters for results below 2^-126, which are
// alpha = 0 < 1/255 either way.
__device__ __forceinline__ float gsd_gauss(float power) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(power, 1.4426950408889634f)));
    return r;
}
// MUFU.RCP without the range fix-ups of __fdividef (callers guarantee a normal, non-huge argument)
__device__ __forceinline__ float gsd_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Transposed butterfly: N per-lane values are summed over the 32 lanes with ~N shuffles; afterwards the lane with
// holder_id() == k holds the warp total of value k in v[0].
template <int N, int BIT>
__device__ __forceinline__ void xreduce(float *v, int lane) {
    if constexpr (BIT >= 1) {
        constexpr int Hh = (N + 1) / 2;
        const bool upper = (lane & BIT) != 0;
#pragma unroll
        for (int k = 0; k < Hh; ++k) {
            const float lo = v[k];
            const float hi = (Hh + k < N) ? v[Hh + k] : 0.f;
            const float recv = __shfl_xor_sync(0xffffffffu, upper ? lo : hi, BIT);
            v[k] = (upper ? hi : lo) + recv;
        }
        xreduce<Hh, BIT / 2>(v, lane);
    }
}
// Mirrors xreduce's index bookkeeping: every stage halves the (zero-padded) value range [base, base+n) for all lanes
// alike; the lane ends up with value `base`, which is real only if it lies inside the unpadded range.
__device__ __forceinline__ int holder_id(int N, int lane) {
    int base = 0, n = N, end = N;
    for (int bit = 16; bit >= 1; bit >>= 1) {
        const int Hh = (n + 1) / 2;
        if (lane & bit) {
            base += Hh;
        } else {
            end = min(end, base + Hh);
        }
        n = Hh;
    }
    return base < end ? base : -1;
}


#endif
