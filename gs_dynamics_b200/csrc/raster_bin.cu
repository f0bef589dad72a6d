// raster_bin.cu — tile binning: prefix sum -> (tile | depth) keys -> radix sort -> per-tile ranges and
// packed 64-byte per-instance records that the blend kernels stream with 1-D bulk (TMA) copies.
//
// Replaces InclusiveSum / duplicateWithKeys / SortPairs / identifyTileRanges of the upstream rasterizer
// (SURVEY.md §2.1). Differences by design: the instance count stays on the device (no D2H sync), and the
// sorted list is materialised as contiguous records so that fwd and bwd blend read linear memory.
#include <cub/cub.cuh>

#include "common.cuh"

// ---- workspace carving -----------------------------------------------------------------------------
int gsd_carve_geom(int G, void *base, GsdGeomWs *ws) {
    size_t n = (size_t)(G > 0 ? G : 1);
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? (void *)(p + off) : nullptr;
        off += gsd_align_up(bytes);
        return r;
    };
    ws->xy = (float2 *)take(n * sizeof(float2));
    ws->conic_o = (float4 *)take(n * sizeof(float4));
    ws->ext = (float2 *)take(n * sizeof(float2));
    ws->depth = (float *)take(n * sizeof(float));
    ws->rect = (uint2 *)take(n * sizeof(uint2));
    ws->tiles = (uint32_t *)take(n * sizeof(uint32_t));
    ws->offsets = (uint32_t *)take(n * sizeof(uint32_t));
    size_t tmp = 0;
    cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    if (e != cudaSuccess) {
        gsd_set_error("cub scan size query failed: %s", cudaGetErrorString(e));
        return GSD_ERR_CUDA;
    }
    ws->scan_tmp_bytes = tmp;
    ws->scan_tmp = take(tmp);
    ws->total = off;
    return GSD_OK;
}

int gsd_carve_bin(int64_t capacity, int tiles, void *base, GsdBinWs *ws) {
    size_t n = (size_t)(capacity > 0 ? capacity : 1);
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? (void *)(p + off) : nullptr;
        off += gsd_align_up(bytes);
        return r;
    };
    ws->keys_a = (uint64_t *)take(n * 8);
    ws->keys_b = (uint64_t *)take(n * 8);
    ws->vals_a = (uint32_t *)take(n * 4);
    ws->vals_b = (uint32_t *)take(n * 4);
    ws->ranges = (uint2 *)take((size_t)tiles * sizeof(uint2));
    ws->records = (float4 *)take(n * GSD_REC_FLOATS * 4);
    size_t tmp = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                                    (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 64);
    if (e != cudaSuccess) {
        gsd_set_error("cub sort size query failed: %s", cudaGetErrorString(e));
        return GSD_ERR_CUDA;
    }
    ws->sort_tmp_bytes = tmp;
    ws->sort_tmp = take(tmp);
    ws->total = off;
    return GSD_OK;
}

int gsd_carve_img(int W, int H, void *base, GsdImgWs *ws) {
    size_t n = (size_t)W * H;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? (void *)(p + off) : nullptr;
        off += gsd_align_up(bytes);
        return r;
    };
    ws->final_T = (float *)take(n * 4);
    ws->n_contrib = (int32_t *)take(n * 4);
    ws->total = off;
    return GSD_OK;
}

// ---- kernels ------------------------------------------------------------------------------------------
// status[0] = R, status[1] = overflow, status[2] = n_visible (filled by the count kernel)
__global__ void gsd_finish_count_kernel(int G, const uint32_t *__restrict__ offsets, int64_t capacity,
                                        int32_t *__restrict__ status) {
    uint32_t R = G > 0 ? offsets[G - 1] : 0u;
    status[0] = (int32_t)R;
    status[1] = ((int64_t)R > capacity) ? 1 : 0;
}

// one thread per Gaussian: emit (tile<<32 | depth bits, gaussian id) for every touched tile, row-major
__global__ void __launch_bounds__(256)
gsd_duplicate_keys_kernel(int G, int gx, const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ tiles,
                          const uint2 *__restrict__ rect, const float *__restrict__ depth, int64_t capacity,
                          uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    uint32_t n = tiles[i];
    if (n == 0) return;
    uint32_t off = offsets[i] - n;
    uint2 rc = rect[i];
    int minx = rc.x & 0xffff, miny = rc.x >> 16, maxx = rc.y & 0xffff, maxy = rc.y >> 16;
    uint32_t dbits = __float_as_uint(depth[i]);
    for (int y = miny; y < maxy; ++y)
        for (int x = minx; x < maxx; ++x) {
            if ((int64_t)off < capacity) {
                keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                vals[off] = (uint32_t)i;
            }
            ++off;
        }
}

// slots [R, capacity) get a sentinel key that sorts behind every real key
__global__ void gsd_fill_sentinel_kernel(int64_t capacity, const int32_t *__restrict__ status, uint64_t *__restrict__ keys,
                                         uint32_t *__restrict__ vals) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= capacity) return;
    int64_t R = (uint32_t)status[0];
    if (j >= R) {
        keys[j] = ~0ull;
        vals[j] = 0xffffffffu;
    }
}

// one thread per sorted instance: tile range boundaries + the packed record
//   plane0 = (x, y, ext_x, ext_y)   plane1 = (A, B, C, opacity)
//   plane2 = (c0, c1, c2, depth)    plane3 = (slot bits, c3, c4, c5)
// slot = position of this instance in the *unsorted* per-Gaussian order (offset + rank of the tile inside the
// Gaussian's rectangle): the backward writes its partial gradient there so that a Gaussian's partials are
// contiguous and can be summed in a fixed order without atomics.
__global__ void __launch_bounds__(256)
gsd_pack_records_kernel(int64_t capacity, int gx, const int32_t *__restrict__ status, const uint64_t *__restrict__ keys,
                        const uint32_t *__restrict__ vals, const float2 *__restrict__ xy, const float4 *__restrict__ conic_o,
                        const float2 *__restrict__ ext, const float *__restrict__ depth, const uint2 *__restrict__ rect,
                        const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ tiles,
                        const float *__restrict__ colors0, const float *__restrict__ colors1, uint2 *__restrict__ ranges,
                        float4 *__restrict__ records) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t R = (uint32_t)status[0];
    if (R > capacity) R = capacity;
    if (j >= R) return;
    uint64_t key = keys[j];
    uint32_t tile = (uint32_t)(key >> 32);
    if (j == 0) {
        ranges[tile].x = 0;
    } else {
        uint32_t prev = (uint32_t)(keys[j - 1] >> 32);
        if (prev != tile) {
            ranges[prev].y = (uint32_t)j;
            ranges[tile].x = (uint32_t)j;
        }
    }
    if (j == R - 1) ranges[tile].y = (uint32_t)R;

    uint32_t g = vals[j];
    uint2 rc = rect[g];
    int minx = rc.x & 0xffff, miny = rc.x >> 16, maxx = rc.y & 0xffff;
    int tx = tile % gx, ty = tile / gx;
    uint32_t slot = offsets[g] - tiles[g] + (uint32_t)((ty - miny) * (maxx - minx) + (tx - minx));
    float2 p = xy[g];
    float2 e = ext[g];
    float4 co = conic_o[g];
    float d = depth[g];
    float c0 = colors0[3 * g], c1 = colors0[3 * g + 1], c2 = colors0[3 * g + 2];
    float c3 = 0.f, c4 = 0.f, c5 = 0.f;
    if (colors1) {
        c3 = colors1[3 * g];
        c4 = colors1[3 * g + 1];
        c5 = colors1[3 * g + 2];
    }
    // four SoA planes of [capacity] float4 (each plane is one contiguous bulk-copy source per tile batch)
    records[j] = make_float4(p.x, p.y, e.x, e.y);
    records[capacity + j] = co;
    records[2 * capacity + j] = make_float4(c0, c1, c2, d);
    records[3 * capacity + j] = make_float4(__uint_as_float(slot), c3, c4, c5);
}

// ---- host launchers -------------------------------------------------------------------------------------
int gsd_launch_scan(int G, const GsdGeomWs &g, int64_t capacity, int32_t *status, cudaStream_t st) {
    if (G > 0) {
        size_t tmp = g.scan_tmp_bytes;
        GSD_CUDA_CHECK(cub::DeviceScan::InclusiveSum(g.scan_tmp, tmp, g.tiles, g.offsets, G, st));
        gsd_count_launch(0, 2);
    }
    gsd_finish_count_kernel<<<1, 1, 0, st>>>(G, g.offsets, capacity, status);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

static int highest_bit(uint32_t n) {
    int b = 0;
    while (n) { ++b; n >>= 1; }
    return b;
}

int gsd_launch_binning(int G, const GsdCam &cam, const GsdRasterFwd *a, const GsdGeomWs &g, const GsdBinWs &b,
                       cudaStream_t st) {
    const int64_t cap = a->capacity;
    const int tiles = cam.gx * cam.gy;
    GSD_CUDA_CHECK(cudaMemsetAsync(b.ranges, 0, (size_t)tiles * sizeof(uint2), st));
    if (G == 0 || cap == 0) return GSD_OK;
    gsd_fill_sentinel_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(cap, a->status, b.keys_a, b.vals_a);
    GSD_LAUNCH_CHECK();
    gsd_duplicate_keys_kernel<<<(G + 255) / 256, 256, 0, st>>>(G, cam.gx, g.offsets, g.tiles, g.rect, g.depth, cap,
                                                               b.keys_a, b.vals_a);
    GSD_LAUNCH_CHECK();
    size_t tmp = b.sort_tmp_bytes;
    int end_bit = 32 + highest_bit((uint32_t)tiles); // sentinel tile id (all ones) stays above every real tile
    GSD_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, b.keys_a, b.keys_b, b.vals_a, b.vals_b, (int)cap, 0,
                                                   end_bit, st));
    gsd_count_launch(0, 2 + (end_bit + 7) / 8);
    gsd_pack_records_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(
        cap, cam.gx, a->status, b.keys_b, b.vals_b, g.xy, g.conic_o, g.ext, g.depth, g.rect, g.offsets, g.tiles,
        a->colors0, a->n_sets == 2 ? a->colors1 : nullptr, b.ranges, b.records);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
