// raster_bin.cu — tile binning without a global sort:
//   pass 1 (in the preprocess kernel)  per-tile instance counts + per-Gaussian slot ranges
//   tile_scan     one CTA: exclusive scan of the tile counts -> per-tile ranges, and the tile schedule sorted by
//                 descending count (longest-processing-time-first) for the persistent blend kernels
//   fill          one thread per Gaussian: append (depth bits << 32 | id) to each touched tile's segment
//   sort_pack     one CTA per tile: bitonic sort of the segment in shared memory (ties by Gaussian id, i.e. exactly the
//                 order of the reference's stable (tile | depth) radix sort), then gather the Gaussian data into the four
//                 SoA record planes that the blend kernels stream with 1-D bulk (TMA) copies
//
// Replaces InclusiveSum / duplicateWithKeys / 6-pass SortPairs / identifyTileRanges of the upstream rasterizer
// (SURVEY.md §2.1): the R-element 43-bit global radix sort (75 us at R = 181k on B200, profiles/r1_*) becomes
// per-tile sorts of ~150 keys that run concurrently in shared memory, and the instance count never leaves the device.
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
    ws->slot_base = (uint32_t *)take(n * sizeof(uint32_t));
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
    // the three small integer arrays are contiguous so that one memset clears them
    ws->tile_count = (int32_t *)take((size_t)tiles * 4);
    ws->tile_fill = (int32_t *)take((size_t)tiles * 4);
    ws->counters = (int32_t *)take(8 * 4);
    ws->tile_order = (int32_t *)take((size_t)tiles * 4);
    ws->ranges = (uint2 *)take((size_t)tiles * sizeof(uint2));
    ws->keys = (uint64_t *)take(n * 8);
    ws->records = (float4 *)take(n * GSD_REC_FLOATS * 4);
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

// ---- tile scan + schedule ----------------------------------------------------------------------------
// One CTA of 1024 threads. status[0] already holds R (accumulated by the preprocess kernel).
#define SCAN_THREADS 1024
#define ORDER_MAX 8192 // tiles sortable in shared memory (64 KB of keys); larger images keep the natural order

__global__ void __launch_bounds__(SCAN_THREADS)
gsd_tile_scan_kernel(int n_tiles, int64_t capacity, const int32_t *__restrict__ tile_count, uint2 *__restrict__ ranges,
                     int32_t *__restrict__ tile_order, int32_t *__restrict__ status) {
    extern __shared__ unsigned long long skey[]; // [pow2 >= n_tiles] when n_tiles <= ORDER_MAX
    __shared__ int sbuf[SCAN_THREADS];
    __shared__ int carry;
    const int t = threadIdx.x;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
        const int i = base + t;
        const int v = i < n_tiles ? tile_count[i] : 0;
        sbuf[t] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_THREADS; o <<= 1) {
            int add = t >= o ? sbuf[t - o] : 0;
            __syncthreads();
            sbuf[t] += add;
            __syncthreads();
        }
        if (i < n_tiles) {
            long long s = (long long)carry + sbuf[t] - v, e = s + v;
            if (s > capacity) s = capacity;
            if (e > capacity) e = capacity;
            ranges[i] = make_uint2((uint32_t)s, (uint32_t)e);
        }
        __syncthreads();
        if (t == SCAN_THREADS - 1) carry += sbuf[SCAN_THREADS - 1];
        __syncthreads();
    }
    if (t == 0) status[1] = ((long long)(uint32_t)status[0] > capacity) ? 1 : 0;

    // schedule: tiles by descending count (ties by tile id) — bitonic sort of (~count << 32 | tile)
    if (n_tiles <= ORDER_MAX) {
        int P = 1;
        while (P < n_tiles) P <<= 1;
        for (int i = t; i < P; i += SCAN_THREADS)
            skey[i] = i < n_tiles ? (((unsigned long long)(0xffffffffu - (uint32_t)tile_count[i]) << 32) | (uint32_t)i) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j >= 1; j >>= 1) {
                for (int i = t; i < P; i += SCAN_THREADS) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        unsigned long long a = skey[i], b = skey[ixj];
                        bool up = (i & k) == 0;
                        if ((a > b) == up) { skey[i] = b; skey[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = t; i < n_tiles; i += SCAN_THREADS) tile_order[i] = (int32_t)(skey[i] & 0xffffffffu);
    } else {
        for (int i = t; i < n_tiles; i += SCAN_THREADS) tile_order[i] = i;
    }
}

// ---- fill ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gsd_fill_keys_kernel(int G, int gx, int64_t capacity, const uint32_t *__restrict__ tiles, const uint2 *__restrict__ rect,
                     const float *__restrict__ depth, const uint2 *__restrict__ ranges, int32_t *__restrict__ tile_fill,
                     uint64_t *__restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    if (tiles[i] == 0) return;
    uint2 rc = rect[i];
    int minx = rc.x & 0xffff, miny = rc.x >> 16, maxx = rc.y & 0xffff, maxy = rc.y >> 16;
    const uint64_t key = ((uint64_t)__float_as_uint(depth[i]) << 32) | (uint32_t)i;
    for (int y = miny; y < maxy; ++y)
        for (int x = minx; x < maxx; ++x) {
            const int tile = y * gx + x;
            const uint2 r = ranges[tile];
            const uint32_t pos = r.x + (uint32_t)atomicAdd(&tile_fill[tile], 1);
            if (pos < r.y) keys[pos] = key; // r.y is clipped to capacity
        }
}

// ---- per-tile sort + pack ----------------------------------------------------------------------------------
// Bitonic network in its "mirror" form: every compare-exchange moves the smaller key to the lower index, so the
// virtual +inf padding above n never moves and any n (not only powers of two) sorts correctly.
template <typename KeyPtr>
__device__ __forceinline__ void cta_bitonic_sort(KeyPtr key, int n, int tid, int nthreads) {
    int P = 1;
    while (P < n) P <<= 1;
    const int half = P >> 1;
    for (int k = 2; k <= P; k <<= 1) {
        const int hk = k >> 1;
        for (int i = tid; i < half; i += nthreads) {
            const int blk = i / hk, off = i % hk;
            const int a = blk * k + off, b = blk * k + k - 1 - off;
            if (b < n) {
                unsigned long long x = key[a], y = key[b];
                if (x > y) { key[a] = y; key[b] = x; }
            }
        }
        __syncthreads();
        for (int j = hk >> 1; j >= 1; j >>= 1) {
            for (int i = tid; i < half; i += nthreads) {
                const int a = (i / j) * 2 * j + (i % j), b = a + j;
                if (b < n) {
                    unsigned long long x = key[a], y = key[b];
                    if (x > y) { key[a] = y; key[b] = x; }
                }
            }
            __syncthreads();
        }
    }
}

#define SORT_THREADS 256
#define SORT_SMEM_KEYS 6000 // 48 KB static limit; longer tile lists are sorted in place in global memory

// plane0 = (x, y, ext_x, ext_y)   plane1 = (A, B, C, opacity)   plane2 = (c0, c1, c2, depth)   plane3 = (slot bits, c3, c4, c5)
// slot = slot_base[g] + rank of the tile inside the Gaussian's rectangle: the blend backward writes this instance's
// partial gradient there, so a Gaussian's partials are contiguous and are summed in a fixed order without atomics.
__global__ void __launch_bounds__(SORT_THREADS)
gsd_tile_sort_pack_kernel(int n_tiles, int gx, int64_t capacity, const int32_t *__restrict__ tile_order,
                          int32_t *__restrict__ next_tile, const uint2 *__restrict__ ranges, uint64_t *__restrict__ keys,
                          const float2 *__restrict__ xy, const float4 *__restrict__ conic_o, const float2 *__restrict__ ext,
                          const float *__restrict__ depth, const uint2 *__restrict__ rect, const uint32_t *__restrict__ slot_base,
                          const float *__restrict__ colors0, const float *__restrict__ colors1, float4 *__restrict__ records) {
    __shared__ unsigned long long skeys[SORT_SMEM_KEYS];
    __shared__ int s_next;
    const int t = threadIdx.x;
    for (;;) {
        if (t == 0) s_next = atomicAdd(next_tile, 1);
        __syncthreads();
        const int q = s_next;
        __syncthreads();
        if (q >= n_tiles) break;
        const int tile = tile_order[q];
        const uint2 r = ranges[tile];
        const int n = (int)(r.y - r.x);
        if (n == 0) continue; // (with the LPT schedule everything after is empty too; natural order needs the scan)
        unsigned long long *gk = reinterpret_cast<unsigned long long *>(keys) + r.x;
        const bool in_smem = n <= SORT_SMEM_KEYS;
        if (in_smem) {
            for (int i = t; i < n; i += SORT_THREADS) skeys[i] = gk[i];
            __syncthreads();
            cta_bitonic_sort(skeys, n, t, SORT_THREADS);
        } else {
            cta_bitonic_sort(gk, n, t, SORT_THREADS);
        }
        const int tx = tile % gx, ty = tile / gx;
        for (int i = t; i < n; i += SORT_THREADS) {
            const uint32_t g = (uint32_t)((in_smem ? skeys[i] : gk[i]) & 0xffffffffull);
            const uint2 rc = rect[g];
            const int minx = rc.x & 0xffff, miny = rc.x >> 16, maxx = rc.y & 0xffff;
            const uint32_t slot = slot_base[g] + (uint32_t)((ty - miny) * (maxx - minx) + (tx - minx));
            const float2 p = xy[g];
            const float2 e = ext[g];
            const float4 co = conic_o[g];
            const float d = depth[g];
            const float c0 = colors0[3 * g], c1 = colors0[3 * g + 1], c2 = colors0[3 * g + 2];
            float c3 = 0.f, c4 = 0.f, c5 = 0.f;
            if (colors1) { c3 = colors1[3 * g]; c4 = colors1[3 * g + 1]; c5 = colors1[3 * g + 2]; }
            const int64_t j = (int64_t)r.x + i;
            records[j] = make_float4(p.x, p.y, e.x, e.y);
            records[capacity + j] = co;
            records[2 * capacity + j] = make_float4(c0, c1, c2, d);
            records[3 * capacity + j] = make_float4(__uint_as_float(slot), c3, c4, c5);
        }
        __syncthreads();
    }
}

// ---- host launchers -------------------------------------------------------------------------------------
// b.tile_count / tile_fill / counters must be zero on entry (one memset by the caller); status[0] holds R.
int gsd_launch_binning(int G, const GsdCam &cam, const GsdRasterFwd *a, const GsdGeomWs &g, const GsdBinWs &b,
                       cudaStream_t st) {
    const int64_t cap = a->capacity;
    const int tiles = cam.gx * cam.gy;
    size_t smem = 0;
    if (tiles <= ORDER_MAX) {
        int P = 1;
        while (P < tiles) P <<= 1;
        smem = (size_t)P * 8;
    }
    static bool attr_set = false;
    if (!attr_set) {
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_tile_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ORDER_MAX * 8));
        attr_set = true;
    }
    gsd_tile_scan_kernel<<<1, SCAN_THREADS, smem, st>>>(tiles, cap, b.tile_count, b.ranges, b.tile_order, a->status);
    GSD_LAUNCH_CHECK();
    if (G == 0 || cap == 0) return GSD_OK;
    gsd_fill_keys_kernel<<<(G + 255) / 256, 256, 0, st>>>(G, cam.gx, cap, g.tiles, g.rect, g.depth, b.ranges, b.tile_fill, b.keys);
    GSD_LAUNCH_CHECK();
    int grid = tiles < 148 * 8 ? tiles : 148 * 8;
    gsd_tile_sort_pack_kernel<<<grid, SORT_THREADS, 0, st>>>(tiles, cam.gx, cap, b.tile_order, b.counters + 0, b.ranges, b.keys,
                                                              g.xy, g.conic_o, g.ext, g.depth, g.rect, g.slot_base, a->colors0,
                                                              a->n_sets == 2 ? a->colors1 : nullptr, b.records);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
