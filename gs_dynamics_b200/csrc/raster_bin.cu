// raster_bin.cu — tile binning without a global sort and without global atomics:
//   (preprocess)  per-Gaussian tile rectangles + per-256-block instance sums; zero-fills the tile totals
//   hist          one CTA per 1024 Gaussians: shared-memory histogram of touched tiles -> table[tile][block] and tile totals
//   scan          one launch, five roles (see gsd_bin_scan_kernel): slot bases / R; per-tile bases, ranges, blend work items
//                 and sort units; per-tile row scans of the table = scatter bases
//   scatter       one CTA per 1024 Gaussians: append (depth bits << 32 | id) to each touched tile's segment at
//                 base[tile][block] + shared-memory cursor
//   tile_sort     one launch: monotone bucket sort of every tile list in shared memory (keys are unique: ties in depth by
//                 Gaussian id, i.e. exactly the order of the reference's stable (tile | depth) radix sort)
// The blend forward then gathers the Gaussian data by sorted key into the four SoA record planes (raster_render.cu).
//
// Replaces InclusiveSum / duplicateWithKeys / 6-pass SortPairs / identifyTileRanges of the upstream rasterizer
// (SURVEY.md §2.1).  Measured on B200 (profiles/): the 43-bit CUB radix sort of R = 181k keys costs 75 us; per-tile atomic
// cursors are worse (a returning atomic on one address retires every ~18 ns and the benchmark scene puts up to 1900 instances
// on one tile), hence the two-level counting sort with shared-memory histograms.  The instance count never leaves the device.
#include "common.cuh"

#define MAX_TILES_SMEM 12288 // 48 KB of 32-bit bins

// ---- workspace carving -----------------------------------------------------------------------------
int gsd_carve_geom(int G, void *base, GsdGeomWs *ws) {
    size_t n = (size_t)(G > 0 ? G : 1);
    size_t nblk = (n + 255) / 256;
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
    ws->block_sum = (uint32_t *)take(nblk * 4);
    ws->block_base = (uint32_t *)take(nblk * 4);
    ws->total = off;
    return GSD_OK;
}

int gsd_carve_bin(int G, int64_t capacity, int tiles, void *base, GsdBinWs *ws) {
    size_t n = (size_t)(capacity > 0 ? capacity : 1);
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? (void *)(p + off) : nullptr;
        off += gsd_align_up(bytes);
        return r;
    };
    ws->n_bb = (G + GSD_BIN_BLOCK - 1) / GSD_BIN_BLOCK;
    if (ws->n_bb < 1) ws->n_bb = 1;
    ws->max_items = (int)(n / GSD_CHUNK) + tiles;
    ws->max_units = (int)(n / 1024) + tiles;
    ws->table = (int32_t *)take((size_t)tiles * ws->n_bb * 4);
    ws->tile_total = (int32_t *)take((size_t)tiles * 4);
    ws->tile_base = (int32_t *)take((size_t)tiles * 4);
    ws->ranges = (uint2 *)take((size_t)tiles * sizeof(uint2));
    ws->chunk_ptr = (int32_t *)take((size_t)(tiles + 1) * 4);
    ws->item_tile = (int4 *)take((size_t)ws->max_items * 16);
    ws->counters = (int32_t *)take(8 * 4);
    ws->unit_tile = (int32_t *)take((size_t)ws->max_units * 4);
    ws->unit_seg = (int32_t *)take((size_t)ws->max_units * 4);
    ws->long_tile = (int32_t *)take((size_t)tiles * 4);
    ws->exec_item = (int32_t *)take((size_t)ws->max_items * 4);
    ws->keys = (uint64_t *)take(n * 8);
    ws->keys_tmp = (uint64_t *)take(n * 8);
    ws->records = (float4 *)take(n * GSD_REC_FLOATS * 4);
    ws->total = off;
    return GSD_OK;
}

int gsd_carve_img(int W, int H, int n_sets, int max_items, void *base, GsdImgWs *ws) {
    size_t n = (size_t)W * H;
    const int tiles = ((W + GSD_TILE - 1) / GSD_TILE) * ((H + GSD_TILE - 1) / GSD_TILE);
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? (void *)(p + off) : nullptr;
        off += gsd_align_up(bytes);
        return r;
    };
    ws->final_T = (float *)take(n * 4);
    ws->n_contrib = (int32_t *)take(n * 4);
    ws->chunk_state = (float *)take(gsd_chunk_state_floats(n_sets, max_items) * 4);
    ws->term_state = (float *)take(gsd_term_state_floats(n_sets, tiles) * 4);
    ws->chunk_flags = (int32_t *)take((size_t)(max_items > 0 ? max_items : 1) * 8 * 4);
    ws->total = off;
    return GSD_OK;
}

// ---- hist / scatter: one CTA per GSD_BIN_BLOCK Gaussians ------------------------------------------------------
// BIN_THREADS threads per CTA walk the block's Gaussians (the shared-memory histogram stays per GSD_BIN_BLOCK Gaussians).  Smaller
// CTAs would be placed sooner when the tracking iteration's side branch fills the SMs (a 1 024-thread CTA needs a whole SM's worth
// of thread slots to come free at once; DESIGN.md "priors kernel"), but 256 threads walking four Gaussians each measured
// +12 us per iteration, and 256- / 512-Gaussian blocks with one thread per Gaussian (GSD_BIN_BLOCK) +3.5 / -0.5 us: 1 024 stays.
#ifndef BIN_THREADS
#define BIN_THREADS GSD_BIN_BLOCK
#endif
template <bool SCATTER>
__global__ void __launch_bounds__(BIN_THREADS)
gsd_bin_kernel(int G, int gx, int n_tiles, int n_bb, const uint32_t *__restrict__ tiles, const uint2 *__restrict__ rect,
               const float *__restrict__ depth, int32_t *__restrict__ table, const uint2 *__restrict__ ranges,
               uint64_t *__restrict__ keys, uint32_t *__restrict__ slot_base, const uint32_t *__restrict__ block_base,
               int32_t *__restrict__ tile_total, int32_t *__restrict__ counters, int32_t *__restrict__ zero_fill, int n_zero) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    extern __shared__ int bins[]; // [n_tiles] counts (hist) or cursors (scatter)
    const int bb = blockIdx.x;
    for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) bins[i] = 0;
    if (!SCATTER)   // the "published" flags of the blend work items, cleared for this forward pass
        for (int z = bb * BIN_THREADS + threadIdx.x; z < n_zero; z += gridDim.x * BIN_THREADS) zero_fill[z] = 0;
    if (!SCATTER && bb == 0 && threadIdx.x == 0) { counters[3] = 0; counters[4] = 0; }   // "tile bases published" flag of the scan launch; replay-pair count of the blend forward
    __syncthreads();
    for (int gi = threadIdx.x; gi < GSD_BIN_BLOCK; gi += BIN_THREADS) {
    const int i = bb * GSD_BIN_BLOCK + gi;
    if (i < G) {
        const uint32_t n = tiles[i];
        if (SCATTER) slot_base[i] += block_base[i >> 8]; // block-local offset -> global slot
        if (n) {
            const uint2 rc = rect[i];
            const int minx = rc.x & 0xffff, miny = rc.x >> 16, maxx = rc.y & 0xffff, maxy = rc.y >> 16;
            const uint64_t key = SCATTER ? (((uint64_t)__float_as_uint(depth[i]) << 32) | (uint32_t)i) : 0ull;
            for (int y = miny; y < maxy; ++y)
                for (int x = minx; x < maxx; ++x) {
                    const int tile = y * gx + x;
                    if (SCATTER) {
                        const uint32_t pos = (uint32_t)table[(size_t)tile * n_bb + bb] + (uint32_t)atomicAdd(&bins[tile], 1);
                        if (pos < ranges[tile].y) keys[pos] = key; // ranges are clipped to capacity
                    } else {
                        atomicAdd(&bins[tile], 1);
                    }
                }
        }
    }
    }
    if (!SCATTER) {
        __syncthreads();
        // per-(tile, block) counts for the scatter bases, and the tile totals: one non-returning atomic per non-empty bin
        // (<= n_bb per address over the whole launch — not the per-instance same-address storm measured in profiles/r1_*)
        for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
            const int c = bins[t];
            table[(size_t)t * n_bb + bb] = c;
            if (c) atomicAdd(&tile_total[t], c);
        }
    }
}

// ---- scans: ONE launch, five roles (independent chains run side by side instead of one after the other in a single CTA) ----
//   CTA 0      slot bases of the preprocess blocks (exclusive scan of their instance sums), R, overflow flags
//   CTA 1      tile bases / clipped ranges (published through a flag for the row scans), then the blend work items
//   CTA 2      execution order of the blend work items (chunk index major, see raster_render.cu)
//   CTA 3      sort work lists (small units, long tiles)
//   CTA >= 4   one warp per tile: exclusive scan of the tile's row of the (tile, binning block) table + the tile base = the
//              scatter bases.  They wait for CTA 1's flag (lower CTA indices are dispatched first, so the producer is always
//              resident before any consumer spins).
// Roles 1-3 each rebuild the clipped per-tile counts in shared memory (one 1200-element scan) rather than wait for one another.
#define SCAN_THREADS 1024
// exclusive scan of `n` ints in place (global or shared); each thread owns a contiguous run. Returns the total (valid in all threads).
__device__ int cta_exclusive_scan(int *data, int n, int *s_warp /* [33] */) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const int beg = min(n, t * per), end = min(n, beg + per);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += data[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    int run = s_warp[wid] + incl - sum;
    for (int i = beg; i < end; ++i) {
        int v = data[i];
        data[i] = run;
        run += v;
    }
    const int total = s_warp[32];
    __syncthreads();
    return total;
}

#define GSD_SORT_SEG 1024     // keys per small sort unit (one 256-thread CTA, shared memory)
#define GSD_SORT_LONG 4096    // longest list one 1024-thread CTA sorts in shared memory (80 KB: two CTAs per SM); longer lists: 1024-key segments + merge rounds
#define EXEC_LEVELS 1024      // chunk-index levels of the execution-order counting sort (deeper chunks share the last level)

// per-tile instance counts that fit the buffers, in shared memory: s_n[i] = clipped count, s_x[i] = unclipped exclusive base
__device__ void tile_counts(int n_tiles, int64_t capacity, const int32_t *__restrict__ tile_total, int *s_n, int *s_x, int *s_warp) {
    const int t = threadIdx.x;
    for (int i = t; i < n_tiles; i += SCAN_THREADS) { const int v = tile_total[i]; s_n[i] = v; s_x[i] = v; }
    __syncthreads();
    cta_exclusive_scan(s_x, n_tiles, s_warp);
    for (int i = t; i < n_tiles; i += SCAN_THREADS) {
        long long s = s_x[i], e = s + s_n[i];
        if (s > capacity) s = capacity;
        if (e > capacity) e = capacity;
        s_n[i] = (int)(e - s);
    }
    __syncthreads();
}
__device__ __forceinline__ int sort_units_of(int n) {   // small units: lists of 2..1024 keys (longer lists get a CTA of their own)
    return (n >= 2 && n <= GSD_SORT_SEG) ? 1 : 0;
}

__global__ void __launch_bounds__(SCAN_THREADS)
gsd_bin_scan_kernel(int n_pre_blocks, int n_tiles, int n_bb, int64_t capacity, int max_items, int max_units,
                    const uint32_t *__restrict__ block_sum, uint32_t *__restrict__ block_base, const int32_t *__restrict__ tile_total,
                    int32_t *__restrict__ tile_base, uint2 *__restrict__ ranges, int32_t *__restrict__ chunk_ptr,
                    int4 *__restrict__ item_tile, int32_t *__restrict__ counters, int32_t *__restrict__ unit_tile,
                    int32_t *__restrict__ unit_seg, int32_t *__restrict__ long_tile, int32_t *__restrict__ exec_item,
                    int32_t *__restrict__ table, int32_t *__restrict__ status, int32_t *__restrict__ sticky) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    extern __shared__ int sm[];       // roles 1-3: [2][n_tiles]
    __shared__ int s_warp[33];
    __shared__ int s_cnt;
    __shared__ int s_lvl[2][EXEC_LEVELS];
    const int t = threadIdx.x;
    int *s_n = sm, *s_x = sm + n_tiles;
    if (blockIdx.x == 0) {
        for (int i = t; i < n_pre_blocks; i += SCAN_THREADS) block_base[i] = block_sum[i];
        __syncthreads();
        const int R = cta_exclusive_scan((int *)block_base, n_pre_blocks, s_warp);
        if (t == 0) {
            status[0] = R;
            status[1] = ((long long)R > capacity) ? 1 : 0;
            if (sticky) {   // single writer (this thread; forward calls on one stream are ordered)
                sticky[0] = max(sticky[0], R);
                if ((long long)R > capacity) sticky[1] += 1;
            }
        }
        return;
    }
    if (blockIdx.x == 1) {
        // tile bases and clipped ranges, published for the row scans; then the blend work items (128-record chunks)
        tile_counts(n_tiles, capacity, tile_total, s_n, s_x, s_warp);
        for (int i = t; i < n_tiles; i += SCAN_THREADS) {
            long long s = s_x[i];
            tile_base[i] = (int)s;
            if (s > capacity) s = capacity;
            ranges[i] = make_uint2((uint32_t)s, (uint32_t)(s + s_n[i]));
        }
        __threadfence();
        __syncthreads();
        if (t == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(counters + 3), "r"(1) : "memory");
        for (int i = t; i < n_tiles; i += SCAN_THREADS) s_x[i] = (s_n[i] + GSD_CHUNK - 1) / GSD_CHUNK;
        __syncthreads();
        const int n_items = cta_exclusive_scan(s_x, n_tiles, s_warp);
        if (t == 0) {
            chunk_ptr[n_tiles] = n_items;
            counters[0] = n_items < max_items ? n_items : max_items;
        }
        for (int i = t; i < n_tiles; i += SCAN_THREADS) {
            const int c0 = s_x[i], c1 = c0 + (s_n[i] + GSD_CHUNK - 1) / GSD_CHUNK;
            chunk_ptr[i] = c0;
            // everything a blend CTA needs to know about its item in one 16-byte record (was: item -> tile -> chunk_ptr / ranges,
            // three dependent loads in front of the first TMA of every CTA)
            long long s = tile_base[i];
            if (s > capacity) s = capacity;
            const int rx = (int)s, ry = rx + s_n[i];
            for (int c = c0; c < c1 && c < max_items; ++c) {
                const int start = rx + (c - c0) * GSD_CHUNK;
                item_tile[c] = make_int4(i, c - c0, start, min(GSD_CHUNK, ry - start));
            }
        }
        return;
    }
    if (blockIdx.x == 2) {
        // execution order of the forward chunk kernel: work items sorted by CHUNK INDEX (all first chunks, then all second
        // chunks, ...), so that by the time a deep chunk of a tile is dispatched its predecessors have usually finished and it
        // can see that its pixels are already opaque (raster_render.cu).  Counting sort over the chunk index in shared memory;
        // the order inside a level only affects scheduling, never results.
        tile_counts(n_tiles, capacity, tile_total, s_n, s_x, s_warp);
        for (int l = t; l < EXEC_LEVELS; l += SCAN_THREADS) { s_lvl[0][l] = 0; s_lvl[1][l] = 0; }
        for (int i = t; i < n_tiles; i += SCAN_THREADS) s_x[i] = (s_n[i] + GSD_CHUNK - 1) / GSD_CHUNK;
        __syncthreads();
        for (int i = t; i < n_tiles; i += SCAN_THREADS) {
            const int nc = s_x[i];
            for (int c = 0; c < nc; ++c) atomicAdd(&s_lvl[0][min(c, EXEC_LEVELS - 1)], 1);
        }
        __syncthreads();
        cta_exclusive_scan(s_lvl[0], EXEC_LEVELS, s_warp);
        cta_exclusive_scan(s_x, n_tiles, s_warp);     // first work item of every tile (same scan as role 1)
        for (int i = t; i < n_tiles; i += SCAN_THREADS) {
            const int nc = (s_n[i] + GSD_CHUNK - 1) / GSD_CHUNK, c0 = s_x[i];
            for (int c = 0; c < nc; ++c) {
                const int l = min(c, EXEC_LEVELS - 1);
                const int pos = s_lvl[0][l] + atomicAdd(&s_lvl[1][l], 1);
                if (pos < max_items && c0 + c < max_items) exec_item[pos] = c0 + c;
            }
        }
        return;
    }
    if (blockIdx.x == 3) {
        // sort work lists: small units (lists of 2..1024 keys, and the 1024-key segments of lists > 8192) and long tiles (> 1024)
        tile_counts(n_tiles, capacity, tile_total, s_n, s_x, s_warp);
        if (t == 0) s_cnt = 0;
        for (int i = t; i < n_tiles; i += SCAN_THREADS) s_x[i] = sort_units_of(s_n[i]);
        __syncthreads();
        const int n_units = cta_exclusive_scan(s_x, n_tiles, s_warp);
        for (int i = t; i < n_tiles; i += SCAN_THREADS) {
            const int nu = sort_units_of(s_n[i]);
            const int u0 = s_x[i];
            for (int u = 0; u < nu && u0 + u < max_units; ++u) { unit_tile[u0 + u] = i; unit_seg[u0 + u] = u; }
            if (s_n[i] > GSD_SORT_SEG) long_tile[atomicAdd(&s_cnt, 1)] = i;   // order only affects scheduling, never results
        }
        __syncthreads();
        if (t == 0) {
            counters[1] = n_units < max_units ? n_units : max_units;
            counters[2] = s_cnt;
        }
        return;
    }
    // row scans: one warp per tile; wait for role 1's tile bases
    if (t == 0) {
        int ready = 0;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(ready) : "l"(counters + 3) : "memory");
            if (!ready) __nanosleep(64);
        } while (!ready);
    }
    __syncthreads();
    const int lane = t & 31;
    const int tile = (blockIdx.x - 4) * (SCAN_THREADS / 32) + (t >> 5);
    if (tile >= n_tiles) return;
    int run = __ldcg(tile_base + tile);
    for (int b0 = 0; b0 < n_bb; b0 += 32) {
        const int b = b0 + lane;
        const int v = b < n_bb ? table[(size_t)tile * n_bb + b] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (b < n_bb) table[(size_t)tile * n_bb + b] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ---- per-tile sort -----------------------------------------------------------------------------------------------------
// Keys are (depth bits << 32 | Gaussian id): unique inside a tile, ties in depth ordered by id — exactly the order of the
// reference's stable (tile | depth) radix sort.  Lists are sorted by a MONOTONE BUCKET SORT in shared memory:
//   1. min / max depth of the list;  2. bucket = floor((d - dmin) * nb / (dmax - dmin)) — monotone in d, so every key of a
//   bucket precedes every key of the next one and equal depths share a bucket;  3. counting sort into the buckets (shared-memory
//   atomics: ~2 keys per bucket, few conflicts);  4. every key counts the smaller keys of its own bucket (full 64-bit compare)
//   and is written to its final position.
// ~10 shared-memory operations per key where the rank-merge sort it replaces needed ~160 bank-conflicting binary-search reads
// (measured: 55 us for the 100k benchmark scene with one CTA per tile, 24 + 20 us as segment sort + multiway merge — shared-
// memory wavefront bound).  A degenerate list (all depths equal) costs n^2 / threads compares per CTA: bounded, never wrong.
// (gsd_tile_sort_kernel below: one launch, long lists by whole CTAs, four short lists per CTA.)
// NT threads (a whole CTA, or one 256-thread group of it: `bar` is the named barrier of the group, tid the index inside it)
template <int NT>
__device__ __forceinline__ void group_sync(int bar) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(NT) : "memory"); }

template <int NT>
__device__ __forceinline__ void bucket_sort_smem(const unsigned long long *__restrict__ src, unsigned long long *__restrict__ dst,
                                                 unsigned long long *buf, int *s_start, int *s_cur, float *s_red, int n, int nb,
                                                 int tid, int bar) {
    const int lane = tid & 31, wid = tid >> 5;
    float lo = 3.4e38f, hi = 0.f;                 // depths are positive (> 0.2): float order == bit order
    for (int i = tid; i < n; i += NT) {
        const float d = __uint_as_float((unsigned)(src[i] >> 32));
        lo = fminf(lo, d); hi = fmaxf(hi, d);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if (lane == 0) { s_red[wid] = lo; s_red[32 + wid] = hi; }
    for (int b = tid; b < nb; b += NT) s_cur[b] = 0;
    group_sync<NT>(bar);
    lo = s_red[0]; hi = s_red[32];
    for (int w = 1; w < NT / 32; ++w) { lo = fminf(lo, s_red[w]); hi = fmaxf(hi, s_red[32 + w]); }
    const float scale = hi > lo ? (float)nb / (hi - lo) : 0.f;
    auto bucket_of = [&](unsigned long long k) { return min(nb - 1, (int)((__uint_as_float((unsigned)(k >> 32)) - lo) * scale)); };
    for (int i = tid; i < n; i += NT) atomicAdd(&s_cur[bucket_of(src[i])], 1);
    group_sync<NT>(bar);
    // exclusive scan of the bucket counts (nb <= 4 * NT): each thread owns a contiguous run of buckets
    {
        const int per = (nb + NT - 1) / NT;
        const int beg = min(nb, tid * per), end = min(nb, beg + per);
        int sum = 0;
        for (int b = beg; b < end; ++b) sum += s_cur[b];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int *s_w = reinterpret_cast<int *>(s_red);   // min / max already consumed by every thread? no: guard with a barrier
        group_sync<NT>(bar);
        if (lane == 31) s_w[wid] = incl;
        group_sync<NT>(bar);
        int wbase = 0;
        for (int w = 0; w < wid; ++w) wbase += s_w[w];
        int run = wbase + incl - sum;
        for (int b = beg; b < end; ++b) { const int c = s_cur[b]; s_start[b] = run; s_cur[b] = run; run += c; }
    }
    group_sync<NT>(bar);
    for (int i = tid; i < n; i += NT) {
        const unsigned long long k = src[i];
        buf[atomicAdd(&s_cur[bucket_of(k)], 1)] = k;     // s_cur[b] ends as the bucket's end
    }
    group_sync<NT>(bar);
    for (int i = tid; i < n; i += NT) {
        const unsigned long long k = buf[i];
        const int b = bucket_of(k);
        const int s0 = s_start[b], s1 = s_cur[b];
        int rank = 0;
        for (int j = s0; j < s1; ++j) rank += buf[j] < k;
        dst[s0 + rank] = k;
    }
}

#define MERGE_THREADS 1024
#define MERGE_FAN 16
// number of elements of the sorted run a[0, len) that are < key (fixed-trip branchless lower bound; len <= cap, cap a power of two)
template <typename Ptr>
__device__ __forceinline__ int lower_bound_pow2(Ptr a, int len, int cap, unsigned long long key) {
    int lo = 0;
    for (int step = cap; step >= 1; step >>= 1) {
        const int idx = lo + step - 1;
        if (idx < len && a[idx] < key) lo += step;
    }
    return lo;
}

// one round: runs of `run` sorted keys of src[0, n) are merged in groups of up to MERGE_FAN into dst
__device__ __forceinline__ void merge_round(const unsigned long long *src, unsigned long long *dst, int n, long long run, int cap, int tid) {
    const long long group = run * MERGE_FAN;
    for (int i = tid; i < n; i += MERGE_THREADS) {
        const unsigned long long key = src[i];
        const long long base = (i / group) * group;
        const long long glen = min(group, (long long)n - base);
        const int self = (int)((i - base) / run);
        int rank = (int)((i - base) - self * run);
        const int nruns = (int)((glen + run - 1) / run);
        for (int q = 0; q < nruns; ++q) {
            if (q == self) continue;
            rank += lower_bound_pow2(src + base + q * run, (int)min(run, glen - q * run), cap, key);
        }
        dst[base + rank] = key;
    }
}

// ONE launch sorts every tile list.  Work index w of a 1024-thread CTA:
//   w < n_long    a list of 1025..4096 keys: bucket sort by the whole CTA.  A longer list: the CTA bucket-sorts its 4096-key
//                 segments one after the other, then runs 16-way rank-merge rounds over them (ping-pong through keys_tmp in L2)
//   w >= n_long   four lists of 2..1024 keys at once, one per 256-thread group (named barriers 1..4)
__global__ void __launch_bounds__(MERGE_THREADS)
gsd_tile_sort_kernel(const uint2 *__restrict__ ranges, uint64_t *__restrict__ keys, uint64_t *__restrict__ keys_tmp,
                     const int32_t *__restrict__ unit_tile, const int32_t *__restrict__ unit_seg, const int32_t *__restrict__ long_tile,
                     const int32_t *__restrict__ counters) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    extern __shared__ unsigned long long smk[];   // long role: [2][GSD_SORT_LONG] keys + [2][GSD_SORT_LONG / 2] ints;  small role: 4 x ([2][1024] keys + [2][512] ints)
    __shared__ float s_red[4][64];
    const int tid = threadIdx.x;
    const int n_long = counters[2];
    const int n_small = (counters[1] + 3) / 4;
    for (int w = blockIdx.x; w < n_long + n_small; w += gridDim.x) {   // static grid (CUDA-graph friendly), device-side counts
        if (w < n_long) {
            const uint2 r = ranges[long_tile[w]];
            const int n = (int)(r.y - r.x);
            unsigned long long *gk = reinterpret_cast<unsigned long long *>(keys) + r.x;
            int *s_start = reinterpret_cast<int *>(smk + 2 * GSD_SORT_LONG), *s_cur = s_start + GSD_SORT_LONG / 2;
            for (int off = 0; off < n; off += GSD_SORT_LONG) {     // one pass for lists up to 4096 keys
                const int m = min(GSD_SORT_LONG, n - off);
                for (int i = tid; i < m; i += MERGE_THREADS) smk[i] = gk[off + i];
                __syncthreads();
                int nb = 512;
                while (nb * 2 < m && nb < GSD_SORT_LONG / 2) nb <<= 1;
                bucket_sort_smem<MERGE_THREADS>(smk, gk + off, smk + GSD_SORT_LONG, s_start, s_cur, s_red[0], m, nb, tid, 0);
                __syncthreads();
            }
            if (n > GSD_SORT_LONG) {
                // very long list: 16-way rank-merge rounds over its sorted 4096-key segments, in L2
                unsigned long long *src = gk, *dst = reinterpret_cast<unsigned long long *>(keys_tmp) + r.x;
                long long run = GSD_SORT_LONG;
                int cap = GSD_SORT_LONG;
                while (run < n) {
                    merge_round(src, dst, n, run, cap, tid);
                    __syncthreads();
                    unsigned long long *tmp = src; src = dst; dst = tmp;
                    run *= MERGE_FAN;
                    cap = (int)min((long long)cap * MERGE_FAN, (long long)1 << 30);
                }
                if (src != gk)
                    for (int i = tid; i < n; i += MERGE_THREADS) gk[i] = src[i];
            }
        } else {
            const int grp = tid >> 8, gtid = tid & 255;
            const int unit = (w - n_long) * 4 + grp;
            if (unit < counters[1]) {     // uniform inside the 256-thread group
                unsigned long long *sk = smk + (size_t)grp * (2 * GSD_SORT_SEG + GSD_SORT_SEG / 2);   // 2 x 1024 keys + 2 x 512 ints = 20 KB per group
                int *s_start = reinterpret_cast<int *>(sk + 2 * GSD_SORT_SEG), *s_cur = s_start + GSD_SORT_SEG / 2;
                const uint2 r = ranges[unit_tile[unit]];
                const int off = unit_seg[unit] * GSD_SORT_SEG;
                const int n = min(GSD_SORT_SEG, (int)(r.y - r.x) - off);
                unsigned long long *gk = reinterpret_cast<unsigned long long *>(keys) + r.x + off;
                for (int i = gtid; i < n; i += 256) sk[i] = gk[i];
                group_sync<256>(1 + grp);
                int nb = 32;
                while (nb * 2 < n && nb < GSD_SORT_SEG / 2) nb <<= 1;     // ~2 keys per bucket
                bucket_sort_smem<256>(sk, gk, sk + GSD_SORT_SEG, s_start, s_cur, s_red[grp], n, nb, gtid, 1 + grp);
            }
        }
        __syncthreads();
    }
}

#define LONG_SORT_SMEM (2 * GSD_SORT_LONG * 8 + 2 * (GSD_SORT_LONG / 2) * 4)
// ---- host launchers -------------------------------------------------------------------------------------
static int set_bin_attrs() {
    static bool done = false;
    if (!done) {
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_TILES_SMEM * 4));
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_TILES_SMEM * 4));
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_bin_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_TILES_SMEM * 8));
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LONG_SORT_SMEM));
        done = true;
    }
    return GSD_OK;
}

// count only: scan of the block sums -> status[0] = R  (what upstream copies to the host before binning)
__global__ void __launch_bounds__(SCAN_THREADS)
gsd_count_kernel(int n_pre_blocks, const uint32_t *__restrict__ block_sum, int32_t *__restrict__ status) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    __shared__ long long red[SCAN_THREADS / 32];
    long long s = 0;
    for (int i = threadIdx.x; i < n_pre_blocks; i += SCAN_THREADS) s += block_sum[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) tot += red[k];
        status[0] = (int32_t)tot;
        status[1] = 0;
    }
}

int gsd_launch_count(int G, const GsdGeomWs &g, int32_t *status, cudaStream_t st) {
    gsd_launch(gsd_count_kernel, dim3(1), dim3(SCAN_THREADS), 0, st, G > 0 ? (G + 255) / 256 : 0, g.block_sum, status);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

int gsd_launch_binning(int G, const GsdCam &cam, const GsdRasterFwd *a, const GsdGeomWs &g, const GsdBinWs &b,
                       int32_t *zero_flags, int n_flags, cudaStream_t st) {
    const int64_t cap = a->capacity;
    const int tiles = cam.gx * cam.gy;
    if (tiles > MAX_TILES_SMEM) {
        gsd_set_error("image has %d tiles; the binning histogram supports up to %d (e.g. 2048x1536)", tiles, MAX_TILES_SMEM);
        return GSD_ERR_UNSUPPORTED;
    }
    int rc;
    if ((rc = set_bin_attrs())) return rc;
    const int n_pre = G > 0 ? (G + 255) / 256 : 0;
    const size_t smem = (size_t)tiles * 4;
    // tile totals are accumulated with atomics by the histogram pass: the preprocess kernel zero-fills them (G > 0)
    if (G > 0) {
        gsd_launch((gsd_bin_kernel<false>), dim3(b.n_bb), dim3(BIN_THREADS), smem, st, G, cam.gx, tiles, b.n_bb, g.tiles, g.rect, g.depth, b.table,
                                                                    b.ranges, b.keys, g.slot_base, g.block_base, b.tile_total, b.counters, zero_flags, n_flags);
        GSD_LAUNCH_CHECK();
    } else {
        GSD_CUDA_CHECK(cudaMemsetAsync(b.table, 0, (size_t)tiles * b.n_bb * 4, st));
        GSD_CUDA_CHECK(cudaMemsetAsync(b.tile_total, 0, (size_t)tiles * 4, st));
        GSD_CUDA_CHECK(cudaMemsetAsync(b.counters, 0, 8 * 4, st));
    }
    const int row_ctas = (tiles + SCAN_THREADS / 32 - 1) / (SCAN_THREADS / 32);
    gsd_launch(gsd_bin_scan_kernel, dim3(4 + row_ctas), dim3(SCAN_THREADS), (size_t)tiles * 8, st, n_pre, tiles, b.n_bb, cap, b.max_items, b.max_units, g.block_sum, g.block_base, b.tile_total, b.tile_base, b.ranges, b.chunk_ptr,
        b.item_tile, b.counters, b.unit_tile, b.unit_seg, b.long_tile, b.exec_item, b.table, a->status, a->sticky);
    GSD_LAUNCH_CHECK();
    if (G == 0 || cap == 0) return GSD_OK;
    gsd_launch((gsd_bin_kernel<true>), dim3(b.n_bb), dim3(BIN_THREADS), smem, st, G, cam.gx, tiles, b.n_bb, g.tiles, g.rect, g.depth, b.table, b.ranges,
                                                               b.keys, g.slot_base, g.block_base, b.tile_total, b.counters, nullptr, 0);
    GSD_LAUNCH_CHECK();
    gsd_launch(gsd_tile_sort_kernel, dim3(148 * 2), dim3(MERGE_THREADS), LONG_SORT_SMEM, st, b.ranges, b.keys, b.keys_tmp, b.unit_tile, b.unit_seg, b.long_tile,
                                                                         b.counters);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
