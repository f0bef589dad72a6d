// raster_bin.cu — tile binning without a global sort and without global atomics:
//   (preprocess)  per-Gaussian tile rectangles + per-256-block instance sums
//   hist          one CTA per 1024 Gaussians: shared-memory histogram of touched tiles -> table[tile][block]
//   scan          one CTA: exclusive scans of the block sums (slot bases, R), of the tile-major table (scatter bases, per-tile
//                 ranges) and of the chunks per tile (blend work items)
//   scatter       one CTA per 1024 Gaussians: append (depth bits << 32 | id) to each touched tile's segment at
//                 base[tile][block] + shared-memory cursor
//   sort_pack     one CTA per tile: bitonic sort of the segment (ties by Gaussian id, i.e. exactly the order of the reference's
//                 stable (tile | depth) radix sort), then gather the Gaussian data into the four SoA record planes that the blend
//                 kernels stream with 1-D bulk (TMA) copies
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
    ws->table = (int32_t *)take((size_t)tiles * ws->n_bb * 4);
    ws->tile_base = (int32_t *)take((size_t)tiles * 4);
    ws->ranges = (uint2 *)take((size_t)tiles * sizeof(uint2));
    ws->chunk_ptr = (int32_t *)take((size_t)(tiles + 1) * 4);
    ws->item_tile = (int32_t *)take((size_t)ws->max_items * 4);
    ws->counters = (int32_t *)take(8 * 4);
    ws->sort_order = (int32_t *)take((size_t)tiles * 4);
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
    ws->total = off;
    return GSD_OK;
}

// ---- hist / scatter: one CTA per GSD_BIN_BLOCK Gaussians ------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(GSD_BIN_BLOCK)
gsd_bin_kernel(int G, int gx, int n_tiles, int n_bb, const uint32_t *__restrict__ tiles, const uint2 *__restrict__ rect,
               const float *__restrict__ depth, int32_t *__restrict__ table, const uint2 *__restrict__ ranges,
               uint64_t *__restrict__ keys, uint32_t *__restrict__ slot_base, const uint32_t *__restrict__ block_base) {
    extern __shared__ int bins[]; // [n_tiles] counts (hist) or cursors (scatter)
    const int bb = blockIdx.x;
    for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    const int i = bb * GSD_BIN_BLOCK + threadIdx.x;
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
    if (!SCATTER) {
        __syncthreads();
        for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) table[(size_t)t * n_bb + bb] = bins[t];
    }
}

// ---- scans (one CTA) -------------------------------------------------------------------------------------------
#define SCAN_THREADS 1024
// exclusive scan of `n` ints in place; each thread owns a contiguous run. Returns the total (valid in all threads).
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

// S1: one warp per tile: total of the tile's row of the table
__global__ void __launch_bounds__(256)
gsd_bin_tile_sum_kernel(int n_tiles, int n_bb, const int32_t *__restrict__ table, int32_t *__restrict__ tile_total) {
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    int s = 0;
    for (int b = lane; b < n_bb; b += 32) s += table[(size_t)tile * n_bb + b];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) tile_total[tile] = s;
}

// S2: one CTA: slot bases of the preprocess blocks (R), tile bases / ranges, chunk work items
__global__ void __launch_bounds__(SCAN_THREADS)
gsd_bin_scan_kernel(int n_pre_blocks, int n_tiles, int64_t capacity, int max_items, uint32_t *__restrict__ block_sum,
                    uint32_t *__restrict__ block_base, int32_t *__restrict__ tile_base, uint2 *__restrict__ ranges,
                    int32_t *__restrict__ chunk_ptr, int32_t *__restrict__ item_tile, int32_t *__restrict__ counters,
                    int32_t *__restrict__ sort_order, int32_t *__restrict__ status, int32_t *__restrict__ sticky) {
    __shared__ int s_warp[33];
    __shared__ int s_cls[3], s_fill[3];
    const int t = threadIdx.x;
    for (int i = t; i < n_pre_blocks; i += SCAN_THREADS) block_base[i] = block_sum[i];
    __syncthreads();
    const int R = cta_exclusive_scan((int *)block_base, n_pre_blocks, s_warp);
    if (t == 0) {
        status[0] = R;
        status[1] = ((long long)R > capacity) ? 1 : 0;
        if (sticky) {   // single writer (this thread of this one-CTA kernel; forward calls on one stream are ordered)
            sticky[0] = max(sticky[0], R);
            if ((long long)R > capacity) sticky[1] += 1;
        }
    }
    for (int i = t; i < n_tiles; i += SCAN_THREADS) chunk_ptr[i] = tile_base[i]; // keep the totals: chunk_ptr is scratch here
    __syncthreads();
    const int total = cta_exclusive_scan(tile_base, n_tiles, s_warp);
    for (int i = t; i < n_tiles; i += SCAN_THREADS) {
        long long s = tile_base[i], e = s + chunk_ptr[i];
        if (s > capacity) s = capacity;
        if (e > capacity) e = capacity;
        ranges[i] = make_uint2((uint32_t)s, (uint32_t)e);
    }
    __syncthreads();
    if (t < 3) { s_cls[t] = 0; s_fill[t] = 0; }
    __syncthreads();
    // work list of the per-tile sort: only tiles with >= 2 instances, the longest lists first (the sort of a 2000-key tile
    // is the critical path of that kernel; the order inside a class only affects scheduling, never results)
    auto cls_of = [](int n) { return n > 1024 ? 0 : (n > 256 ? 1 : 2); };
    for (int i = t; i < n_tiles; i += SCAN_THREADS) {
        const uint2 r = ranges[i];
        const int n = (int)(r.y - r.x);
        chunk_ptr[i] = (n + GSD_CHUNK - 1) / GSD_CHUNK;
        if (n >= 2) atomicAdd(&s_cls[cls_of(n)], 1);
    }
    __syncthreads();
    for (int i = t; i < n_tiles; i += SCAN_THREADS) {
        const uint2 r = ranges[i];
        const int n = (int)(r.y - r.x);
        if (n >= 2) {
            const int c = cls_of(n);
            const int base = (c > 0 ? s_cls[0] : 0) + (c > 1 ? s_cls[1] : 0);
            sort_order[base + atomicAdd(&s_fill[c], 1)] = i;
        }
    }
    if (t == 0) counters[1] = s_cls[0] + s_cls[1] + s_cls[2];
    __syncthreads();
    const int n_items = cta_exclusive_scan(chunk_ptr, n_tiles, s_warp);
    if (t == 0) {
        chunk_ptr[n_tiles] = n_items;
        counters[0] = n_items < max_items ? n_items : max_items;
    }
    __syncthreads();
    for (int i = t; i < n_tiles; i += SCAN_THREADS) {
        const int c0 = chunk_ptr[i], c1 = (i + 1 < n_tiles) ? chunk_ptr[i + 1] : n_items;
        for (int c = c0; c < c1 && c < max_items; ++c) item_tile[c] = i;
    }
    (void)total;
}

// S3: one warp per tile: exclusive scan of the tile's row + tile base -> scatter bases
__global__ void __launch_bounds__(256)
gsd_bin_tile_scan_kernel(int n_tiles, int n_bb, int32_t *__restrict__ table, const int32_t *__restrict__ tile_base) {
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    int run = tile_base[tile];
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

// ---- per-tile sort + pack ----------------------------------------------------------------------------------
// Keys are unique inside a tile (the Gaussian id is part of the key), so a rank-based merge sort needs no tie handling:
//   1. every warp sorts groups of 32 keys in registers (bitonic network over shuffles, no barriers);
//   2. log2(n/32) merge rounds: every key binary-searches the sibling run and is written to its final position of the
//      merged run in the other buffer (one barrier per round).
// ~6 rounds of ~11 dependent shared-memory reads for the heaviest benchmark tile (1900 keys) instead of the 66 barrier-
// separated steps of a bitonic network; lists longer than SORT_SMEM_KEYS run the same rounds in global memory (L2).
#define SORT_THREADS 512
#define SORT_SMEM_KEYS 2048

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ unsigned long long warp_sort32(unsigned long long key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 1; j >>= 1) {
            const unsigned long long other = shfl_xor_u64(key, j);
            const bool up = (lane & k) == 0;       // ascending block (k = 32: always ascending)
            const bool lower = (lane & j) == 0;    // this lane keeps the smaller key of the pair when ascending
            const bool take_min = (lower == up);
            key = take_min ? (key < other ? key : other) : (key > other ? key : other);
        }
    }
    return key;
}

template <typename Ptr>
__device__ __forceinline__ int lower_bound_u64(Ptr a, int len, unsigned long long key) {
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// sorts n keys; src/dst are two buffers of n keys; returns the buffer holding the result
template <typename Ptr>
__device__ __forceinline__ Ptr cta_merge_sort(Ptr src, Ptr dst, int n, int tid, int nthreads) {
    const int lane = tid & 31, wid = tid >> 5, nwarps = nthreads >> 5;
    for (int g = wid; g * 32 < n; g += nwarps) {
        const int i = g * 32 + lane;
        unsigned long long k = i < n ? src[i] : ~0ull;
        k = warp_sort32(k, lane);
        if (i < n) src[i] = k;
    }
    __syncthreads();
    for (int run = 32; run < n; run <<= 1) {
        for (int i = tid; i < n; i += nthreads) {
            const unsigned long long key = src[i];
            const int r = i / run, base_self = r * run, base_sib = (r ^ 1) * run;
            const int sib_len = max(0, min(run, n - base_sib));
            const int rank = sib_len > 0 ? lower_bound_u64(src + base_sib, sib_len, key) : 0;
            dst[min(base_self, base_sib) + (i - base_self) + rank] = key;
        }
        __syncthreads();
        Ptr tmp = src; src = dst; dst = tmp;
    }
    return src;
}

// shared-memory variant: buffers addressed as sk[cur] / sk[cur ^ 1] (the compiler keeps LDS/STS instead of generic accesses),
// four keys per thread searched in lockstep with a fixed-trip branchless lower bound (the searches of a round are independent,
// so their shared-memory latencies overlap instead of adding up).  Returns the index of the buffer holding the result.
__device__ __forceinline__ int cta_merge_sort_smem(unsigned long long (*sk)[SORT_SMEM_KEYS], int n, int tid) {
    constexpr int KPT = 4;
    const int lane = tid & 31, wid = tid >> 5;
    for (int g = wid; g * 32 < n; g += SORT_THREADS / 32) {
        const int i = g * 32 + lane;
        unsigned long long k = i < n ? sk[0][i] : ~0ull;
        k = warp_sort32(k, lane);
        if (i < n) sk[0][i] = k;
    }
    __syncthreads();
    int cur = 0;
    for (int run = 32; run < n; run <<= 1) {
        for (int i0 = tid; i0 < n; i0 += KPT * SORT_THREADS) {
            unsigned long long key[KPT];
            int sib[KPT], len[KPT], lo[KPT], out[KPT];
#pragma unroll
            for (int u = 0; u < KPT; ++u) {
                const int i = i0 + u * SORT_THREADS;
                const bool v = i < n;
                key[u] = v ? sk[cur][i] : 0ull;
                const int r = i / run, base_self = r * run;
                sib[u] = (r ^ 1) * run;
                len[u] = v ? max(0, min(run, n - sib[u])) : 0;
                out[u] = min(base_self, sib[u]) + (i - base_self);
                lo[u] = 0;
            }
            for (int step = run; step >= 1; step >>= 1) {
#pragma unroll
                for (int u = 0; u < KPT; ++u) {
                    const int idx = lo[u] + step - 1;
                    if (idx < len[u] && sk[cur][sib[u] + idx] < key[u]) lo[u] += step;
                }
            }
#pragma unroll
            for (int u = 0; u < KPT; ++u)
                if (i0 + u * SORT_THREADS < n) sk[cur ^ 1][out[u] + lo[u]] = key[u];
        }
        __syncthreads();
        cur ^= 1;
    }
    return cur;
}

// one CTA per tile: sort the tile's keys in place (the blend forward gathers the Gaussian data by sorted key)
__global__ void __launch_bounds__(SORT_THREADS)
gsd_tile_sort_kernel(const uint2 *__restrict__ ranges, uint64_t *__restrict__ keys, uint64_t *__restrict__ keys_tmp,
                     const int32_t *__restrict__ sort_order, const int32_t *__restrict__ counters) {
    __shared__ unsigned long long sk[2][SORT_SMEM_KEYS];
    const int t = threadIdx.x;
    // static grid (CUDA-graph friendly); only the first counters[1] CTAs have work: the tiles with >= 2 keys, longest first
    if ((int)blockIdx.x >= counters[1]) return;
    const uint2 r = ranges[sort_order[blockIdx.x]];
    const int n = (int)(r.y - r.x);
    if (n <= 1) return;
    unsigned long long *gk = reinterpret_cast<unsigned long long *>(keys) + r.x;
    if (n <= SORT_SMEM_KEYS) {
        for (int i = t; i < n; i += SORT_THREADS) sk[0][i] = gk[i];
        __syncthreads();
        const int res = cta_merge_sort_smem(sk, n, t);
        for (int i = t; i < n; i += SORT_THREADS) gk[i] = sk[res][i];
    } else {
        unsigned long long *tmp = reinterpret_cast<unsigned long long *>(keys_tmp) + r.x;
        const unsigned long long *sorted = cta_merge_sort(gk, tmp, n, t, SORT_THREADS);
        if (sorted != gk)
            for (int i = t; i < n; i += SORT_THREADS) gk[i] = sorted[i];
    }
}

// ---- host launchers -------------------------------------------------------------------------------------
static int set_bin_attrs() {
    static bool done = false;
    if (!done) {
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_TILES_SMEM * 4));
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_TILES_SMEM * 4));
        done = true;
    }
    return GSD_OK;
}

// count only: scan of the block sums -> status[0] = R  (what upstream copies to the host before binning)
__global__ void __launch_bounds__(SCAN_THREADS)
gsd_count_kernel(int n_pre_blocks, const uint32_t *__restrict__ block_sum, int32_t *__restrict__ status) {
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
    gsd_count_kernel<<<1, SCAN_THREADS, 0, st>>>(G > 0 ? (G + 255) / 256 : 0, g.block_sum, status);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

int gsd_launch_binning(int G, const GsdCam &cam, const GsdRasterFwd *a, const GsdGeomWs &g, const GsdBinWs &b,
                       cudaStream_t st) {
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
    if (G > 0) {
        gsd_bin_kernel<false><<<b.n_bb, GSD_BIN_BLOCK, smem, st>>>(G, cam.gx, tiles, b.n_bb, g.tiles, g.rect, g.depth, b.table,
                                                                    b.ranges, b.keys, g.slot_base, g.block_base);
        GSD_LAUNCH_CHECK();
    } else {
        GSD_CUDA_CHECK(cudaMemsetAsync(b.table, 0, (size_t)tiles * b.n_bb * 4, st));
    }
    gsd_bin_tile_sum_kernel<<<(tiles + 7) / 8, 256, 0, st>>>(tiles, b.n_bb, b.table, b.tile_base);
    GSD_LAUNCH_CHECK();
    gsd_bin_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(n_pre, tiles, cap, b.max_items, g.block_sum, g.block_base, b.tile_base,
                                                     b.ranges, b.chunk_ptr, b.item_tile, b.counters, b.sort_order, a->status, a->sticky);
    GSD_LAUNCH_CHECK();
    gsd_bin_tile_scan_kernel<<<(tiles + 7) / 8, 256, 0, st>>>(tiles, b.n_bb, b.table, b.tile_base);
    GSD_LAUNCH_CHECK();
    if (G == 0 || cap == 0) return GSD_OK;
    gsd_bin_kernel<true><<<b.n_bb, GSD_BIN_BLOCK, smem, st>>>(G, cam.gx, tiles, b.n_bb, g.tiles, g.rect, g.depth, b.table, b.ranges,
                                                               b.keys, g.slot_base, g.block_base);
    GSD_LAUNCH_CHECK();
    gsd_tile_sort_kernel<<<tiles, SORT_THREADS, 0, st>>>(b.ranges, b.keys, b.keys_tmp, b.sort_order, b.counters);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
