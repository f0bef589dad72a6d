// raster_render.cu — the blend kernels (forward and backward) of the tile rasterizer, sm_100a.
//
// Replaces renderCUDA<3> forward/backward of the upstream rasterizer, reached from
// /root/reference/src/tracking/train_utils.py:178,192 (forward) and train_gs.py:31 (backward).
//
// Design (B200-first, not a translation).  A pixel's front-to-back recurrence is serial in the length of its tile list, and
// real scenes concentrate the instances on few tiles (benchmark scene: 317 of 1200 tiles, up to 1900 instances; a tracked
// object: tens of tiles with thousands).  One CTA per tile therefore leaves most SMs idle behind a few long serial chains
// (profiles/r1_*).  Here every tile list is cut into CHUNKS of 256 records and each (tile, chunk) is an independent work item:
//   fwd A1  per item   local composite of the chunk (transmittance product P, colour/depth sums with T starting at 1)
//   fwd B   per tile   combine the chunk composites in order; the chunk in which T_in * P crosses 1e-4 holds the reference's
//                      stopping point — that pixel's chunk is replayed sequentially with the true T_in; -> colour, depth,
//                      final_T, n_contrib
//   bwd B'  per tile   per chunk and pixel: T_in and C_pre (the colour composited in front of the chunk), from the stored chunk
//                      composites — independent of dL/dC, so it can run beside the loss kernels
//   bwd A'  per item   gradients of the chunk's Gaussians from T_in and Q_in = dL/dC . (C_final - T_final bg - C_pre), the colour
//                      still to come: front-to-back like the forward (T rebuilt by
//                      multiplication, not division); per-pixel partials summed across the warp with a transposed butterfly
//                      (13 shuffles for 12 values), across warps in fixed order by a flusher warp, one 64-byte record per
//                      instance: no atomics, bit-reproducible gradients
// fwd C   per pair   the crossing chunk replayed for one 8x4 rectangle by one warp (list of (item, rectangle) pairs from pass B)
// Inside an item the chunk's records are one TMA bulk copy per SoA plane (cp.async.bulk + mbarrier); 8 warps own 8x4 pixel
// rectangles; every 32 records the lanes test one record each against the rectangle (conservative extents of the
// alpha >= 1/255 ellipse) and only ballot survivors are blended, GSD_ILP at a time.
#include "common.cuh"
#include <type_traits>

#define GSD_SUB 32     // records per cull group / backward sub-batch
#define GSD_CWARPS 8   // consumer warps = 8x4 pixel rectangles of a 16x16 tile
#define GSD_ILP 2
#define T_EPS 0.0001f
#define TERMINAL_NONE (-1)

// per-item state (floats, SoA over the 256 pixels): P, D, C[CH], last(int) | bwd: T_in, Cpre[CH] (colour composited in front of the item)
template <int CH> struct ItemState {
    static constexpr int P = 0, D = 1, C = 2, LAST = 2 + CH, TIN = 3 + CH, CPRE = 4 + CH, NF = 4 + 2 * CH;
};
// per-tile terminal record (written by the one chunk that terminates a pixel): T_stop, D_abs, C_abs[CH], last(int), cstar(int)
template <int CH> struct TermState {
    static constexpr int T = 0, D = 1, C = 2, LAST = 2 + CH, CSTAR = 3 + CH, NF = 4 + CH;
};

// The accumulator ring of the blend backward hands a stage over with ONE mbarrier arrival per warp (lane 0, after __syncwarp()
// has ordered the other lanes' shared-memory accesses before it).  compute-sanitizer's racecheck attributes an arrival only to
// the arriving thread, so it reports the other 31 lanes' accesses as hazards; -DGSD_RACECHECK_ARRIVE_ALL makes every lane arrive
// (same protocol, 32x the arrival count) for the sanitizer run recorded in profiles/.
#ifdef GSD_RACECHECK_ARRIVE_ALL
#define GSD_RING_ARRIVALS 32
#else
#define GSD_RING_ARRIVALS 1
#endif
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Exact test "does any pixel of the warp's 8x4 rectangle receive alpha >= 1/255 from this Gaussian": minimum of the
// quadratic form q(d) = A dx^2 + 2 B dx dy + C dy^2 over the rectangle (in offset space) against the bound 2 ln(255 o).
// Evaluated by ONE lane per record (32 records per warp instruction), so even ~40 flops here are ~1 warp instruction per
// record, while every false survivor costs ~30. Anything non-finite or degenerate passes (conservative).
__device__ __forceinline__ bool cull_pass(const float4 g0, const float4 g1, float rx0, float rx1, float ry0, float ry1) {
    const float x0 = g0.x - rx1, x1 = g0.x - rx0, y0 = g0.y - ry1, y1 = g0.y - ry0; // offsets d = centre - pixel
    const float A = g1.x, B = g1.y, C = g1.z;
    if (!(A > 0.f) || !(C > 0.f)) return true;
    const bool in_x = x0 <= 0.f && x1 >= 0.f, in_y = y0 <= 0.f && y1 >= 0.f;
    float qmin = 0.f;
    if (!(in_x && in_y)) {
        const float nBA = -B * gsd_rcp_approx(A), nBC = -B * gsd_rcp_approx(C);   // A, C > 0; the 1e-4 margin below absorbs the ulp
        float ys = fminf(fmaxf(nBC * x0, y0), y1);
        float q1 = A * x0 * x0 + 2.f * B * x0 * ys + C * ys * ys;
        ys = fminf(fmaxf(nBC * x1, y0), y1);
        float q2 = A * x1 * x1 + 2.f * B * x1 * ys + C * ys * ys;
        float xs = fminf(fmaxf(nBA * y0, x0), x1);
        float q3 = A * xs * xs + 2.f * B * xs * y0 + C * y0 * y0;
        xs = fminf(fmaxf(nBA * y1, x0), x1);
        float q4 = A * xs * xs + 2.f * B * xs * y1 + C * y1 * y1;
        qmin = fminf(fminf(q1, q2), fminf(q3, q4));
    }
    return !(qmin * 0.9999f > g0.z);
}

struct ItemInfo { int tile, chunk, start, cnt, px, py, pix; bool inside; float rx0, rx1, ry0, ry1; };

__device__ __forceinline__ bool item_setup(const GsdRenderParams &p, int item, int warp, int lane, ItemInfo &I) {
    const int n_items = *p.n_items;            // the two loads are independent (item < max_items: in bounds): issued together
    const int4 rec = p.item_tile[item];
    if (item >= n_items) return false;
    I.tile = rec.x;
    I.chunk = rec.y;
    I.start = rec.z;
    I.cnt = rec.w;
    const int tx = I.tile % p.gx, ty = I.tile / p.gx;
    const int wx0 = tx * GSD_TILE + (warp & 1) * 8, wy0 = ty * GSD_TILE + (warp >> 1) * 4;
    I.px = wx0 + (lane & 7);
    I.py = wy0 + (lane >> 3);
    I.inside = I.px < p.W && I.py < p.H;
    I.pix = warp * 32 + lane;
    I.rx0 = (float)wx0; I.rx1 = (float)(wx0 + 7); I.ry0 = (float)wy0; I.ry1 = (float)(wy0 + 3);
    return true;
}

template <int NPLANES>
__device__ __forceinline__ void load_chunk(const GsdRenderParams &p, float4 (*planes)[GSD_CHUNK], uint64_t *bar, int start, int cnt) {
    const uint32_t bytes = (uint32_t)cnt * 16u;
    mbar_expect_tx(bar, bytes * NPLANES);
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) bulk_g2s(&planes[k][0], p.planes + (int64_t)k * p.plane_stride + start, bytes, bar);
}

// ------------------------------------------------------------------------------------------------------
// forward A1: local composite of one chunk
// ------------------------------------------------------------------------------------------------------
// (Measured alternative, not kept: launching A1 in phases by chunk index so that later chunks know the transmittance left by
// the earlier ones — exact termination in chunks 0/1 without A2, upper-bound skipping behind them.  In the benchmark scene 29 %
// of the items lie behind the termination point of every pixel of their tile, but all of them at chunk index >= 4, where only a
// sequential dependence could expose them; the phases cost more in lost occupancy (26 + 25 + 39 us) than the single launch (59 us).)
// Execution order and look-back.  blockIdx.x is mapped to a work item through p.exec_item, which lists the items by CHUNK INDEX
// first (all first chunks of all tiles, then all second chunks, ...): CTAs are dispatched in blockIdx order, so when a deep
// chunk starts, the chunks in front of it have usually finished.  Every warp publishes a flag per (item, 8x4 rectangle) after
// writing its composite; a warp of chunk c first reads the flags of the chunks in front of it, takes the contiguous published
// prefix 0..j, rebuilds T = prod P_k (the same multiplications, in the same order, as the termination pass A2) and, where that
// product is already below 1e-4, switches the pixel off: the reference's loop stopped in an earlier chunk, nothing of this chunk
// is ever read for that pixel (A2 declares it dead at the same k, the combine pass breaks at the terminating chunk, the backward
// prefix stops at n_contrib).  The look-back never waits — an unpublished predecessor just shortens the prefix — so there is no
// forward-progress dependence between CTAs, and skipping never changes a result bit: it only removes work the sequential
// reference would not have done either.  In the 100k benchmark scene 64 % of the (rectangle, chunk) pairs lie behind the
// termination depth of their rectangle (53 % of the chunks behind that of their whole tile).
__device__ __forceinline__ int ld_acquire_s32(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int32_t *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int CH>
__global__ void __launch_bounds__(GSD_CWARPS * 32, 6)
gsd_blend_fwd_chunk_kernel(GsdRenderParams p) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    constexpr int NPL = (CH == 3) ? 3 : 4;
    using IS = ItemState<CH>;
    __shared__ __align__(128) float4 planes[NPL][GSD_CHUNK];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    ItemInfo I;
    const int n_items_l = *p.n_items;          // independent of the next load (exec_item is max_items long): issued together
    const int item = p.exec_item[blockIdx.x];
    if ((int)blockIdx.x >= n_items_l) return;
    if (!item_setup(p, item, warp, lane, I)) return;
    const bool gone = !I.inside;
    float *st = p.chunk_state + (size_t)item * IS::NF * 256;
    // look-back (issued before the gather so that its loads overlap it)
    bool lb_dead = false;
    if (I.chunk > 0) {
        const int item0 = item - I.chunk;
        const int npred = min(I.chunk, 32);
        const int ready = lane < npred ? ld_acquire_s32(p.chunk_flags + (size_t)(item0 + lane) * GSD_CWARPS + warp) : 0;
        const unsigned m = __ballot_sync(0xffffffffu, ready != 0);
        __syncwarp();   // memory ordering between the lane that acquired flag k and the lanes that read chunk k's composite
        const int nready = min(npred, (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1));
        float T = 1.0f;
        bool d = gone;
        for (int c0 = 0; c0 < nready; c0 += 8) {
            if (__all_sync(0xffffffffu, d)) break;
            float Pc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                Pc[u] = (c0 + u < nready) ? __ldcg(p.chunk_state + (size_t)(item0 + c0 + u) * IS::NF * 256 + IS::P * 256 + I.pix) : 1.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (c0 + u < nready && !d) {
                    const float Tn = __fmul_rn(T, Pc[u]);
                    if (Tn < T_EPS) d = true;
                    T = Tn;
                }
            }
        }
        lb_dead = d && !gone;
    }
    // the whole tile is already opaque in front of this chunk: nothing will ever composite its records; the backward only needs
    // every instance's partial-gradient slot (it zero-fills the slots of instances that reached no pixel)
    const bool cta_dead = I.chunk > 0 && __syncthreads_and((lb_dead || gone) ? 1 : 0);
    // gather the chunk's records by sorted key into shared memory AND into the global record planes (the backward and
    // the termination pass stream them with TMA). One record per thread; the gathers of 8 CTAs per SM overlap.
    for (int i = t; i < I.cnt; i += GSD_CWARPS * 32) {
        const uint32_t g = (uint32_t)(p.keys[I.start + i] & 0xffffffffull);
        const uint2 rc = p.g_rect[g];
        const int minx = rc.x & 0xffff, miny = rc.x >> 16, maxx = rc.y & 0xffff;
        const int tx = I.tile % p.gx, ty = I.tile / p.gx;
        const uint32_t slot = p.g_slot_base[g] + (uint32_t)((ty - miny) * (maxx - minx) + (tx - minx));
        if (cta_dead) {
            p.planes_w[3 * p.plane_stride + (int64_t)I.start + i] = make_float4(__uint_as_float(slot), 0.f, 0.f, 0.f);
            continue;
        }
        const float2 pc = p.g_xy[g];
        const float2 e = p.g_ext[g];
        const float4 co = p.g_conic_o[g];
        const float d = p.g_depth[g];
        const float c0 = p.colors0[3 * g], c1 = p.colors0[3 * g + 1], c2 = p.colors0[3 * g + 2];
        float c3 = 0.f, c4 = 0.f, c5 = 0.f;
        if (CH == 6) { c3 = p.colors1[3 * g]; c4 = p.colors1[3 * g + 1]; c5 = p.colors1[3 * g + 2]; }
        const float4 r0 = make_float4(pc.x, pc.y, e.x, e.y), r2 = make_float4(c0, c1, c2, d);
        const float4 r3 = make_float4(__uint_as_float(slot), c3, c4, c5);
        const int64_t j = (int64_t)I.start + i;
        planes[0][i] = r0; planes[1][i] = co; planes[2][i] = r2;
        if (NPL == 4) planes[NPL - 1][i] = r3;
        p.planes_w[j] = r0;
        p.planes_w[p.plane_stride + j] = co;
        p.planes_w[2 * p.plane_stride + j] = r2;
        p.planes_w[3 * p.plane_stride + j] = r3;
    }
    int *term_i = reinterpret_cast<int *>(p.term_state + (size_t)I.tile * TermState<CH>::NF * 256);
    const bool first = I.chunk == 0; // T_in = 1 is known: the reference's termination rule is applied right here
    __syncthreads();
    const float pxf = (float)I.px, pyf = (float)I.py;
    float T = 1.0f, D = 0.f;
    float C[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) C[c] = 0.f;
    int last = 0;
    bool off = gone || lb_dead;   // pixel outside the image, already opaque before this chunk (look-back), or (chunk 0) stopped here
    bool stopped = false;         // chunk 0 only: the reference's "done"
    bool dead = off;              // nothing of this chunk can be used any more (A2 replays the terminating chunk of a pixel)
    // two instantiations of the chunk loop: only chunk 0 carries the termination rule
    auto run = [&](auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        for (int grp = 0; grp < I.cnt; grp += GSD_SUB) {
            if (__all_sync(0xffffffffu, dead)) break;
            const int idx = grp + lane;
            bool pass = false;
            if (idx < I.cnt) {
                pass = cull_pass(planes[0][idx], planes[1][idx], I.rx0, I.rx1, I.ry0, I.ry1);
            }
            unsigned m = __ballot_sync(0xffffffffu, pass);
            while (m) {
                int j[GSD_ILP];
                bool valid[GSD_ILP];
                float alpha[GSD_ILP], power[GSD_ILP];
                float4 col[GSD_ILP], ext[GSD_ILP];
    #pragma unroll
                for (int u = 0; u < GSD_ILP; ++u) {
                    valid[u] = m != 0;
                    j[u] = valid[u] ? (grp + __ffs(m) - 1) : grp;
                    m &= m - 1;
                }
    #pragma unroll
                for (int u = 0; u < GSD_ILP; ++u) {
                    const float4 g0 = planes[0][j[u]];
                    const float4 g1 = planes[1][j[u]];
                    col[u] = planes[2][j[u]];
                    if (CH == 6) ext[u] = planes[NPL - 1][j[u]];
                    power[u] = gsd_power(g1.x, g1.y, g1.z, g0.x - pxf, g0.y - pyf);
                    alpha[u] = fminf(0.99f, __fmul_rn(g1.w, gsd_gauss(power[u])));
                }
    #pragma unroll
                for (int u = 0; u < GSD_ILP; ++u) {
                    // straight-line (predicated) update: a rejected Gaussian contributes with weight 0
                    bool ok = valid[u] && (!off) && (power[u] <= 0.0f) && (alpha[u] >= 1.0f / 255.0f);
                    const float om = __fsub_rn(1.0f, alpha[u]);
                    const float Tn = __fmul_rn(T, om);
                    if (FIRST) {
                        const bool stop = ok && Tn < T_EPS;
                        stopped = stopped || stop;
                        off = off || stop;
                        ok = ok && !stop;
                    }
                    const float w = ok ? alpha[u] * T : 0.f;
                    C[0] += col[u].x * w;
                    C[1] += col[u].y * w;
                    C[2] += col[u].z * w;
                    if (CH == 6) {
                        C[3 % CH] += ext[u].y * w;
                        C[4 % CH] += ext[u].z * w;
                        C[5 % CH] += ext[u].w * w;
                    }
                    D += col[u].w * w;
                    T = ok ? Tn : T;
                    last = ok ? j[u] + 1 : last;
                }
            }
            dead = off || (T < T_EPS);
        }
    };
    if (first) run(std::true_type{});
    else run(std::false_type{});
    if (first) {
        using TS = TermState<CH>;
        float *ts = p.term_state + (size_t)I.tile * TS::NF * 256;
        if (stopped) {
            ts[TS::T * 256 + I.pix] = T;
            ts[TS::D * 256 + I.pix] = D;
#pragma unroll
            for (int c = 0; c < CH; ++c) ts[(TS::C + c) * 256 + I.pix] = C[c];
            term_i[TS::LAST * 256 + I.pix] = last;
        }
        term_i[TS::CSTAR * 256 + I.pix] = stopped ? 0 : TERMINAL_NONE;
        if (stopped) T = 0.f; // later chunks see a dead pixel
    }
    if (lb_dead) T = 0.f;   // switched off by the look-back: later chunks / passes see a dead pixel (C, D, last are still 0)
    st[IS::P * 256 + I.pix] = T;
    st[IS::D * 256 + I.pix] = D;
#pragma unroll
    for (int c = 0; c < CH; ++c) st[(IS::C + c) * 256 + I.pix] = C[c];
    reinterpret_cast<int *>(st)[IS::LAST * 256 + I.pix] = last;
    // publish this rectangle's composite for the look-back of the chunks behind it (2: the whole rectangle was already
    // opaque in front of this chunk — the termination pass skips it without rebuilding T)
    const bool rect_dead = __all_sync(0xffffffffu, lb_dead || gone);
    __syncwarp();   // orders the lanes' composite stores before lane 0's fence + flag
    if (lane == 0) {
        __threadfence();
        st_release_s32(p.chunk_flags + (size_t)item * GSD_CWARPS + warp, rect_dead && I.chunk > 0 ? 2 : 1);
    }
}

// ------------------------------------------------------------------------------------------------------
// forward B: per tile — combine the chunk composites in order and find the chunk in which each pixel's transmittance crosses 1e-4
// ------------------------------------------------------------------------------------------------------
// One pass over a tile's chunk composites (round 1 rebuilt T_in = prod P of ALL preceding chunks inside every work item of the
// termination pass: O(chunks^2) loads per tile):
//   1. fold chunk after chunk: C += T C_c, D += T D_c, T *= P_c (the loads of four chunks are in flight at a time); the first
//      chunk c > 0 with T * P_c < 1e-4 is the pixel's terminating chunk: the reference's loop stops somewhere inside it
//      (chunk 0 applies the stop rule itself: its T_in = 1 is known);
//   2. pixels that do not terminate behind chunk 0: write colour (+ T * background), depth, final T and n_contrib;
//      pixels that do: hand (terminating chunk, incoming T, composite so far) to pass C, which replays that chunk for them with
//      the reference's per-Gaussian test T (1 - alpha) < 1e-4 and writes their outputs.
// (Measured and not kept: replaying inside this kernel — per-tile serialisation of up to ~20 distinct terminating chunks made it
// 79-96 us against 47 us for the two launches it replaced; the replays are independent across items and belong in their own grid.)
template <int CH>
__global__ void __launch_bounds__(GSD_CWARPS * 32)
gsd_blend_fwd_finish_kernel(GsdRenderParams p) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    using IS = ItemState<CH>;
    using TS = TermState<CH>;
    constexpr int NPL = (CH == 3) ? 3 : 4;
    constexpr int BATCH = 4;   // chunks whose composites are in flight at once (8: 124 registers, 7.0 -> 9.5 us)
    const int tile = blockIdx.x;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tx = tile % p.gx, ty = tile / p.gx;
    const int wx0 = tx * GSD_TILE + (warp & 1) * 8, wy0 = ty * GSD_TILE + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const int item0 = p.chunk_ptr[tile];
    const int nc = min(p.chunk_ptr[tile + 1], p.max_items) - item0;
    if (nc <= 0) {   // empty tile: background only (uniform for the CTA)
        if (inside) {
            const size_t pid = (size_t)py * p.W + px;
            const size_t plane = (size_t)p.W * p.H;
            p.final_T[pid] = 1.0f;
            p.n_contrib[pid] = 0;
#pragma unroll
            for (int c = 0; c < CH; ++c) p.out_color[c * plane + pid] = (c < 3) ? __ldg(p.bg0 + c) : (p.bg1 ? __ldg(p.bg1 + (c % 3)) : 0.f);
            p.out_depth[pid] = 0.f;
        }
        return;
    }
    float T = 1.0f, D = 0.f;
    float C[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) C[c] = 0.f;
    int last = 0;
    int cstar = -1;       // chunk (> 0) in which this pixel's T crosses 1e-4; T then holds its incoming transmittance
    if (nc > 0 && inside) {
        const float *ts = p.term_state + (size_t)tile * TS::NF * 256;
        if (reinterpret_cast<const int *>(ts)[TS::CSTAR * 256 + t] == 0) {
            // stopped inside chunk 0: A1 applied the exact rule and left the absolute values
            const int lc = reinterpret_cast<const int *>(ts)[TS::LAST * 256 + t];
#pragma unroll
            for (int k = 0; k < CH; ++k) C[k] = ts[(TS::C + k) * 256 + t];
            D = ts[TS::D * 256 + t];
            T = ts[TS::T * 256 + t];
            last = lc;
        } else {
            bool open = true;
            for (int c0 = 0; c0 < nc && open; c0 += BATCH) {
                float bP[BATCH], bD[BATCH], bC[BATCH][CH];
                int bl[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {   // independent loads in flight
                    const bool v = c0 + u < nc;
                    const float *st = p.chunk_state + (size_t)(item0 + (v ? c0 + u : 0)) * IS::NF * 256;
                    bl[u] = v ? reinterpret_cast<const int *>(st)[IS::LAST * 256 + t] : 0;
                    bP[u] = v ? st[IS::P * 256 + t] : 1.0f;
                    bD[u] = v ? st[IS::D * 256 + t] : 0.f;
#pragma unroll
                    for (int k = 0; k < CH; ++k) bC[u][k] = v ? st[(IS::C + k) * 256 + t] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const int c = c0 + u;
                    if (c < nc && open) {
                        const float Tn = __fmul_rn(T, bP[u]);
                        if (c > 0 && Tn < T_EPS) {   // the reference stops inside this chunk: replay it below
                            cstar = c;
                            open = false;
                        } else if (bl[u] > 0) {
#pragma unroll
                            for (int k = 0; k < CH; ++k) C[k] += T * bC[u][k];
                            D += T * bD[u];
                            T = Tn;
                            last = c * GSD_CHUNK + bl[u];
                        }
                    }
                }
            }
        }
    }
    // work list of the replay pass: one entry per (work item, rectangle) that holds a terminating pixel — a rectangle is this
    // warp, so the lanes that share a terminating chunk elect one of them to append.  The list lives in the look-back flag
    // array of pass A, which is dead by now (this kernel runs after it; the next forward clears it again).
    {
        const bool term = inside && cstar >= 0;
        const unsigned tm = __ballot_sync(0xffffffffu, term);
        if (term) {
            const unsigned same = __match_any_sync(tm, cstar);
            if (lane == __ffs(same) - 1) p.chunk_flags[atomicAdd(p.replay_count, 1)] = (item0 + cstar) * GSD_CWARPS + warp;
        }
    }
    if (!inside) return;
    float *tsw = p.term_state + (size_t)tile * TS::NF * 256;
    reinterpret_cast<int *>(tsw)[TS::CSTAR * 256 + t] = cstar;     // -1, or the chunk the replay pass finishes this pixel in
    if (cstar >= 0) {
        // hand the pixel over to the replay of chunk cstar: incoming transmittance and the composite of the chunks in front
        tsw[TS::T * 256 + t] = T;
        tsw[TS::D * 256 + t] = D;
#pragma unroll
        for (int k = 0; k < CH; ++k) tsw[(TS::C + k) * 256 + t] = C[k];
        reinterpret_cast<int *>(tsw)[TS::LAST * 256 + t] = last;
        return;
    }
    const size_t pid = (size_t)py * p.W + px;
    const size_t plane = (size_t)p.W * p.H;
    p.final_T[pid] = T;
    p.n_contrib[pid] = last;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float bgc = (c < 3) ? __ldg(p.bg0 + c) : (p.bg1 ? __ldg(p.bg1 + (c % 3)) : 0.f);
        p.out_color[c * plane + pid] = C[c] + T * bgc;
    }
    p.out_depth[pid] = D;
}

// ------------------------------------------------------------------------------------------------------
// forward C: replay of a terminating chunk for the pixels of one 8x4 rectangle that stop inside it
// ------------------------------------------------------------------------------------------------------
// Every pixel terminates in at most one chunk, so every output pixel has exactly one writer (pass B or one warp of this pass).
// One WARP per (work item, rectangle) pair of pass B's list (~1.5 pairs per rectangle of a saturated tile): warps are independent
// (their own 2 KB shared-memory stage for the 32 records of a cull group, no block barrier, no TMA), so every resident warp has
// work.  (Round 2's first version ran one CTA per work item with the chunk staged by TMA: one to three live warps per 8-warp
// CTA, SMSPs active 56 % of the elapsed cycles, 24 us; the order of the list is arbitrary and never affects a result.)
#ifndef GSD_REPLAY_MIN_CTAS
#define GSD_REPLAY_MIN_CTAS 4   // 63 registers; 5 / 6 CTAs per SM (48 / 40 registers, 24 / 72 bytes of spills) measured +2.9 / +4.4 us per iteration
#endif
template <int CH>
__global__ void __launch_bounds__(GSD_CWARPS * 32, GSD_REPLAY_MIN_CTAS)
gsd_blend_fwd_replay_kernel(GsdRenderParams p) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    constexpr int NPL = (CH == 3) ? 3 : 4;
    using TS = TermState<CH>;
    __shared__ __align__(16) float4 stage[GSD_CWARPS][NPL][GSD_SUB];
    const int t = threadIdx.x, wid = t >> 5, lane = t & 31;
    const int n_pairs = *p.replay_count;
    const int pi = blockIdx.x * GSD_CWARPS + wid;
    if (pi >= n_pairs) return;
    const int pair = p.chunk_flags[pi];
    const int item = pair / GSD_CWARPS, rect = pair % GSD_CWARPS;
    ItemInfo I;
    if (!item_setup(p, item, rect, lane, I)) return;
    float *ts = p.term_state + (size_t)I.tile * TS::NF * 256;
    const bool mine = I.inside && reinterpret_cast<const int *>(ts)[TS::CSTAR * 256 + I.pix] == I.chunk;
    float T = 1.0f, D = 0.f;
    float C[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) C[c] = 0.f;
    int last = 0;
    if (mine) {
        T = ts[TS::T * 256 + I.pix];
        D = ts[TS::D * 256 + I.pix];
#pragma unroll
        for (int c = 0; c < CH; ++c) C[c] = ts[(TS::C + c) * 256 + I.pix];
        last = reinterpret_cast<const int *>(ts)[TS::LAST * 256 + I.pix];
    }
    // sequential replay for the terminating lanes: the reference's loop (renderCUDA forward) with the true incoming T
    const float pxf = (float)I.px, pyf = (float)I.py;
    const float4 *src = p.planes + I.start;
    float4 (*st)[GSD_SUB] = stage[wid];
    int lastl = 0;
    bool done = !mine;
    // lane l holds record grp + l of every plane (coalesced 512-byte reads, next group in flight during the current one)
    float4 r[NPL];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NPL; ++k) r[k] = lane < I.cnt ? src[(int64_t)k * p.plane_stride + lane] : zero4;
    for (int grp = 0; grp < I.cnt; grp += GSD_SUB) {
        if (__all_sync(0xffffffffu, done)) break;
        const int idx = grp + lane;
        const bool pass = idx < I.cnt && cull_pass(r[0], r[1], I.rx0, I.rx1, I.ry0, I.ry1);
        __syncwarp();      // the previous group's survivors have been read by every lane
#pragma unroll
        for (int k = 0; k < NPL; ++k) st[k][lane] = r[k];
        const int nidx = idx + GSD_SUB;
#pragma unroll
        for (int k = 0; k < NPL; ++k) r[k] = nidx < I.cnt ? src[(int64_t)k * p.plane_stride + nidx] : zero4;
        __syncwarp();      // stage stores visible to the whole warp
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int jl = __ffs(m) - 1;
            m &= m - 1;
            const float4 g0 = st[0][jl];
            const float4 g1 = st[1][jl];
            const float4 g2 = st[2][jl];
            float4 g3 = zero4;
            if (CH == 6) g3 = st[NPL - 1][jl];
            const float power = gsd_power(g1.x, g1.y, g1.z, g0.x - pxf, g0.y - pyf);
            const float alpha = fminf(0.99f, __fmul_rn(g1.w, gsd_gauss(power)));
            // straight-line (predicated) update, as in pass A
            bool ok = (!done) && (power <= 0.0f) && (alpha >= 1.0f / 255.0f);
            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
            const bool stop = ok && test_T < T_EPS;
            done = done || stop;
            ok = ok && !stop;
            const float w = ok ? alpha * T : 0.f;
            C[0] += g2.x * w;
            C[1] += g2.y * w;
            C[2] += g2.z * w;
            if (CH == 6) {
                C[3 % CH] += g3.y * w;
                C[4 % CH] += g3.z * w;
                C[5 % CH] += g3.w * w;
            }
            D += g2.w * w;
            T = ok ? test_T : T;
            lastl = ok ? grp + jl + 1 : lastl;
        }
    }
    if (!mine) return;
    if (lastl > 0) last = I.chunk * GSD_CHUNK + lastl;   // 0: nothing of this chunk contributed before the stop
    const size_t pid = (size_t)I.py * p.W + I.px;
    const size_t plane = (size_t)p.W * p.H;
    p.final_T[pid] = T;
    p.n_contrib[pid] = last;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float bgc = (c < 3) ? __ldg(p.bg0 + c) : (p.bg1 ? __ldg(p.bg1 + (c % 3)) : 0.f);
        p.out_color[c * plane + pid] = C[c] + T * bgc;
    }
    p.out_depth[pid] = D;
}


// ------------------------------------------------------------------------------------------------------
// backward B': per chunk and pixel, the incoming transmittance and the colour composited in front of the chunk
// ------------------------------------------------------------------------------------------------------
// Depends only on the forward's results (chunk composites, n_contrib) — NOT on dL/dcolour: the tracking iteration runs it on a side
// branch while the photometric kernels compute the image gradient, and the chunk kernel forms
// Q_in = dL/dC . (colour still to come) = dL/dC . (C_final - T_final bg - C_pre) itself.  (Round 2's first version folded
// dL/dC in here — one scalar per chunk and pixel instead of CH — and therefore sat on the critical path behind the SSIM gradient.)
template <int CH>
__global__ void __launch_bounds__(GSD_CWARPS * 32)
gsd_blend_bwd_prefix_kernel(GsdRenderParams p) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    using IS = ItemState<CH>;
    const int tile = blockIdx.x;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tx = tile % p.gx, ty = tile / p.gx;
    const int px = tx * GSD_TILE + (warp & 1) * 8 + (lane & 7), py = ty * GSD_TILE + (warp >> 1) * 4 + (lane >> 3);
    const int item0 = p.chunk_ptr[tile];
    const int nc = min(p.chunk_ptr[tile + 1], p.max_items) - item0;
    if (nc <= 0) return;
    const bool inside = px < p.W && py < p.H;
    const int last = inside ? p.n_contrib[(size_t)py * p.W + px] : 0;
    float T = 1.0f;
    float Cp[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) Cp[k] = 0.f;
    float nP = 0.f, nC[CH];
    int nl = 0;
    auto fetch = [&](int c) {
        const float *st = p.chunk_state + (size_t)(item0 + c) * IS::NF * 256;
        nl = reinterpret_cast<const int *>(st)[IS::LAST * 256 + t];
        nP = st[IS::P * 256 + t];
#pragma unroll
        for (int k = 0; k < CH; ++k) nC[k] = st[(IS::C + k) * 256 + t];
    };
    fetch(0);
    for (int c = 0; c < nc; ++c) {
        float *st = p.chunk_state + (size_t)(item0 + c) * IS::NF * 256;
        st[IS::TIN * 256 + t] = T;
#pragma unroll
        for (int k = 0; k < CH; ++k) st[(IS::CPRE + k) * 256 + t] = Cp[k];
        if (c * GSD_CHUNK >= last) { // nothing of this or later chunks contributed to the pixel (the chunk kernel skips them by n_contrib)
            for (int c2 = c + 1; c2 < nc; ++c2) p.chunk_state[(size_t)(item0 + c2) * IS::NF * 256 + IS::TIN * 256 + t] = 0.f;
            break;
        }
        const float cP = nP;
        const int lc = nl;
        float cC[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) cC[k] = nC[k];
        if (c + 1 < nc) fetch(c + 1);
        if (lc > 0) {   // the same fold as the forward's pass B
#pragma unroll
            for (int k = 0; k < CH; ++k) Cp[k] += T * cC[k];
            T = __fmul_rn(T, cP);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// backward A': gradients of one chunk
// ------------------------------------------------------------------------------------------------------
// GEOM: only the geometry gradients (mean2D, conic) are produced — the steady state of the tracker freezes colours and
// opacities (train_utils.py:370-373), which shrinks the per-Gaussian warp reduction from 12 to 5 values.
// (The same half-tile split was measured for the forward chunk and termination kernels: 3 440 / 3 475 it/s against 3 504 — not kept.)
// HALVES = 2 (geometry mode): an item is processed by two CTAs of four consumer warps + flusher, one per half tile; their
// five sums go to floats [0,5) and [8,13) of the same 64-byte record (summed by the preprocess backward).  A CTA lives as long
// as its slowest rectangle, so half-tile CTAs idle less behind uneven rectangles and pack twice as finely.
template <int CH, bool GEOM, int HALVES>
__global__ void __launch_bounds__((GSD_CWARPS / HALVES + 1) * 32, HALVES == 2 ? 8 : 5)
gsd_blend_bwd_chunk_kernel(GsdRenderParams p) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    constexpr int NW = GSD_CWARPS / HALVES; // consumer warps of this CTA
    constexpr int NV = GEOM ? 5 : CH + 6; // [colours,] mean2D(2), conic(3) [, opacity(1)]
    constexpr int OG = GEOM ? 0 : CH;     // offset of the geometry values inside a partial record
    constexpr int S = GEOM ? 4 : (CH == 3 ? 3 : 2); // ring of per-warp partial-sum stages (static shared memory budget); 4 = a whole chunk: consumers never wait for the flusher
    using IS = ItemState<CH>;
    __shared__ __align__(128) float4 planes[4][GSD_CHUNK];
    __shared__ float acc[S][NW][GSD_SUB][NV];
    __shared__ unsigned wmask[S][NW];
    __shared__ __align__(8) uint64_t bar, done_bar[S], empty[S];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    ItemInfo I;
    // heavy items first: the forward's chunk-index-major order is also longest-work-first for this kernel (front chunks touch
    // every pixel, deep chunks lie mostly behind the pixels' last contributor and leave early), so the grid's tail is made of
    // light items instead of whatever tile happens to come last
    const int half = blockIdx.x % HALVES;
    if ((int)(blockIdx.x / HALVES) >= *p.n_items) return;
    const int item = p.exec_item[blockIdx.x / HALVES];
    if (!item_setup(p, item, half * NW + (warp < NW ? warp : 0), lane, I)) return;
    if (t == 0) {
        mbar_init(&bar, 1);
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&done_bar[s], NW * GSD_RING_ARRIVALS); mbar_init(&empty[s], GSD_RING_ARRIVALS); }
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) load_chunk<4>(p, planes, &bar, I.start, I.cnt);
    const int nsub = (I.cnt + GSD_SUB - 1) / GSD_SUB;

    if (warp == NW) {
        // ===== flusher: fixed-order sum over the consumer warps, one 64-byte partial record per instance =====
        mbar_wait(&bar, 0);
        for (int sb = 0; sb < nsub; ++sb) {
            const int s = sb % S;
            mbar_wait(&done_bar[s], (sb / S) & 1);
            const int cnt = min(GSD_SUB, I.cnt - sb * GSD_SUB);
            unsigned wm[NW], any = 0u;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) { wm[w2] = wmask[s][w2]; any |= wm[w2]; }
            // the preprocess backward reads the first NVP floats of a record (8 in geometry-only mode, 12 otherwise)
            constexpr int NVP = GEOM ? 8 : 12;
            if (any == 0u) {
                // nothing of this sub-batch reached any pixel (e.g. a chunk behind the termination depth): zero records
                if (lane < cnt) {
                    const uint32_t slot = __float_as_uint(planes[3][sb * GSD_SUB + lane].x);
                    if ((int64_t)slot < p.plane_stride) {
                        float4 *r = reinterpret_cast<float4 *>(p.partials + (size_t)slot * GSD_PART_FLOATS + half * NVP);
                        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int k = 0; k < NVP / 4; ++k) r[k] = z;
                    }
                }
            } else {
                for (int idx = lane; idx < cnt * NVP; idx += 32) {
                    const int j = idx / NVP, vv = idx % NVP;
                    float sum = 0.f;
                    if (vv < NV) {
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2)
                            if ((wm[w2] >> j) & 1u) sum += acc[s][w2][j][vv];
                    }
                    const uint32_t slot = __float_as_uint(planes[3][sb * GSD_SUB + j].x);
                    if ((int64_t)slot < p.plane_stride) p.partials[(size_t)slot * GSD_PART_FLOATS + half * NVP + vv] = sum;
                }
            }
            __syncwarp();
            if (GSD_RING_ARRIVALS == 32 || lane == 0) mbar_arrive(&empty[s]);
        }
        return;
    }
    // ===== consumers =====
    const float pxf = (float)I.px, pyf = (float)I.py;
    const int my_val = holder_id(NV, lane);
    const int base = I.chunk * GSD_CHUNK; // index of the chunk's first record in the tile list
    float T = 0.f, Q = 0.f, tail = 0.f;
    float dLdC[CH];
    int last = 0;
    if (I.inside) {
        const size_t pid = (size_t)I.py * p.W + I.px;
        const size_t plane = (size_t)p.W * p.H;
        const float *st = p.chunk_state + (size_t)item * IS::NF * 256;
        T = st[IS::TIN * 256 + I.pix];
        last = p.n_contrib[pid];
        const float Tfin = p.final_T[pid];
        float bgdot = 0.f;
        // Q = dL/dC . (colour composited behind the front of this chunk) = dL/dC . (C_final - T_final bg - C_pre)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const float bgc = (c < 3) ? __ldg(p.bg0 + c) : (p.bg1 ? __ldg(p.bg1 + (c % 3)) : 0.f);
            dLdC[c] = p.dL_dcolor[c * plane + pid];
            bgdot += bgc * dLdC[c];
            Q += dLdC[c] * ((p.out_color[c * plane + pid] - Tfin * bgc) - st[(IS::CPRE + c) * 256 + I.pix]);
        }
        tail = Tfin * bgdot;
    } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) dLdC[c] = 0.f;
    }
    int wlast = last;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) wlast = max(wlast, __shfl_xor_sync(0xffffffffu, wlast, o));
    mbar_wait(&bar, 0);
    for (int sb = 0; sb < nsub; ++sb) {
        const int s = sb % S;
        if (sb >= S) mbar_wait(&empty[s], ((sb / S) - 1) & 1);
        const int grp = sb * GSD_SUB;
        unsigned touched = 0u;
        if (base + grp < wlast) {
            const int idx = grp + lane;
            bool pass = false;
            if (idx < I.cnt && base + idx < wlast) {
                pass = cull_pass(planes[0][idx], planes[1][idx], I.rx0, I.rx1, I.ry0, I.ry1);
            }
            unsigned m = __ballot_sync(0xffffffffu, pass);
            while (m) {
                const int jl = __ffs(m) - 1;
                m &= m - 1;
                const int j = grp + jl;
                const float4 g0 = planes[0][j];
                const float4 g1 = planes[1][j];
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = gsd_power(g1.x, g1.y, g1.z, dx, dy);
                const float Gr = gsd_gauss(power);
                const float alpha = fminf(0.99f, __fmul_rn(g1.w, Gr));
                const bool ok = (base + j < last) && (power <= 0.0f) && (alpha >= 1.0f / 255.0f);
                if (!__any_sync(0xffffffffu, ok)) continue;
                const float4 g2 = planes[2][j];
                float col[CH];
                col[0] = g2.x; col[1] = g2.y; col[2] = g2.z;
                if (CH == 6) {
                    const float4 g3 = planes[3][j];
                    col[3 % CH] = g3.y; col[4 % CH] = g3.z; col[5 % CH] = g3.w;
                }
                const float one_m = __fsub_rn(1.0f, alpha);
                const float w = ok ? alpha * T : 0.f;
                float cd = 0.f;
#pragma unroll
                for (int c = 0; c < CH; ++c) cd += col[c] * dLdC[c];
                const float Qn = Q - cd * w;
                const float dL_dalpha = ok ? (T * cd - (Qn + tail) * gsd_rcp_approx(one_m)) : 0.f;   // one_m in [0.01, 1]
                const float Ge = ok ? Gr : 0.f;
                float v[NV];
                if (!GEOM) {
#pragma unroll
                    for (int c = 0; c < CH; ++c) v[c % NV] = w * dLdC[c];
                    v[(CH + 5) % NV] = Ge * dL_dalpha;
                }
                // geometry: raw moments of s = dL/dG * G; the (per-Gaussian) conic is applied once in the preprocess backward
                const float sg = g1.w * dL_dalpha * Ge;
                const float sdx = sg * dx, sdy = sg * dy;
                v[OG + 0] = sdx;
                v[OG + 1] = sdy;
                v[OG + 2] = sdx * dx;
                v[OG + 3] = sdx * dy;
                v[OG + 4] = sdy * dy;
                if (ok) {
                    Q = Qn;
                    T = __fmul_rn(T, one_m);
                }
                xreduce<NV, 16>(v, lane);
                if (my_val >= 0) acc[s][warp][jl][my_val] = v[0];
                touched |= 1u << jl;
            }
        }
        if (lane == 0) wmask[s][warp] = touched;
        __syncwarp();
        if (GSD_RING_ARRIVALS == 32 || lane == 0) mbar_arrive(&done_bar[s]);
    }
}

// ------------------------------------------------------------------------------------------------------
size_t gsd_chunk_state_floats(int n_sets, int max_items) { return (size_t)(4 + 6 * n_sets) * 256 * (size_t)max_items; }
size_t gsd_term_state_floats(int n_sets, int tiles) { return (size_t)(4 + 3 * n_sets) * 256 * (size_t)tiles; }

int gsd_launch_render_fwd(const GsdRenderParams &p, int tiles, int n_sets, cudaStream_t st) {
    if (tiles == 0) return GSD_OK;
    const int threads = GSD_CWARPS * 32;
    if (p.max_items > 0) {
        if (n_sets == 1) gsd_launch((gsd_blend_fwd_chunk_kernel<3>), dim3(p.max_items), dim3(threads), 0, st, p);
        else gsd_launch((gsd_blend_fwd_chunk_kernel<6>), dim3(p.max_items), dim3(threads), 0, st, p);
        GSD_LAUNCH_CHECK();
    }
    if (n_sets == 1) gsd_launch((gsd_blend_fwd_finish_kernel<3>), dim3(tiles), dim3(threads), 0, st, p);
    else gsd_launch((gsd_blend_fwd_finish_kernel<6>), dim3(tiles), dim3(threads), 0, st, p);
    GSD_LAUNCH_CHECK();
    if (p.max_items > 0) {
        if (n_sets == 1) gsd_launch((gsd_blend_fwd_replay_kernel<3>), dim3(p.max_items), dim3(threads), 0, st, p);
        else gsd_launch((gsd_blend_fwd_replay_kernel<6>), dim3(p.max_items), dim3(threads), 0, st, p);
    }
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// which: 1 = the per-chunk prefix pass (needs only the forward's state), 2 = the chunk kernel, 3 = both
int gsd_launch_render_bwd(const GsdRenderParams &p, int tiles, int n_sets, int which, cudaStream_t st) {
    if (tiles == 0 || p.max_items == 0) return GSD_OK;
    if (which & 1) {
        if (n_sets == 1) gsd_launch((gsd_blend_bwd_prefix_kernel<3>), dim3(tiles), dim3(GSD_CWARPS * 32), 0, st, p);
        else gsd_launch((gsd_blend_bwd_prefix_kernel<6>), dim3(tiles), dim3(GSD_CWARPS * 32), 0, st, p);
        GSD_LAUNCH_CHECK();
    }
    if (!(which & 2)) return GSD_OK;
    const int threads = (GSD_CWARPS + 1) * 32;
    if (p.geom_only) {   // half-tile CTAs (see the kernel)
        const int th2 = (GSD_CWARPS / 2 + 1) * 32;
        if (n_sets == 1) gsd_launch((gsd_blend_bwd_chunk_kernel<3, true, 2>), dim3(2 * p.max_items), dim3(th2), 0, st, p);
        else gsd_launch((gsd_blend_bwd_chunk_kernel<6, true, 2>), dim3(2 * p.max_items), dim3(th2), 0, st, p);
    } else {
        if (n_sets == 1) gsd_launch((gsd_blend_bwd_chunk_kernel<3, false, 1>), dim3(p.max_items), dim3(threads), 0, st, p);
        else gsd_launch((gsd_blend_bwd_chunk_kernel<6, false, 1>), dim3(p.max_items), dim3(threads), 0, st, p);
    }
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
