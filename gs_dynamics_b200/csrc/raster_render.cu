// raster_render.cu — the blend kernels (forward and backward) of the tile rasterizer, sm_100a.
//
// Replaces renderCUDA<3> forward/backward of the upstream rasterizer, reached from
// /root/reference/src/tracking/train_utils.py:178,192 (forward) and train_gs.py:31 (backward).
//
// Design (B200-first, not a translation):
//   * persistent CTAs pull 16x16 tiles from a work queue ordered by descending instance count (LPT schedule written by
//     the binning pass), so the heaviest tiles start first and the tail is short;
//   * warp-specialised: one PRODUCER warp streams the tile's depth-sorted record planes into a ring of shared-memory
//     stages with the TMA engine (cp.async.bulk + mbarrier full/empty pairs); 8 CONSUMER warps each own an 8x4 pixel
//     rectangle and never meet at a CTA-wide barrier inside a tile; (backward) one FLUSHER warp sums the per-warp
//     partials of a finished stage in fixed order and writes one 64-byte record per instance;
//   * every 32 records the lanes of a consumer warp test one record each against the warp's rectangle (conservative
//     extents of the alpha >= 1/255 ellipse) and only the survivors of the ballot are blended — exact, ~2.3x fewer pairs;
//   * survivors are processed GSD_ILP at a time: the alpha evaluations (the long dependent chains: LDS -> FMA x6 ->
//     MUFU.EX2) are independent and overlap; only the short transmittance recurrence is serial;
//   * backward runs FRONT-TO-BACK like the forward (suffix colour = final colour - prefix), so T is rebuilt by the same
//     multiplications as in the forward; the per-pixel partials of a Gaussian are summed across the warp with a transposed
//     butterfly (13 shuffles for 12 values): no atomics anywhere, bit-reproducible gradients.
#include "common.cuh"

#define GSD_BATCH 32     // records per pipeline stage
#define GSD_STAGES_F 6   // forward ring depth
#define GSD_STAGES_B 3   // backward ring depth (each stage also carries the per-warp partial sums)
#define GSD_CWARPS 8     // consumer warps
#define GSD_ILP 4

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NPLANES>
__device__ __forceinline__ void issue_batch(const GsdRenderParams &p, float4 (*stage)[GSD_BATCH], uint64_t *bar,
                                            uint32_t start, int cnt) {
    const uint32_t bytes = (uint32_t)cnt * 16u;
    mbar_expect_tx(bar, bytes * NPLANES);
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) bulk_g2s(&stage[k][0], p.planes + (int64_t)k * p.plane_stride + start, bytes, bar);
}

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__((GSD_CWARPS + 1) * 32)
gsd_render_fwd_kernel(GsdRenderParams p) {
    constexpr int NPL = (CH == 3) ? 3 : 4;
    constexpr int S = GSD_STAGES_F;
    __shared__ __align__(128) float4 stage[S][NPL][GSD_BATCH];
    __shared__ __align__(8) uint64_t full[S], empty[S];
    __shared__ int s_tile;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GSD_CWARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t gb = 0; // batches streamed so far by this CTA (identical in every warp): stage = gb % S, use = gb / S

    for (;;) {
        if (t == 0) s_tile = atomicAdd(p.next_tile, 1);
        __syncthreads();
        const int q = s_tile;
        __syncthreads();
        if (q >= p.n_tiles) break;
        const int tile = p.tile_order[q];
        const uint2 range = p.ranges[tile];
        const int n = (int)(range.y - range.x);
        const int nb = (n + GSD_BATCH - 1) / GSD_BATCH;

        if (warp == GSD_CWARPS) {
            // ===== producer =====
            if (lane == 0) {
                for (int b = 0; b < nb; ++b) {
                    const uint32_t g = gb + b;
                    const int s = g % S;
                    if (g >= S) mbar_wait(&empty[s], ((g / S) - 1) & 1);
                    issue_batch<NPL>(p, stage[s], &full[s], range.x + b * GSD_BATCH, min(GSD_BATCH, n - b * GSD_BATCH));
                }
            }
        } else {
            // ===== consumers =====
            const int tx = tile % p.gx, ty = tile / p.gx;
            const int wx0 = tx * GSD_TILE + (warp & 1) * 8, wy0 = ty * GSD_TILE + (warp >> 1) * 4;
            const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
            const float pxf = (float)px, pyf = (float)py;
            const bool inside = px < p.W && py < p.H;
            const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);
            bool done = !inside;
            float T = 1.0f, D = 0.f;
            float C[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) C[c] = 0.f;
            int last = 0;
            for (int b = 0; b < nb; ++b) {
                const uint32_t g = gb + b;
                const int s = g % S;
                mbar_wait(&full[s], (g / S) & 1);
                const int cnt = min(GSD_BATCH, n - b * GSD_BATCH);
                if (!__all_sync(0xffffffffu, done)) {
                    bool pass = false;
                    if (lane < cnt) {
                        const float4 g0 = stage[s][0][lane];
                        pass = (g0.x + g0.z >= rx0) && (g0.x - g0.z <= rx1) && (g0.y + g0.w >= ry0) && (g0.y - g0.w <= ry1);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, pass);
                    while (m) {
                        int j[GSD_ILP];
                        bool valid[GSD_ILP];
                        float alpha[GSD_ILP], power[GSD_ILP];
                        float4 col[GSD_ILP];
#pragma unroll
                        for (int u = 0; u < GSD_ILP; ++u) {
                            valid[u] = m != 0;
                            j[u] = valid[u] ? (__ffs(m) - 1) : 0;
                            m &= m - 1;
                        }
#pragma unroll
                        for (int u = 0; u < GSD_ILP; ++u) {
                            const float4 g0 = stage[s][0][j[u]];
                            const float4 g1 = stage[s][1][j[u]];
                            col[u] = stage[s][2][j[u]];
                            power[u] = gsd_power(g1.x, g1.y, g1.z, g0.x - pxf, g0.y - pyf);
                            alpha[u] = fminf(0.99f, __fmul_rn(g1.w, gsd_gauss(power[u])));
                        }
#pragma unroll
                        for (int u = 0; u < GSD_ILP; ++u) {
                            bool ok = valid[u] && (!done) && (power[u] <= 0.0f) && (alpha[u] >= 1.0f / 255.0f);
                            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha[u]));
                            if (ok && test_T < 0.0001f) {
                                done = true;
                                ok = false;
                            }
                            if (ok) {
                                const float w = alpha[u] * T;
                                C[0] += col[u].x * w;
                                C[1] += col[u].y * w;
                                C[2] += col[u].z * w;
                                if (CH == 6) {
                                    const float4 g3 = stage[s][NPL - 1][j[u]];
                                    C[3 % CH] += g3.y * w;
                                    C[4 % CH] += g3.z * w;
                                    C[5 % CH] += g3.w * w;
                                }
                                D += col[u].w * w;
                                T = test_T;
                                last = b * GSD_BATCH + j[u] + 1;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            if (inside) {
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
        }
        gb += nb;
    }
}

// ------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------
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

template <int CH>
__global__ void __launch_bounds__((GSD_CWARPS + 2) * 32, 3)
gsd_render_bwd_kernel(GsdRenderParams p) {
    constexpr int NV = CH + 6; // colours, mean2D(2), conic(3), opacity(1)
    constexpr int S = GSD_STAGES_B;
    __shared__ __align__(128) float4 stage[S][4][GSD_BATCH];
    __shared__ float acc[S][GSD_CWARPS][GSD_BATCH][NV];
    __shared__ unsigned wmask[S][GSD_CWARPS];
    __shared__ __align__(8) uint64_t full[S], done_bar[S], empty[S];
    __shared__ int s_tile;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&done_bar[s], GSD_CWARPS); mbar_init(&empty[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t gb = 0;
    const int my_val = holder_id(NV, lane);
    const float ddelx_dx = 0.5f * p.W, ddely_dy = 0.5f * p.H;

    for (;;) {
        if (t == 0) s_tile = atomicAdd(p.next_tile, 1);
        __syncthreads();
        const int q = s_tile;
        __syncthreads();
        if (q >= p.n_tiles) break;
        const int tile = p.tile_order[q];
        const uint2 range = p.ranges[tile];
        const int n = (int)(range.y - range.x);
        const int nb = (n + GSD_BATCH - 1) / GSD_BATCH;

        if (warp == GSD_CWARPS) {
            // ===== producer =====
            if (lane == 0) {
                for (int b = 0; b < nb; ++b) {
                    const uint32_t g = gb + b;
                    const int s = g % S;
                    if (g >= S) mbar_wait(&empty[s], ((g / S) - 1) & 1);
                    issue_batch<4>(p, stage[s], &full[s], range.x + b * GSD_BATCH, min(GSD_BATCH, n - b * GSD_BATCH));
                }
            }
        } else if (warp == GSD_CWARPS + 1) {
            // ===== flusher: fixed-order sum over the consumer warps, one 64-byte partial record per instance =====
            for (int b = 0; b < nb; ++b) {
                const uint32_t g = gb + b;
                const int s = g % S;
                mbar_wait(&done_bar[s], (g / S) & 1);
                const int cnt = min(GSD_BATCH, n - b * GSD_BATCH);
                unsigned wm[GSD_CWARPS];
#pragma unroll
                for (int w2 = 0; w2 < GSD_CWARPS; ++w2) wm[w2] = wmask[s][w2];
                for (int idx = lane; idx < cnt * GSD_PART_FLOATS; idx += 32) {
                    const int j = idx / GSD_PART_FLOATS, vv = idx % GSD_PART_FLOATS;
                    float sum = 0.f;
                    if (vv < NV) {
#pragma unroll
                        for (int w2 = 0; w2 < GSD_CWARPS; ++w2)
                            if ((wm[w2] >> j) & 1u) sum += acc[s][w2][j][vv];
                    }
                    const uint32_t slot = __float_as_uint(stage[s][3][j].x);
                    if ((int64_t)slot < p.plane_stride) p.partials[(size_t)slot * GSD_PART_FLOATS + vv] = sum;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        } else {
            // ===== consumers =====
            const int tx = tile % p.gx, ty = tile / p.gx;
            const int wx0 = tx * GSD_TILE + (warp & 1) * 8, wy0 = ty * GSD_TILE + (warp >> 1) * 4;
            const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
            const float pxf = (float)px, pyf = (float)py;
            const bool inside = px < p.W && py < p.H;
            float T = 1.0f, Tfin = 0.f, Q = 0.f, bgdot = 0.f;
            float dLdC[CH];
            int last = 0;
            if (inside && n > 0) {
                const size_t pid = (size_t)py * p.W + px;
                const size_t plane = (size_t)p.W * p.H;
                Tfin = p.final_T[pid];
                last = p.n_contrib[pid];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const float bgc = (c < 3) ? __ldg(p.bg0 + c) : (p.bg1 ? __ldg(p.bg1 + (c % 3)) : 0.f);
                    dLdC[c] = p.dL_dcolor[c * plane + pid];
                    Q += dLdC[c] * (p.out_color[c * plane + pid] - Tfin * bgc);
                    bgdot += bgc * dLdC[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < CH; ++c) dLdC[c] = 0.f;
            }
            const float tail = Tfin * bgdot;
            int wlast = last;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) wlast = max(wlast, __shfl_xor_sync(0xffffffffu, wlast, o));
            const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);

            for (int b = 0; b < nb; ++b) {
                const uint32_t g = gb + b;
                const int s = g % S;
                mbar_wait(&full[s], (g / S) & 1);
                const int cnt = min(GSD_BATCH, n - b * GSD_BATCH);
                unsigned touched = 0u;
                if (b * GSD_BATCH < wlast) {
                    bool pass = false;
                    if (lane < cnt && b * GSD_BATCH + lane < wlast) {
                        const float4 g0 = stage[s][0][lane];
                        pass = (g0.x + g0.z >= rx0) && (g0.x - g0.z <= rx1) && (g0.y + g0.w >= ry0) && (g0.y - g0.w <= ry1);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, pass);
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        const float4 g0 = stage[s][0][j];
                        const float4 g1 = stage[s][1][j];
                        const float dx = g0.x - pxf, dy = g0.y - pyf;
                        const float power = gsd_power(g1.x, g1.y, g1.z, dx, dy);
                        const float Gr = gsd_gauss(power);
                        const float alpha = fminf(0.99f, __fmul_rn(g1.w, Gr));
                        const bool ok = (b * GSD_BATCH + j < last) && (power <= 0.0f) && (alpha >= 1.0f / 255.0f);
                        if (!__any_sync(0xffffffffu, ok)) continue;
                        const float4 g2 = stage[s][2][j];
                        float col[CH];
                        col[0] = g2.x; col[1] = g2.y; col[2] = g2.z;
                        if (CH == 6) {
                            const float4 g3 = stage[s][3][j];
                            col[3 % CH] = g3.y; col[4 % CH] = g3.z; col[5 % CH] = g3.w;
                        }
                        const float one_m = __fsub_rn(1.0f, alpha);
                        const float w = ok ? alpha * T : 0.f;
                        float cd = 0.f;
#pragma unroll
                        for (int c = 0; c < CH; ++c) cd += col[c] * dLdC[c];
                        const float Qn = Q - cd * w;
                        const float dL_dalpha = ok ? (T * cd - __fdividef(Qn + tail, one_m)) : 0.f;
                        const float Ge = ok ? Gr : 0.f;
                        float v[NV];
#pragma unroll
                        for (int c = 0; c < CH; ++c) v[c] = w * dLdC[c];
                        const float dL_dG = g1.w * dL_dalpha;
                        const float gdx = Ge * dx, gdy = Ge * dy;
                        v[CH + 0] = dL_dG * (-gdx * g1.x - gdy * g1.y) * ddelx_dx;
                        v[CH + 1] = dL_dG * (-gdy * g1.z - gdx * g1.y) * ddely_dy;
                        v[CH + 2] = -0.5f * gdx * dx * dL_dG;
                        v[CH + 3] = -0.5f * gdx * dy * dL_dG;
                        v[CH + 4] = -0.5f * gdy * dy * dL_dG;
                        v[CH + 5] = Ge * dL_dalpha;
                        if (ok) {
                            Q = Qn;
                            T = __fmul_rn(T, one_m);
                        }
                        xreduce<NV, 16>(v, lane);
                        if (my_val >= 0) acc[s][warp][j][my_val] = v[0];
                        touched |= 1u << j;
                    }
                }
                if (lane == 0) wmask[s][warp] = touched;
                __syncwarp();
                if (lane == 0) mbar_arrive(&done_bar[s]);
            }
        }
        gb += nb;
    }
}

// ------------------------------------------------------------------------------------------------------
template <typename K>
static int blend_grid(K kernel, int threads, int tiles) {
    int dev = 0, sms = 148, per_sm = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int g = sms * per_sm;
    return tiles < g ? tiles : g;
}

int gsd_launch_render_fwd(const GsdRenderParams &p, int tiles, int n_sets, cudaStream_t st) {
    if (tiles == 0) return GSD_OK;
    const int threads = (GSD_CWARPS + 1) * 32;
    static int grid3 = 0, grid6 = 0;
    if (n_sets == 1) {
        if (!grid3) grid3 = blend_grid(gsd_render_fwd_kernel<3>, threads, 1 << 30);
        gsd_render_fwd_kernel<3><<<tiles < grid3 ? tiles : grid3, threads, 0, st>>>(p);
    } else {
        if (!grid6) grid6 = blend_grid(gsd_render_fwd_kernel<6>, threads, 1 << 30);
        gsd_render_fwd_kernel<6><<<tiles < grid6 ? tiles : grid6, threads, 0, st>>>(p);
    }
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

int gsd_launch_render_bwd(const GsdRenderParams &p, int tiles, int n_sets, cudaStream_t st) {
    if (tiles == 0) return GSD_OK;
    const int threads = (GSD_CWARPS + 2) * 32;
    static int grid3 = 0, grid6 = 0;
    if (n_sets == 1) {
        if (!grid3) grid3 = blend_grid(gsd_render_bwd_kernel<3>, threads, 1 << 30);
        gsd_render_bwd_kernel<3><<<tiles < grid3 ? tiles : grid3, threads, 0, st>>>(p);
    } else {
        if (!grid6) grid6 = blend_grid(gsd_render_bwd_kernel<6>, threads, 1 << 30);
        gsd_render_bwd_kernel<6><<<tiles < grid6 ? tiles : grid6, threads, 0, st>>>(p);
    }
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
