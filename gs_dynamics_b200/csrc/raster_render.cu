// raster_render.cu — the blend kernels (forward and backward) of the tile rasterizer, sm_100a.
//
// Replaces renderCUDA<3> forward/backward of the upstream rasterizer, reached from
// /root/reference/src/tracking/train_utils.py:178,192 (forward) and train_gs.py:31 (backward).
//
// Design (B200-first, not a translation):
//   * one CTA (8 warps) per 16x16 tile, one pixel per thread, each warp owns an 8x4 pixel rectangle;
//   * the tile's depth-sorted instance list is a contiguous run of packed records (4 float4 planes) that is
//     streamed into shared memory by the TMA engine (cp.async.bulk + mbarrier, GSD_STAGES-deep ring);
//   * every 32 records the lanes test one record each against the warp's rectangle (conservative extents of
//     the alpha >= 1/255 ellipse) and only the survivors of the ballot are blended — exact, ~2.3x fewer pairs;
//   * backward runs FRONT-TO-BACK like the forward (suffix colour = final colour - prefix), so T is rebuilt
//     by the same multiplications as in the forward instead of divisions; the per-pixel partials of a
//     Gaussian are summed across the warp with a transposed butterfly (13 shuffles for 12 values), across
//     warps in fixed order in shared memory, and written to a per-instance slot: no atomics, bit-reproducible.
#include "common.cuh"

#define GSD_BATCH 64
#define GSD_STAGES 2
#define GSD_WARPS 8


template <int NPLANES>
__device__ __forceinline__ void issue_batch(const GsdRenderParams &p, float4 (*stage)[GSD_BATCH], uint64_t *bar,
                                            uint32_t start, int cnt) {
    uint32_t bytes = (uint32_t)cnt * 16u;
    mbar_expect_tx(bar, bytes * NPLANES);
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) bulk_g2s(&stage[k][0], p.planes + (int64_t)k * p.plane_stride + start, bytes, bar);
}

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256)
gsd_render_fwd_kernel(GsdRenderParams p) {
    constexpr int NPL = (CH == 3) ? 3 : 4;
    __shared__ __align__(128) float4 stage[GSD_STAGES][4][GSD_BATCH];
    __shared__ __align__(8) uint64_t full[GSD_STAGES];

    const int tile = blockIdx.x;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tx = tile % p.gx, ty = tile / p.gx;
    const int wx0 = tx * GSD_TILE + (warp & 1) * 8, wy0 = ty * GSD_TILE + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
    const float pxf = (float)px, pyf = (float)py;
    const bool inside = px < p.W && py < p.H;
    const uint2 range = p.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nb = (n + GSD_BATCH - 1) / GSD_BATCH;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < GSD_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < GSD_STAGES; ++s)
            if (s < nb) issue_batch<NPL>(p, stage[s], &full[s], range.x + s * GSD_BATCH, min(GSD_BATCH, n - s * GSD_BATCH));
    }

    bool done = !inside;
    float T = 1.0f, D = 0.f;
    float C[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) C[c] = 0.f;
    int last = 0;
    // warp rectangle in pixel-centre coordinates
    const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);

    for (int b = 0; b < nb; ++b) {
        const int s = b % GSD_STAGES;
        mbar_wait(&full[s], (uint32_t)((b / GSD_STAGES) & 1));
        const int cnt = min(GSD_BATCH, n - b * GSD_BATCH);
        if (!__all_sync(0xffffffffu, done)) {
            for (int grp = 0; grp < cnt; grp += 32) {
                const int idx = grp + lane;
                bool pass = false;
                if (idx < cnt) {
                    float4 g0 = stage[s][0][idx];
                    pass = (g0.x + g0.z >= rx0) && (g0.x - g0.z <= rx1) && (g0.y + g0.w >= ry0) && (g0.y - g0.w <= ry1);
                }
                unsigned m = __ballot_sync(0xffffffffu, pass);
                while (m) {
                    const int j = grp + __ffs(m) - 1;
                    m &= m - 1;
                    const float4 g0 = stage[s][0][j];
                    const float4 g1 = stage[s][1][j];
                    const float4 g2 = stage[s][2][j];
                    const float dx = g0.x - pxf, dy = g0.y - pyf;
                    const float power = gsd_power(g1.x, g1.y, g1.z, dx, dy);
                    const float alpha = fminf(0.99f, __fmul_rn(g1.w, gsd_gauss(power)));
                    bool ok = (!done) && (power <= 0.0f) && (alpha >= 1.0f / 255.0f);
                    const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                    if (ok && test_T < 0.0001f) {
                        done = true;
                        ok = false;
                    }
                    if (ok) {
                        const float w = alpha * T;
                        C[0] += g2.x * w;
                        C[1] += g2.y * w;
                        C[2] += g2.z * w;
                        if (CH == 6) {
                            const float4 g3 = stage[s][3][j];
                            C[3 % CH] += g3.y * w;
                            C[4 % CH] += g3.z * w;
                            C[5 % CH] += g3.w * w;
                        }
                        D += g2.w * w;
                        T = test_T;
                        last = b * GSD_BATCH + j + 1;
                    }
                }
            }
        }
        const int all_done = __syncthreads_and(done ? 1 : 0); // also releases stage s
        if (all_done) {
            // drain copies already in flight into our shared memory before the CTA may retire
            for (int b2 = b + 1; b2 < nb && b2 < b + GSD_STAGES; ++b2)
                mbar_wait(&full[b2 % GSD_STAGES], (uint32_t)((b2 / GSD_STAGES) & 1));
            break;
        }
        if (t == 0 && b + GSD_STAGES < nb) {
            const int b2 = b + GSD_STAGES;
            issue_batch<NPL>(p, stage[s], &full[s], range.x + b2 * GSD_BATCH, min(GSD_BATCH, n - b2 * GSD_BATCH));
        }
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

// ------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------
// Transposed butterfly: N per-lane values are summed over the 32 lanes with ~N shuffles; afterwards the lane
// returned by holder_id() == k holds the warp total of value k in v[0].
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
__global__ void __launch_bounds__(256)
gsd_render_bwd_kernel(GsdRenderParams p) {
    constexpr int NV = CH + 6; // colours, mean2D(2), conic(3), opacity(1)
    __shared__ __align__(128) float4 stage[GSD_STAGES][4][GSD_BATCH];
    __shared__ __align__(8) uint64_t full[GSD_STAGES];
    __shared__ float acc[GSD_WARPS][GSD_BATCH][NV];
    __shared__ unsigned long long wmask[GSD_WARPS];

    const int tile = blockIdx.x;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tx = tile % p.gx, ty = tile / p.gx;
    const int wx0 = tx * GSD_TILE + (warp & 1) * 8, wy0 = ty * GSD_TILE + (warp >> 1) * 4;
    const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
    const float pxf = (float)px, pyf = (float)py;
    const bool inside = px < p.W && py < p.H;
    const uint2 range = p.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nb = (n + GSD_BATCH - 1) / GSD_BATCH;
    if (n == 0) return;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < GSD_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < GSD_STAGES; ++s)
            if (s < nb) issue_batch<4>(p, stage[s], &full[s], range.x + s * GSD_BATCH, min(GSD_BATCH, n - s * GSD_BATCH));
    }

    // per-pixel constants
    float T = 1.0f, Tfin = 0.f, Q = 0.f, bgdot = 0.f;
    float dLdC[CH];
    int last = 0;
    if (inside) {
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
    // the largest contributor index of the warp / of the CTA bounds the work
    int wlast = last;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) wlast = max(wlast, __shfl_xor_sync(0xffffffffu, wlast, o));
    const int my_val = holder_id(NV, lane);
    const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);
    const float ddelx_dx = 0.5f * p.W, ddely_dy = 0.5f * p.H;

    int b = 0;
    for (; b < nb; ++b) {
        const int s = b % GSD_STAGES;
        mbar_wait(&full[s], (uint32_t)((b / GSD_STAGES) & 1));
        const int cnt = min(GSD_BATCH, n - b * GSD_BATCH);
        unsigned long long touched = 0ull;
        if (b * GSD_BATCH < wlast) {
            for (int grp = 0; grp < cnt; grp += 32) {
                const int idx = grp + lane;
                bool pass = false;
                if (idx < cnt && b * GSD_BATCH + idx < wlast) {
                    float4 g0 = stage[s][0][idx];
                    pass = (g0.x + g0.z >= rx0) && (g0.x - g0.z <= rx1) && (g0.y + g0.w >= ry0) && (g0.y - g0.w <= ry1);
                }
                unsigned m = __ballot_sync(0xffffffffu, pass);
                while (m) {
                    const int j = grp + __ffs(m) - 1;
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
                    if (my_val >= 0) acc[warp][j][my_val] = v[0];
                    touched |= 1ull << j;
                }
            }
        }
        if (lane == 0) wmask[warp] = touched;
        __syncthreads();
        // flush: fixed-order sum over warps, one 64-byte partial record per instance
        for (int idx = t; idx < cnt * GSD_PART_FLOATS; idx += 256) {
            const int j = idx / GSD_PART_FLOATS, vv = idx % GSD_PART_FLOATS;
            float sum = 0.f;
            if (vv < NV) {
#pragma unroll
                for (int w2 = 0; w2 < GSD_WARPS; ++w2)
                    if ((wmask[w2] >> j) & 1ull) sum += acc[w2][j][vv];
            }
            const uint32_t slot = __float_as_uint(stage[s][3][j].x);
            p.partials[(size_t)slot * GSD_PART_FLOATS + vv] = sum;
        }
        const int cta_more = __syncthreads_or(((b + 1) * GSD_BATCH < wlast) ? 1 : 0); // releases stage s, acc, wmask
        if (!cta_more) {
            for (int b2 = b + 1; b2 < nb && b2 < b + GSD_STAGES; ++b2)
                mbar_wait(&full[b2 % GSD_STAGES], (uint32_t)((b2 / GSD_STAGES) & 1));
            ++b;
            break;
        }
        if (t == 0 && b + GSD_STAGES < nb) {
            const int b2 = b + GSD_STAGES;
            issue_batch<4>(p, stage[s], &full[s], range.x + b2 * GSD_BATCH, min(GSD_BATCH, n - b2 * GSD_BATCH));
        }
    }
    // instances behind the last contributor of every pixel received no gradient: zero their records
    const float4 *plane3 = p.planes + 3 * p.plane_stride;
    for (int64_t idx = (int64_t)b * GSD_BATCH * 4 + t; idx < (int64_t)n * 4; idx += 256) {
        const int64_t j = idx >> 2;
        const uint32_t slot = __float_as_uint(plane3[range.x + j].x);
        reinterpret_cast<float4 *>(p.partials)[(size_t)slot * 4 + (idx & 3)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------------
int gsd_launch_render_fwd(const GsdRenderParams &p, int tiles, int n_sets, cudaStream_t st) {
    if (tiles == 0) return GSD_OK;
    if (n_sets == 1)
        gsd_render_fwd_kernel<3><<<tiles, 256, 0, st>>>(p);
    else
        gsd_render_fwd_kernel<6><<<tiles, 256, 0, st>>>(p);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

int gsd_launch_render_bwd(const GsdRenderParams &p, int tiles, int n_sets, cudaStream_t st) {
    if (tiles == 0) return GSD_OK;
    if (n_sets == 1)
        gsd_render_bwd_kernel<3><<<tiles, 256, 0, st>>>(p);
    else
        gsd_render_bwd_kernel<6><<<tiles, 256, 0, st>>>(p);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
