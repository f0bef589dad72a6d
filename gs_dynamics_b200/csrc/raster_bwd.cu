// raster_bwd.cu — per-Gaussian backward: fixed-order sum of the per-instance partial records written by the
// blend backward, then conic -> cov2D -> (cov3D, mean) and NDC mean2D -> mean3D chain rule, cov3D -> scale/quat.
//
// Replaces computeCov2DCUDA + preprocessCUDA (backward.cu) of the upstream rasterizer;
// gradients returned as the reference's autograd consumes them (/root/reference/src/tracking/train_gs.py:31,
// means2D.grad at /root/reference/src/tracking/external.py:138-142). dL/ddepth is ignored as upstream does.
#include "common.cuh"
#include "track_update.cuh"

// FUSE (steady-state tracking, geometry-only): the gradients do not go to memory — the thread applies normalize backward +
// gradient sum + Adam to its Gaussian right here (track_update.cuh), with the radii bookkeeping and the step advance of
// gsd_track_update: one launch and a 28-byte gradient round trip per Gaussian less per iteration.
template <int CH, bool GEOM, bool FUSE>
__device__ __forceinline__ void
preprocess_bwd_body(int G, const GsdCam &cam, int64_t capacity, const float *__restrict__ means3D,
                    const float *__restrict__ scales, const float *__restrict__ rotations,
                    const int32_t *__restrict__ radii, const uint32_t *__restrict__ slot_base,
                    const uint32_t *__restrict__ tiles, const float4 *__restrict__ conic_o,
                    const float4 *__restrict__ partials, float *__restrict__ dmeans3D, float *__restrict__ dmeans2D, float *__restrict__ dcolors0,
                    float *__restrict__ dcolors1, float *__restrict__ dopac, float *__restrict__ dscales,
                    float *__restrict__ drot, const GsdTrackUpdate &u) {
    __shared__ float sVP[32];
    __shared__ GsdAdamCoef coef;
    if (threadIdx.x < 16) sVP[threadIdx.x] = cam.view[threadIdx.x];
    else if (threadIdx.x < 32) sVP[threadIdx.x] = cam.proj[threadIdx.x - 16];
    if (FUSE && threadIdx.x == 32) gsd_track_update_coef(u, &coef);
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    const bool vis = radii[i] > 0;
    if (vis) {
        uint32_t nt = tiles[i];
        int64_t s0 = (int64_t)slot_base[i];
        if (GEOM) {
            // two half-tile partial sums per record: floats [0,5) and [8,13), added in fixed order.  The trip count is data dependent,
            // so the compiler does not overlap the iterations: the loads of four records are issued together (the kernel is one wave;
            // its time is the slowest thread's chain of record loads), the sums keep the record order
            for (uint32_t k0 = 0; k0 < nt; k0 += 4) {
                float4 A[4], Cc[4];
                float B[4], D[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t sidx = s0 + k0 + u;
                    ok[u] = (k0 + u < nt) && sidx < capacity;
                    const float *r = reinterpret_cast<const float *>(partials + (ok[u] ? sidx : 0) * 4);
                    A[u] = *reinterpret_cast<const float4 *>(r);
                    B[u] = r[4];
                    Cc[u] = *reinterpret_cast<const float4 *>(r + 8);
                    D[u] = r[12];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!ok[u]) continue;
                    acc[0] += A[u].x + Cc[u].x; acc[1] += A[u].y + Cc[u].y; acc[2] += A[u].z + Cc[u].z; acc[3] += A[u].w + Cc[u].w;
                    acc[4] += B[u] + D[u];
                }
            }
        } else {
            for (uint32_t k = 0; k < nt; ++k) {
                int64_t s = s0 + k;
                if (s >= capacity) break;
                const float4 *r = partials + s * 4;
                float4 a = r[0], b = r[1];
                acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
                float4 c = r[2];
                acc[8] += c.x; acc[9] += c.y; acc[10] += c.z; acc[11] += c.w;
            }
        }
    }
    // layout of a partial record: colours[CH], five geometry MOMENTS, opacity.  With s = dL/dG * G per (pixel, instance) and
    // d = centre - pixel, the blend backward accumulates M1 = sum s dx, M2 = sum s dy, M3 = sum s dx^2, M4 = sum s dx dy,
    // M5 = sum s dy^2; the conic is per Gaussian, so the map to dL/dmean2D and dL/dconic is applied once here instead of
    // per (pixel, instance) there:  dL/dmean2D = -(W/2) (A M1 + B M2), -(H/2) (C M2 + B M1);  dL/dconic = -1/2 (M3, M4, M5)
    constexpr int OG = GEOM ? 0 : CH;
    float dm2x = 0.f, dm2y = 0.f, dca = 0.f, dcb = 0.f, dcc = 0.f;
    if (vis) {
        const float4 co = conic_o[i];
        const float M1 = acc[OG], M2 = acc[OG + 1];
        dm2x = -(0.5f * cam.W) * (co.x * M1 + co.y * M2);
        dm2y = -(0.5f * cam.H) * (co.z * M2 + co.y * M1);
        dca = -0.5f * acc[OG + 2]; dcb = -0.5f * acc[OG + 3]; dcc = -0.5f * acc[OG + 4];
    }
    if (!GEOM) {
        if (dcolors0) { dcolors0[3 * i] = acc[0]; dcolors0[3 * i + 1] = acc[1]; dcolors0[3 * i + 2] = acc[2]; }
        if (CH == 6 && dcolors1) { dcolors1[3 * i] = acc[3]; dcolors1[3 * i + 1] = acc[4]; dcolors1[3 * i + 2] = acc[5]; }
        if (dopac) dopac[i] = acc[(CH + 5) % 12];
    }
    if (dmeans2D) { dmeans2D[3 * i] = dm2x; dmeans2D[3 * i + 1] = dm2y; dmeans2D[3 * i + 2] = 0.f; }
    float dmean[3] = {0.f, 0.f, 0.f}, dsc[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    if (vis) {
        const float *V = sVP, *P = sVP + 16;
        const float px_ = means3D[3 * i], py_ = means3D[3 * i + 1], pz_ = means3D[3 * i + 2];
        // recompute Sigma, M = J Rw, cov2D
        const float r = rotations[4 * i], x = rotations[4 * i + 1], y = rotations[4 * i + 2], z = rotations[4 * i + 3];
        const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                               {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                               {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float mod = cam.scale_modifier;
        const float sv[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
        float S[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                S[a][b] = R[a][0] * sv[0] * sv[0] * R[b][0] + R[a][1] * sv[1] * sv[1] * R[b][1] + R[a][2] * sv[2] * sv[2] * R[b][2];
        float tvx = V[0] * px_ + V[4] * py_ + V[8] * pz_ + V[12];
        float tvy = V[1] * px_ + V[5] * py_ + V[9] * pz_ + V[13];
        float tvz = V[2] * px_ + V[6] * py_ + V[10] * pz_ + V[14];
        const float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
        const float txtz = tvx / tvz, tytz = tvy / tvz;
        const float xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float tx = fminf(limx, fmaxf(-limx, txtz)) * tvz, ty = fminf(limy, fmaxf(-limy, tytz)) * tvz;
        const float fx = cam.focal_x, fy = cam.focal_y;
        const float J[2][3] = {{fx / tvz, 0.f, -(fx * tx) / (tvz * tvz)}, {0.f, fy / tvz, -(fy * ty) / (tvz * tvz)}};
        float M[2][3], MS[2][3];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) M[a][b] = J[a][0] * V[b * 4 + 0] + J[a][1] * V[b * 4 + 1] + J[a][2] * V[b * 4 + 2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) MS[a][b] = M[a][0] * S[0][b] + M[a][1] * S[1][b] + M[a][2] * S[2][b];
        const float ca = MS[0][0] * M[0][0] + MS[0][1] * M[0][1] + MS[0][2] * M[0][2] + 0.3f;
        const float cb = MS[0][0] * M[1][0] + MS[0][1] * M[1][1] + MS[0][2] * M[1][2];
        const float cc = MS[1][0] * M[1][0] + MS[1][1] * M[1][1] + MS[1][2] * M[1][2] + 0.3f;
        const float denom = ca * cc - cb * cb;
        const float d2i = 1.0f / (denom * denom + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        if (d2i != 0.f) {
            dL_da = d2i * (-cc * cc * dca + 2.f * cb * cc * dcb + (denom - ca * cc) * dcc);
            dL_dc = d2i * (-ca * ca * dcc + 2.f * ca * cb * dcb + (denom - ca * cc) * dca);
            dL_db = d2i * 2.f * (cb * cc * dca - (denom + 2.f * cb * cb) * dcb + ca * cb * dcc);
        }
        // dL/dSigma as a symmetric matrix Gs = M^T Gm M, Gm = [[da, db/2],[db/2, dc]]
        float Gs[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                Gs[a][b] = M[0][a] * M[0][b] * dL_da + 0.5f * (M[0][a] * M[1][b] + M[1][a] * M[0][b]) * dL_db + M[1][a] * M[1][b] * dL_dc;
        // dL/dM = 2 Gm M Sigma
        float dM[2][3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            dM[0][b] = 2.f * MS[0][b] * dL_da + MS[1][b] * dL_db;
            dM[1][b] = 2.f * MS[1][b] * dL_dc + MS[0][b] * dL_db;
        }
        const float dJ00 = dM[0][0] * V[0] + dM[0][1] * V[4] + dM[0][2] * V[8];
        const float dJ02 = dM[0][0] * V[2] + dM[0][1] * V[6] + dM[0][2] * V[10];
        const float dJ11 = dM[1][0] * V[1] + dM[1][1] * V[5] + dM[1][2] * V[9];
        const float dJ12 = dM[1][0] * V[2] + dM[1][1] * V[6] + dM[1][2] * V[10];
        const float tz = 1.f / tvz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = xmul * -fx * tz2 * dJ02;
        const float dty = ymul * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * tx) * tz3 * dJ02 + (2.f * fy * ty) * tz3 * dJ12;
        dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
        dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
        dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
        // NDC mean2D -> mean3D through the perspective divide
        const float hw = P[3] * px_ + P[7] * py_ + P[11] * pz_ + P[15];
        const float mw = 1.0f / (hw + 0.0000001f);
        const float mul1 = (P[0] * px_ + P[4] * py_ + P[8] * pz_ + P[12]) * mw * mw;
        const float mul2 = (P[1] * px_ + P[5] * py_ + P[9] * pz_ + P[13]) * mw * mw;
        dmean[0] += (P[0] * mw - P[3] * mul1) * dm2x + (P[1] * mw - P[3] * mul2) * dm2y;
        dmean[1] += (P[4] * mw - P[7] * mul1) * dm2x + (P[5] * mw - P[7] * mul2) * dm2y;
        dmean[2] += (P[8] * mw - P[11] * mul1) * dm2x + (P[9] * mw - P[11] * mul2) * dm2y;
        // Sigma = R D R^T
        float GR[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) GR[a][b] = Gs[a][0] * R[0][b] + Gs[a][1] * R[1][b] + Gs[a][2] * R[2][b];
        float g[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float rgr = R[0][k] * GR[0][k] + R[1][k] * GR[1][k] + R[2][k] * GR[2][k];
            dsc[k] = 2.f * sv[k] * rgr * mod;
#pragma unroll
            for (int a = 0; a < 3; ++a) g[a][k] = 2.f * GR[a][k] * sv[k] * sv[k];
        }
        dq[0] = 2.f * (-z * g[0][1] + y * g[0][2] + z * g[1][0] - x * g[1][2] - y * g[2][0] + x * g[2][1]);
        dq[1] = 2.f * (y * g[0][1] + z * g[0][2] + y * g[1][0] - 2.f * x * g[1][1] - r * g[1][2] + z * g[2][0] + r * g[2][1] - 2.f * x * g[2][2]);
        dq[2] = 2.f * (-2.f * y * g[0][0] + x * g[0][1] + r * g[0][2] + x * g[1][0] + z * g[1][2] - r * g[2][0] + z * g[2][1] - 2.f * y * g[2][2]);
        dq[3] = 2.f * (-2.f * z * g[0][0] - r * g[0][1] + x * g[0][2] + r * g[1][0] - 2.f * z * g[1][1] + y * g[1][2] + x * g[2][0] + y * g[2][1]);
    }
    if (FUSE) {
        gsd_track_update_radii_1(u, i);
        gsd_track_update_apply(u, coef, i, dmean, make_float4(dq[0], dq[1], dq[2], dq[3]));
        return;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dmeans3D[3 * i + k] = dmean[k];
        dscales[3 * i + k] = dsc[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) drot[4 * i + k] = dq[k];
}

template <int CH, bool GEOM>
__global__ void __launch_bounds__(256)
gsd_preprocess_bwd_kernel(int G, GsdCam cam, int64_t capacity, const float *__restrict__ means3D,
                          const float *__restrict__ scales, const float *__restrict__ rotations,
                          const int32_t *__restrict__ radii, const uint32_t *__restrict__ slot_base,
                          const uint32_t *__restrict__ tiles, const float4 *__restrict__ conic_o,
                          const float4 *__restrict__ partials, float *__restrict__ dmeans3D, float *__restrict__ dmeans2D, float *__restrict__ dcolors0,
                          float *__restrict__ dcolors1, float *__restrict__ dopac, float *__restrict__ dscales,
                          float *__restrict__ drot) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    GsdTrackUpdate none;
    preprocess_bwd_body<CH, GEOM, false>(G, cam, capacity, means3D, scales, rotations, radii, slot_base, tiles, conic_o, partials, dmeans3D, dmeans2D,
                                         dcolors0, dcolors1, dopac, dscales, drot, none);
}

template <int CH>
__global__ void __launch_bounds__(256)
gsd_preprocess_bwd_update_kernel(int G, GsdCam cam, int64_t capacity, const float *__restrict__ means3D,
                                 const float *__restrict__ scales, const float *__restrict__ rotations,
                                 const int32_t *__restrict__ radii, const uint32_t *__restrict__ slot_base,
                                 const uint32_t *__restrict__ tiles, const float4 *__restrict__ conic_o,
                                 const float4 *__restrict__ partials, GsdTrackUpdate u) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    preprocess_bwd_body<CH, true, true>(G, cam, capacity, means3D, scales, rotations, radii, slot_base, tiles, conic_o, partials, nullptr, nullptr,
                                        nullptr, nullptr, nullptr, nullptr, nullptr, u);
    gsd_track_update_advance(u);
}

// per-Gaussian backward fused with the tracker's update (gsd_track_backward_update): geometry-only, no gradient outputs
int gsd_launch_preprocess_bwd_update(int G, const GsdCam &cam, const GsdRasterBwd *a, const GsdGeomWs &g, const GsdTrackUpdate &u, cudaStream_t st) {
    if (G == 0) return GSD_OK;
    const GsdRasterFwd &f = a->fwd;
    int blocks = (G + 255) / 256;
    if (f.n_sets == 1)
        gsd_launch((gsd_preprocess_bwd_update_kernel<3>), dim3(blocks), dim3(256), 0, st, G, cam, f.capacity, f.means3D, f.scales, f.rotations, f.radii,
                   g.slot_base, g.tiles, g.conic_o, (const float4 *)a->partial_ws, u);
    else
        gsd_launch((gsd_preprocess_bwd_update_kernel<6>), dim3(blocks), dim3(256), 0, st, G, cam, f.capacity, f.means3D, f.scales, f.rotations, f.radii,
                   g.slot_base, g.tiles, g.conic_o, (const float4 *)a->partial_ws, u);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

int gsd_launch_preprocess_bwd(int G, const GsdCam &cam, const GsdRasterBwd *a, const GsdGeomWs &g, int geom_only, cudaStream_t st) {
    if (G == 0) return GSD_OK;
    const GsdRasterFwd &f = a->fwd;
    int blocks = (G + 255) / 256;
#define GSD_PB_ARGS G, cam, f.capacity, f.means3D, f.scales, f.rotations, f.radii, g.slot_base, g.tiles, g.conic_o, (const float4 *)a->partial_ws, \
        a->dL_dmeans3D, a->dL_dmeans2D, a->dL_dcolors0, a->dL_dcolors1, a->dL_dopacities, a->dL_dscales, a->dL_drotations
    if (geom_only) {
        if (f.n_sets == 1) gsd_launch((gsd_preprocess_bwd_kernel<3, true>), dim3(blocks), dim3(256), 0, st, GSD_PB_ARGS);
        else gsd_launch((gsd_preprocess_bwd_kernel<6, true>), dim3(blocks), dim3(256), 0, st, GSD_PB_ARGS);
    } else {
        if (f.n_sets == 1) gsd_launch((gsd_preprocess_bwd_kernel<3, false>), dim3(blocks), dim3(256), 0, st, GSD_PB_ARGS);
        else gsd_launch((gsd_preprocess_bwd_kernel<6, false>), dim3(blocks), dim3(256), 0, st, GSD_PB_ARGS);
    }
#undef GSD_PB_ARGS
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
