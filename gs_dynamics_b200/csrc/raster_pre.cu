// raster_pre.cu — per-Gaussian preprocess (forward). Compiled with -fmad=false so that the integer outputs
// (radii, tile rectangles) are reproducible against the fp32 CPU oracle; the kernel is HBM-bound
// (56 B in, 52 B out per Gaussian), arithmetic is irrelevant to its speed.
//
// Replaces preprocessCUDA of the upstream rasterizer (called through
// /root/reference/src/tracking/train_utils.py:178).
#include "common.cuh"

__device__ __forceinline__ float3 xform4x3(const float *m, float3 p) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}

__global__ void __launch_bounds__(256)
gsd_preprocess_kernel(int G, GsdCam cam, const float *__restrict__ means3D, const float *__restrict__ opacities,
                      const float *__restrict__ scales, const float *__restrict__ rotations, float2 *__restrict__ xy,
                      float4 *__restrict__ conic_o, float2 *__restrict__ ext, float *__restrict__ depth,
                      uint2 *__restrict__ rect, uint32_t *__restrict__ tiles, uint32_t *__restrict__ slot_base,
                      int32_t *__restrict__ radii, uint32_t *__restrict__ block_sum, int32_t *__restrict__ zero_fill, int n_zero,
                      const float4 *__restrict__ unnorm, float4 *__restrict__ rot_out) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    __shared__ float sVP[32];
    // the binning histogram accumulates the per-tile totals with atomics: cleared here (one launch earlier in stream order)
    for (int z = blockIdx.x * blockDim.x + threadIdx.x; z < n_zero; z += gridDim.x * blockDim.x) zero_fill[z] = 0;
    if (threadIdx.x < 16) sVP[threadIdx.x] = cam.view[threadIdx.x];
    else if (threadIdx.x < 32) sVP[threadIdx.x] = cam.proj[threadIdx.x - 16];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i < G;
    // steady-state tracking: the quaternions arrive un-normalised (params['unnorm_rotations']); F.normalize is applied here and the
    // result written for the backward (one launch less in front of every iteration)
    float4 qn = make_float4(1.f, 0.f, 0.f, 0.f);
    if (in_range) {
        if (unnorm) { qn = gsd_quat_normalize(unnorm[i]); rot_out[i] = qn; }
        else qn = *reinterpret_cast<const float4 *>(rotations + 4 * (size_t)i);
    }
    int rad_out = 0;
    uint32_t tiles_out = 0;
    uint2 rect_out = make_uint2(0u, 0u);
    if (in_range) do {
        float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
        const float *V = sVP, *P = sVP + 16;
        float3 pv = xform4x3(V, p);
        if (pv.z <= 0.2f) break;
        float hx = P[0] * p.x + P[4] * p.y + P[8] * p.z + P[12];
        float hy = P[1] * p.x + P[5] * p.y + P[9] * p.z + P[13];
        float hw = P[3] * p.x + P[7] * p.y + P[11] * p.z + P[15];
        float pw = 1.0f / (hw + 0.0000001f);
        float ndcx = hx * pw, ndcy = hy * pw;

        // Sigma = R diag((mod*s)^2) R^T
        float r = qn.x, x = qn.y, y = qn.z, z = qn.w;
        float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                         {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                         {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        float sx = cam.scale_modifier * scales[3 * i], sy = cam.scale_modifier * scales[3 * i + 1],
              sz = cam.scale_modifier * scales[3 * i + 2];
        float M[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            M[0][j] = sx * R[j][0];
            M[1][j] = sy * R[j][1];
            M[2][j] = sz * R[j][2];
        }
        float S[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) S[a][b] = M[0][a] * M[0][b] + M[1][a] * M[1][b] + M[2][a] * M[2][b];
        // force exact symmetry the way the 6-float storage of the reference does
        S[1][0] = S[0][1]; S[2][0] = S[0][2]; S[2][1] = S[1][2];

        // EWA projection
        float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
        float txtz = pv.x / pv.z, tytz = pv.y / pv.z;
        float tx = fminf(limx, fmaxf(-limx, txtz)) * pv.z;
        float ty = fminf(limy, fmaxf(-limy, tytz)) * pv.z;
        float J[2][3] = {{cam.focal_x / pv.z, 0.f, -(cam.focal_x * tx) / (pv.z * pv.z)},
                         {0.f, cam.focal_y / pv.z, -(cam.focal_y * ty) / (pv.z * pv.z)}};
        float Mm[2][3];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) Mm[a][b] = J[a][0] * V[b * 4 + 0] + J[a][1] * V[b * 4 + 1] + J[a][2] * V[b * 4 + 2];
        float MS[2][3];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) MS[a][b] = Mm[a][0] * S[0][b] + Mm[a][1] * S[1][b] + Mm[a][2] * S[2][b];
        float ca = MS[0][0] * Mm[0][0] + MS[0][1] * Mm[0][1] + MS[0][2] * Mm[0][2] + 0.3f;
        float cb = MS[0][0] * Mm[1][0] + MS[0][1] * Mm[1][1] + MS[0][2] * Mm[1][2];
        float cc = MS[1][0] * Mm[1][0] + MS[1][1] * Mm[1][1] + MS[1][2] * Mm[1][2] + 0.3f;
        float det = ca * cc - cb * cb;
        if (det == 0.0f) break;
        float det_inv = 1.f / det;
        float A = cc * det_inv, B = -cb * det_inv, C = ca * det_inv;
        float mid = 0.5f * (ca + cc);
        float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float rad = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
        float px = ((ndcx + 1.0f) * cam.W - 1.0f) * 0.5f;
        float py = ((ndcy + 1.0f) * cam.H - 1.0f) * 0.5f;
        int minx = min(cam.gx, max(0, (int)((px - rad) / GSD_TILE)));
        int miny = min(cam.gy, max(0, (int)((py - rad) / GSD_TILE)));
        int maxx = min(cam.gx, max(0, (int)((px + rad + GSD_TILE - 1) / GSD_TILE)));
        int maxy = min(cam.gy, max(0, (int)((py + rad + GSD_TILE - 1) / GSD_TILE)));
        if ((maxx - minx) * (maxy - miny) == 0) break;

        float o = opacities[i];
        // contribution bound used by the blend kernels' exact rectangle cull: a pixel at offset d contributes iff
        // d^T Q d <= 2 ln(255 o); stored with a safety margin (o < 1/255 gives a negative bound: never contributes)
        float ex = 2.0f * logf(255.0f * o);
        ex = ex * (ex > 0.f ? 1.0005f : 0.9995f) + 1e-3f;
        float ey = 0.f;
        depth[i] = pv.z;
        xy[i] = make_float2(px, py);
        conic_o[i] = make_float4(A, B, C, o);
        ext[i] = make_float2(ex, ey);
        rad_out = (int)rad;
        rect_out = make_uint2((uint32_t)minx | ((uint32_t)miny << 16), (uint32_t)maxx | ((uint32_t)maxy << 16));
        tiles_out = (uint32_t)((maxx - minx) * (maxy - miny));
    } while (0);
    // block-local exclusive scan of tiles touched: slot_base[i] = offset inside this block (the binning scatter adds the
    // block's base). No atomics: slots are deterministic and in Gaussian order.
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = tiles_out;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < wid) wbase += wsum[k];
        total += wsum[k];
    }
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
    if (in_range) {
        radii[i] = rad_out;
        tiles[i] = tiles_out;
        rect[i] = rect_out;
        slot_base[i] = wbase + incl - tiles_out;
    }
}

__global__ void gsd_mark_visible_kernel(int G, GsdCam cam, const float *__restrict__ means3D, uint8_t *__restrict__ vis) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    float Vm[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) Vm[k] = __ldg(cam.view + k);
    float3 pv = xform4x3(Vm, p);
    vis[i] = pv.z > 0.2f ? 1 : 0;
}

// block_sum[b] receives the number of tile instances of Gaussians [256 b, 256 b + 256)
int gsd_launch_preprocess(int G, const GsdCam &cam, const GsdRasterFwd *a, const GsdGeomWs &g, int32_t *zero_fill, int n_zero,
                          cudaStream_t st) {
    if (G == 0) return GSD_OK;
    int blocks = (G + 255) / 256;
    gsd_launch(gsd_preprocess_kernel, dim3(blocks), dim3(256), 0, st, G, cam, a->means3D, a->opacities, a->scales, a->rotations, g.xy,
                                                   g.conic_o, g.ext, g.depth, g.rect, g.tiles, g.slot_base, a->radii,
                                                   g.block_sum, zero_fill, n_zero, (const float4 *)a->unnorm_rotations,
                                                   (float4 *)const_cast<float *>(a->rotations));
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

int gsd_launch_mark_visible(int G, const GsdCam &cam, const float *means3D, uint8_t *vis, cudaStream_t st) {
    if (G == 0) return GSD_OK;
    gsd_launch(gsd_mark_visible_kernel, dim3((G + 255) / 256), dim3(256), 0, st, G, cam, means3D, vis);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
