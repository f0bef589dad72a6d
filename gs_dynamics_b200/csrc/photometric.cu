// photometric.cu — fused photometric loss of the tracking iteration:
//     loss = w_l1 * mean|x - y| + w_ssim * (1 - mean SSIM_11x11,sigma=1.5(x, y))
// and its gradient w.r.t. x, in two kernels (statistics + gradient) instead of the reference's 5 depthwise
// cuDNN convolutions forward + their backward per call, with the window rebuilt on the host every call
// (/root/reference/src/tracking/external.py:101-135 calc_ssim/_ssim, src/tracking/helpers.py:71-72 l1_loss_v1,
//  called at /root/reference/src/tracking/train_utils.py:185,195).
//
// Both kernels are separable 11-tap stencils over 32x32 tiles staged in shared memory (halo 5, zero padding like
// conv2d(padding=5)); HBM traffic per call: read x,y + write 3 partial maps (kernel 1), read 3 maps + x,y, write
// the gradient (kernel 2) = 40 B/pixel/channel.  Block partial sums are written to a buffer and reduced in fixed
// order (deterministic).
#include "common.cuh"

#define PH_T 32          // tile edge
#define PH_R 5           // window radius
#define PH_E (PH_T + 2 * PH_R)

struct PhWin { float g[11]; };

__device__ __forceinline__ float ph_load(const float *img, int W, int H, int x, int y) {
    return (x >= 0 && x < W && y >= 0 && y < H) ? img[(size_t)y * W + x] : 0.f;
}

// kernel 1: per pixel SSIM partials dS/dmu1, dS/ds11, dS/ds12 + block sums of |x-y| and SSIM
__global__ void __launch_bounds__(256)
gsd_ssim_stats_kernel(int C, int H, int W, PhWin win, const float *__restrict__ X, const float *__restrict__ Y,
                      float *__restrict__ dmu, float *__restrict__ ds11, float *__restrict__ ds12,
                      float *__restrict__ block_sums /* [nblocks][2] */) {
    __shared__ float sx[PH_E][PH_E + 1], sy[PH_E][PH_E + 1];
    __shared__ float h[5][PH_E][PH_T + 1]; // horizontally filtered x, y, xx, yy, xy
    __shared__ float red[2][8];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * PH_T, y0 = blockIdx.y * PH_T;
    const float *Xc = X + (size_t)c * H * W, *Yc = Y + (size_t)c * H * W;
    const int t = threadIdx.x;
    for (int i = t; i < PH_E * PH_E; i += 256) {
        int ly = i / PH_E, lx = i % PH_E;
        sx[ly][lx] = ph_load(Xc, W, H, x0 + lx - PH_R, y0 + ly - PH_R);
        sy[ly][lx] = ph_load(Yc, W, H, x0 + lx - PH_R, y0 + ly - PH_R);
    }
    __syncthreads();
    for (int i = t; i < PH_E * PH_T; i += 256) {
        int ly = i / PH_T, lx = i % PH_T;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float xv = sx[ly][lx + k], yv = sy[ly][lx + k], w = win.g[k];
            a += w * xv; b += w * yv; aa += w * xv * xv; bb += w * yv * yv; ab += w * xv * yv;
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = aa; h[3][ly][lx] = bb; h[4][ly][lx] = ab;
    }
    __syncthreads();
    float l1 = 0.f, ss = 0.f;
    for (int i = t; i < PH_T * PH_T; i += 256) {
        int ly = i / PH_T, lx = i % PH_T;
        int gx = x0 + lx, gy = y0 + ly;
        if (gx >= W || gy >= H) continue;
        float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float w = win.g[k];
            mu1 += w * h[0][ly + k][lx]; mu2 += w * h[1][ly + k][lx];
            s11 += w * h[2][ly + k][lx]; s22 += w * h[3][ly + k][lx]; s12 += w * h[4][ly + k][lx];
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        float mu1sq = mu1 * mu1, mu2sq = mu2 * mu2, mu12 = mu1 * mu2;
        float sig1 = s11 - mu1sq, sig2 = s22 - mu2sq, sig12 = s12 - mu12;
        float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2;
        float B1 = mu1sq + mu2sq + C1, B2 = sig1 + sig2 + C2;
        float inv = 1.f / (B1 * B2);
        float S = A1 * A2 * inv;
        // dS/dmu1 (through A1, A2, B1, B2), dS/ds11 (through B2), dS/ds12 (through A2)
        float dS_dA1 = A2 * inv, dS_dA2 = A1 * inv, dS_dB1 = -S / B1, dS_dB2 = -S / B2;
        float d_mu1 = dS_dA1 * 2.f * mu2 + dS_dA2 * (-2.f * mu2) + dS_dB1 * 2.f * mu1 + dS_dB2 * (-2.f * mu1);
        size_t pid = (size_t)c * H * W + (size_t)gy * W + gx;
        dmu[pid] = d_mu1;
        ds11[pid] = dS_dB2;
        ds12[pid] = 2.f * dS_dA2;
        ss += S;
        l1 += fabsf(sx[ly + PH_R][lx + PH_R] - sy[ly + PH_R][lx + PH_R]);
    }
    // block reduce (fixed order)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((t & 31) == 0) { red[0][t >> 5] = l1; red[1][t >> 5] = ss; }
    __syncthreads();
    if (t == 0) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < 8; ++k) { a += red[0][k]; b += red[1][k]; }
        size_t bid = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        block_sums[2 * bid] = a;
        block_sums[2 * bid + 1] = b;
    }
}

// fixed-order final reduction: out[0] = loss, out[1] = mean|x-y|, out[2] = mean SSIM
__global__ void gsd_ssim_finish_kernel(int nblocks, const float *__restrict__ block_sums, float inv_n, float w_l1,
                                       float w_ssim, float *__restrict__ out) {
    __shared__ double r0[256], r1[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { a += block_sums[2 * i]; b += block_sums[2 * i + 1]; }
    r0[threadIdx.x] = a; r1[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if (threadIdx.x < s) { r0[threadIdx.x] += r0[threadIdx.x + s]; r1[threadIdx.x] += r1[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float ml1 = (float)(r0[0] * inv_n), ms = (float)(r1[0] * inv_n);
        out[0] = w_l1 * ml1 + w_ssim * (1.0f - ms);
        out[1] = ml1;
        out[2] = ms;
    }
}

// kernel 2: grad_x = gscale * ( w_l1*sign(x-y)/N - w_ssim/N * (conv(dmu) + 2x*conv(ds11) + y*conv(ds12)) )
__global__ void __launch_bounds__(256)
gsd_ssim_grad_kernel(int C, int H, int W, PhWin win, const float *__restrict__ X, const float *__restrict__ Y,
                     const float *__restrict__ dmu, const float *__restrict__ ds11, const float *__restrict__ ds12,
                     const float *__restrict__ gscale_ptr, float gscale_mul, float w_l1, float w_ssim, float inv_n,
                     float *__restrict__ grad) {
    __shared__ float sm[3][PH_E][PH_E + 1];
    __shared__ float h[3][PH_E][PH_T + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * PH_T, y0 = blockIdx.y * PH_T;
    const size_t coff = (size_t)c * H * W;
    const int t = threadIdx.x;
    for (int i = t; i < PH_E * PH_E; i += 256) {
        int ly = i / PH_E, lx = i % PH_E;
        int gx = x0 + lx - PH_R, gy = y0 + ly - PH_R;
        sm[0][ly][lx] = ph_load(dmu + coff, W, H, gx, gy);
        sm[1][ly][lx] = ph_load(ds11 + coff, W, H, gx, gy);
        sm[2][ly][lx] = ph_load(ds12 + coff, W, H, gx, gy);
    }
    __syncthreads();
    for (int i = t; i < PH_E * PH_T; i += 256) {
        int ly = i / PH_T, lx = i % PH_T;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float w = win.g[k];
            a += w * sm[0][ly][lx + k]; b += w * sm[1][ly][lx + k]; d += w * sm[2][ly][lx + k];
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = d;
    }
    __syncthreads();
    const float gs = (gscale_ptr ? *gscale_ptr : 1.0f) * gscale_mul;
    for (int i = t; i < PH_T * PH_T; i += 256) {
        int ly = i / PH_T, lx = i % PH_T;
        int gx = x0 + lx, gy = y0 + ly;
        if (gx >= W || gy >= H) continue;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float w = win.g[k];
            a += w * h[0][ly + k][lx]; b += w * h[1][ly + k][lx]; d += w * h[2][ly + k][lx];
        }
        size_t pid = coff + (size_t)gy * W + gx;
        float xv = X[pid], yv = Y[pid];
        float df = xv - yv;
        float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
        grad[pid] = gs * inv_n * (w_l1 * sgn - w_ssim * (a + 2.f * xv * b + yv * d));
    }
}

static PhWin make_window() {
    PhWin w;
    // float32 arithmetic like torch.Tensor([...]) / sum  (external.py:56-70)
    float g[11], s = 0.f;
    for (int i = 0; i < 11; ++i) {
        double d = (double)(i - 5);
        g[i] = (float)exp(-(d * d) / (2.0 * 1.5 * 1.5));
        s += g[i];
    }
    for (int i = 0; i < 11; ++i) w.g[i] = g[i] / s;
    return w;
}

extern "C" int gsd_photometric_workspace_bytes(int32_t C, int32_t H, int32_t W, size_t *bytes) {
    if (C <= 0 || H <= 0 || W <= 0 || !bytes) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    size_t n = (size_t)C * H * W;
    size_t nb = (size_t)C * ((H + PH_T - 1) / PH_T) * ((W + PH_T - 1) / PH_T);
    *bytes = gsd_align_up(n * 4) * 3 + gsd_align_up(nb * 8);
    return GSD_OK;
}

// loss_out[3] = {loss, mean|x-y|, mean SSIM}; ws keeps the partial maps for the backward call
extern "C" int gsd_photometric_forward(int32_t C, int32_t H, int32_t W, const float *x, const float *y, float w_l1,
                                       float w_ssim, void *ws, float *loss_out, void *stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !x || !y || !ws || !loss_out) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)C * H * W;
    char *p = (char *)ws;
    float *dmu = (float *)p; p += gsd_align_up(n * 4);
    float *ds11 = (float *)p; p += gsd_align_up(n * 4);
    float *ds12 = (float *)p; p += gsd_align_up(n * 4);
    float *bs = (float *)p;
    dim3 grid((W + PH_T - 1) / PH_T, (H + PH_T - 1) / PH_T, C);
    PhWin win = make_window();
    gsd_ssim_stats_kernel<<<grid, 256, 0, st>>>(C, H, W, win, x, y, dmu, ds11, ds12, bs);
    GSD_LAUNCH_CHECK();
    int nb = (int)(grid.x * grid.y * grid.z);
    gsd_ssim_finish_kernel<<<1, 256, 0, st>>>(nb, bs, 1.0f / (float)n, w_l1, w_ssim, loss_out);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// grad_x = (*gscale_ptr or 1) * gscale_mul * dloss/dx
extern "C" int gsd_photometric_backward(int32_t C, int32_t H, int32_t W, const float *x, const float *y, float w_l1,
                                        float w_ssim, const void *ws, const float *gscale_ptr, float gscale_mul,
                                        float *grad_x, void *stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !x || !y || !ws || !grad_x) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)C * H * W;
    const char *p = (const char *)ws;
    const float *dmu = (const float *)p; p += gsd_align_up(n * 4);
    const float *ds11 = (const float *)p; p += gsd_align_up(n * 4);
    const float *ds12 = (const float *)p;
    dim3 grid((W + PH_T - 1) / PH_T, (H + PH_T - 1) / PH_T, C);
    PhWin win = make_window();
    gsd_ssim_grad_kernel<<<grid, 256, 0, st>>>(C, H, W, win, x, y, dmu, ds11, ds12, gscale_ptr, gscale_mul, w_l1, w_ssim,
                                               1.0f / (float)n, grad_x);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
