// photometric.cu — fused photometric loss of the tracking iteration:
//     loss = w_l1 * mean|x - y| + w_ssim * (1 - mean SSIM_11x11,sigma=1.5(x, y)),   x = scale_c * rendered + shift_c
// and its gradient w.r.t. the rendered image, in two kernels (statistics + gradient) instead of the reference's 5 depthwise
// cuDNN convolutions forward + their backward per call, with the window rebuilt on the host every call
// (/root/reference/src/tracking/external.py:101-135 calc_ssim/_ssim, src/tracking/helpers.py:71-72 l1_loss_v1,
//  called at /root/reference/src/tracking/train_utils.py:182-185,195; the affine is the cam_m / cam_c correction of :182).
// Channels are grouped in sets of 3 (RGB render, seg render) with their own loss weight so that both losses of an
// iteration are one launch.
//
// Both kernels are separable 11-tap stencils over 32x32 tiles staged in shared memory (halo 5, zero padding like
// conv2d(padding=5)); each thread produces 4 consecutive outputs per pass from a 14-value register window (8x fewer
// shared-memory reads than one output per thread).  HBM traffic per call: read x,y + write 3 partial maps (kernel 1),
// read 3 maps + x,y, write the gradient (kernel 2) = 40 B/pixel/channel.  Block partial sums are written to a buffer and
// reduced in fixed order (deterministic).
#include "common.cuh"

#define PH_T 32          // tile edge
#define PH_R 5           // window radius
#define PH_E (PH_T + 2 * PH_R)
#define PH_MAXC 6

struct PhWin { float g[11]; };
struct PhAffine { float scale[PH_MAXC], shift[PH_MAXC]; int on; };

__device__ __forceinline__ float ph_load(const float *img, int W, int H, int x, int y) {
    return (x >= 0 && x < W && y >= 0 && y < H) ? img[(size_t)y * W + x] : 0.f;
}

// affine parameters may live on the device (cam_m / cam_c rows): scale = exp(m[c]), shift = c[c]
__device__ __forceinline__ void ph_affine(const float *log_scale, const float *shift, int c, float &s, float &b) {
    s = log_scale ? __expf(log_scale[c % 3]) : 1.0f;
    b = shift ? shift[c % 3] : 0.0f;
}

// kernel 1: per pixel SSIM partials dS/dmu1, dS/ds11, dS/ds12 + block sums of |x-y| and SSIM.
// MODE 0: everything from x and y.  MODE 1: the target's window statistics (mu2 = conv(y), s22 = conv(y*y)) are read from
// maps precomputed once per target image (they do not change during an episode frame).  MODE 2: write those maps.
template <int MODE>
__global__ void __launch_bounds__(256, 4)
gsd_ssim_stats_kernel(int C, int H, int W, PhWin win, const float *__restrict__ X, const float *__restrict__ Y,
                      const float *__restrict__ log_scale, const float *__restrict__ shift, int affine_channels,
                      float *__restrict__ y_mu, float *__restrict__ y_s22,
                      float *__restrict__ dmu, float *__restrict__ ds11, float *__restrict__ ds12,
                      float *__restrict__ block_sums /* [nblocks][2] */) {
    constexpr int NQ = (MODE == 0) ? 5 : (MODE == 1 ? 3 : 2); // x, xx, xy (, y, yy)   |   MODE 2: y, yy
    __shared__ float sx[PH_E][PH_E + 1], sy[PH_E][PH_E + 1];
    __shared__ float h[NQ][PH_E][PH_T + 1];
    __shared__ float red[2][8];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * PH_T, y0 = blockIdx.y * PH_T;
    const float *Xc = X + (size_t)c * H * W, *Yc = Y + (size_t)c * H * W;
    const int t = threadIdx.x;
    float as = 1.f, ab = 0.f;
    if (MODE != 2 && c < affine_channels) ph_affine(log_scale, shift, c, as, ab);
    for (int i = t; i < PH_E * PH_E; i += 256) {
        int ly = i / PH_E, lx = i % PH_E;
        const int gx = x0 + lx - PH_R, gy = y0 + ly - PH_R;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        if (MODE != 2) sx[ly][lx] = in ? as * Xc[(size_t)gy * W + gx] + ab : 0.f;
        sy[ly][lx] = in ? Yc[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    // horizontal pass: unit = (row, segment of 4 outputs); consecutive lanes take consecutive rows (pitch 43: no conflicts)
    for (int u = t; u < PH_E * (PH_T / 4); u += 256) {
        const int ly = u % PH_E, seg = u / PH_E;
        float q[NQ][14];
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const float yy = sy[ly][seg * 4 + k];
            if (MODE == 2) {
                q[0][k] = yy; q[1][k] = yy * yy;
            } else {
                const float xx = sx[ly][seg * 4 + k];
                q[0][k] = xx; q[1][k] = xx * xx; q[2][k] = xx * yy;
                if (MODE == 0) { q[3 % NQ][k] = yy; q[4 % NQ][k] = yy * yy; }
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
#pragma unroll
            for (int n = 0; n < NQ; ++n) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) a += win.g[k] * q[n][o + k];
                h[n][ly][seg * 4 + o] = a;
            }
        }
    }
    __syncthreads();
    // vertical pass: thread = (column, segment of 4 rows)
    float l1 = 0.f, ss = 0.f;
    {
        const int lx = t % PH_T, seg = t / PH_T; // 32 columns x 8 segments
        float v[NQ][14];
#pragma unroll
        for (int n = 0; n < NQ; ++n)
#pragma unroll
            for (int k = 0; k < 14; ++k) v[n][k] = h[n][seg * 4 + k][lx];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int ly = seg * 4 + o;
            const int gx = x0 + lx, gy = y0 + ly;
            float r[NQ];
#pragma unroll
            for (int n = 0; n < NQ; ++n) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) a += win.g[k] * v[n][o + k];
                r[n] = a;
            }
            if (gx < W && gy < H) {
                const size_t pid = (size_t)c * H * W + (size_t)gy * W + gx;
                if (MODE == 2) {
                    y_mu[pid] = r[0];
                    y_s22[pid] = r[1];
                } else {
                    const float mu1 = r[0], s11 = r[1], s12 = r[2];
                    const float mu2 = (MODE == 0) ? r[3 % NQ] : y_mu[pid];
                    const float s22 = (MODE == 0) ? r[4 % NQ] : y_s22[pid];
                    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
                    const float mu1sq = mu1 * mu1, mu2sq = mu2 * mu2, mu12 = mu1 * mu2;
                    const float sig1 = s11 - mu1sq, sig2 = s22 - mu2sq, sig12 = s12 - mu12;
                    const float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2;
                    const float B1 = mu1sq + mu2sq + C1, B2 = sig1 + sig2 + C2;
                    const float inv = 1.f / (B1 * B2);
                    const float S = A1 * A2 * inv;
                    const float dS_dA1 = A2 * inv, dS_dA2 = A1 * inv, dS_dB1 = -S / B1, dS_dB2 = -S / B2;
                    dmu[pid] = dS_dA1 * 2.f * mu2 + dS_dA2 * (-2.f * mu2) + dS_dB1 * 2.f * mu1 + dS_dB2 * (-2.f * mu1);
                    ds11[pid] = dS_dB2;
                    ds12[pid] = 2.f * dS_dA2;
                    ss += S;
                    l1 += fabsf(sx[ly + PH_R][lx + PH_R] - sy[ly + PH_R][lx + PH_R]);
                }
            }
        }
    }
    if (MODE != 2) {
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
}

// fixed-order final reduction per set of 3 channels: out[3*s + 0] = loss_s, [1] = mean|x-y|, [2] = mean SSIM;
// out[3*n_sets] = sum_s set_weight[s] * loss_s
__global__ void gsd_ssim_finish_kernel(int n_sets, int blocks_per_set, const float *__restrict__ block_sums, float inv_n,
                                       float w_l1, float w_ssim, float sw0, float sw1, float *__restrict__ out) {
    __shared__ double r0[256], r1[256];
    float total = 0.f;
    for (int s = 0; s < n_sets; ++s) {
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < blocks_per_set; i += 256) {
            a += block_sums[2 * ((size_t)s * blocks_per_set + i)];
            b += block_sums[2 * ((size_t)s * blocks_per_set + i) + 1];
        }
        r0[threadIdx.x] = a; r1[threadIdx.x] = b;
        __syncthreads();
        for (int k = 128; k >= 1; k >>= 1) {
            if (threadIdx.x < k) { r0[threadIdx.x] += r0[threadIdx.x + k]; r1[threadIdx.x] += r1[threadIdx.x + k]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            float ml1 = (float)(r0[0] * inv_n), ms = (float)(r1[0] * inv_n);
            float l = w_l1 * ml1 + w_ssim * (1.0f - ms);
            out[3 * s] = l; out[3 * s + 1] = ml1; out[3 * s + 2] = ms;
            total += (s == 0 ? sw0 : sw1) * l;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[3 * n_sets] = total;
}

// kernel 2: d loss / d rendered = gscale * set_weight * scale_c * ( w_l1*sign(x-y)/N - w_ssim/N * (conv(dmu) + 2x*conv(ds11) + y*conv(ds12)) )
__global__ void __launch_bounds__(256, 4)
gsd_ssim_grad_kernel(int C, int H, int W, PhWin win, const float *__restrict__ X, const float *__restrict__ Y,
                     const float *__restrict__ log_scale, const float *__restrict__ shift, int affine_channels,
                     const float *__restrict__ dmu, const float *__restrict__ ds11, const float *__restrict__ ds12,
                     const float *__restrict__ gscale_ptr, float sw0, float sw1, float w_l1, float w_ssim, float inv_n,
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
    for (int u = t; u < PH_E * (PH_T / 4); u += 256) {
        const int ly = u % PH_E, seg = u / PH_E;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float v[14];
#pragma unroll
            for (int k = 0; k < 14; ++k) v[k] = sm[q][ly][seg * 4 + k];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) a += win.g[k] * v[o + k];
                h[q][ly][seg * 4 + o] = a;
            }
        }
    }
    __syncthreads();
    float as = 1.f, ab = 0.f;
    if (c < affine_channels) ph_affine(log_scale, shift, c, as, ab);
    const float gs = (gscale_ptr ? *gscale_ptr : 1.0f) * (c < 3 ? sw0 : sw1) * as;
    const int lx = t % PH_T, seg = t / PH_T;
    float v[3][14];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int k = 0; k < 14; ++k) v[q][k] = h[q][seg * 4 + k][lx];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int ly = seg * 4 + o;
        const int gx = x0 + lx, gy = y0 + ly;
        if (gx >= W || gy >= H) continue;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = win.g[k];
            a += w * v[0][o + k]; b += w * v[1][o + k]; d += w * v[2][o + k];
        }
        const size_t pid = coff + (size_t)gy * W + gx;
        const float xv = as * X[pid] + ab, yv = Y[pid];
        const float df = xv - yv;
        const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
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

static int ph_check(const GsdPhotometric *p) {
    if (!p || p->H <= 0 || p->W <= 0 || (p->n_sets != 1 && p->n_sets != 2) || p->C != (p->n_sets == 2 ? 6 : p->C) || p->C <= 0 ||
        p->C > PH_MAXC || !p->x || !p->y || !p->ws) {
        gsd_set_error("invalid photometric descriptor");
        return GSD_ERR_INVALID;
    }
    return GSD_OK;
}

// loss_out: per set {loss, mean|x-y|, mean SSIM}, then the weighted total  (3*n_sets + 1 floats)
extern "C" int gsd_photometric_forward(const GsdPhotometric *p, float *loss_out, void *stream) {
    int rc;
    if ((rc = ph_check(p))) return rc;
    if (!loss_out) { gsd_set_error("null loss_out"); return GSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int C = p->C, H = p->H, W = p->W;
    size_t n = (size_t)C * H * W;
    char *q = (char *)p->ws;
    float *dmu = (float *)q; q += gsd_align_up(n * 4);
    float *ds11 = (float *)q; q += gsd_align_up(n * 4);
    float *ds12 = (float *)q; q += gsd_align_up(n * 4);
    float *bs = (float *)q;
    dim3 grid((W + PH_T - 1) / PH_T, (H + PH_T - 1) / PH_T, C);
    PhWin win = make_window();
    const int aff = (p->affine_log_scale || p->affine_shift) ? 3 : 0;
    if (p->y_mu && p->y_s22)
        gsd_ssim_stats_kernel<1><<<grid, 256, 0, st>>>(C, H, W, win, p->x, p->y, p->affine_log_scale, p->affine_shift, aff,
                                                        (float *)p->y_mu, (float *)p->y_s22, dmu, ds11, ds12, bs);
    else
        gsd_ssim_stats_kernel<0><<<grid, 256, 0, st>>>(C, H, W, win, p->x, p->y, p->affine_log_scale, p->affine_shift, aff,
                                                        nullptr, nullptr, dmu, ds11, ds12, bs);
    GSD_LAUNCH_CHECK();
    const int per_set_c = C / p->n_sets;
    const int blocks_per_set = (int)(grid.x * grid.y) * per_set_c;
    gsd_ssim_finish_kernel<<<1, 256, 0, st>>>(p->n_sets, blocks_per_set, bs, 1.0f / (float)((size_t)per_set_c * H * W), p->w_l1,
                                              p->w_ssim, p->set_weight[0], p->set_weight[1], loss_out);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// window statistics of a target image, computed once and passed as GsdPhotometric.y_mu / y_s22 on later calls
extern "C" int gsd_photometric_target_stats(int32_t C, int32_t H, int32_t W, const float *y, float *y_mu, float *y_s22, void *stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !y || !y_mu || !y_s22) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    dim3 grid((W + PH_T - 1) / PH_T, (H + PH_T - 1) / PH_T, C);
    PhWin win = make_window();
    gsd_ssim_stats_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(C, H, W, win, y, y, nullptr, nullptr, 0, y_mu, y_s22, nullptr,
                                                                     nullptr, nullptr, nullptr);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// grad = (gscale_ptr ? *gscale_ptr : 1) * set_weight[set] * d loss_set / d rendered
extern "C" int gsd_photometric_backward(const GsdPhotometric *p, const float *gscale_ptr, float *grad, void *stream) {
    int rc;
    if ((rc = ph_check(p))) return rc;
    if (!grad) { gsd_set_error("null grad"); return GSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int C = p->C, H = p->H, W = p->W;
    size_t n = (size_t)C * H * W;
    const char *q = (const char *)p->ws;
    const float *dmu = (const float *)q; q += gsd_align_up(n * 4);
    const float *ds11 = (const float *)q; q += gsd_align_up(n * 4);
    const float *ds12 = (const float *)q;
    dim3 grid((W + PH_T - 1) / PH_T, (H + PH_T - 1) / PH_T, C);
    PhWin win = make_window();
    const int per_set_c = C / p->n_sets;
    gsd_ssim_grad_kernel<<<grid, 256, 0, st>>>(C, H, W, win, p->x, p->y, p->affine_log_scale, p->affine_shift,
                                               (p->affine_log_scale || p->affine_shift) ? 3 : 0, dmu, ds11, ds12, gscale_ptr,
                                               p->set_weight[0], p->set_weight[1], p->w_l1, p->w_ssim,
                                               1.0f / (float)((size_t)per_set_c * H * W), grad);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
