// photometric.cu — fused photometric loss of the tracking iteration:
//     loss = w_l1 * mean|x - y| + w_ssim * (1 - mean SSIM_11x11,sigma=1.5(x, y)),   x = scale_c * rendered + shift_c
// and its gradient w.r.t. the rendered image, in two kernels (statistics + gradient) instead of the reference's 5 depthwise
// cuDNN convolutions forward + their backward per call, with the window rebuilt on the host every call
// (/root/reference/src/tracking/external.py:101-135 calc_ssim/_ssim, src/tracking/helpers.py:71-72 l1_loss_v1,
//  called at /root/reference/src/tracking/train_utils.py:182-185,195; the affine is the cam_m / cam_c correction of :182).
// Channels are grouped in sets of 3 (RGB render, seg render) with their own loss weight so that both losses of an
// iteration are one launch.
//
// Both kernels are separable 11-tap stencils written as ROW-STREAMING warps: a warp owns a strip of 32*NCOL output columns
// and a segment of `seg_rows` output rows of one channel and marches down the rows.  Rows are fetched PH_D-1 rows ahead by
// asynchronous 4-byte copies (LDGSTS, zero-filled outside the image like conv2d(padding=5)) into a per-warp shared-memory
// ring (one commit group per row, one __syncwarp per row, no block barrier), filtered horizontally from a register window and
// scattered into a 10-slot ring of vertical accumulators held in registers (the ring shift is folded into the
// destination registers of the update FMAs, so the row loop is compact: no unrolling, no moves).  There is no vertical halo recompute inside a segment (10 rows at its ends only), the
// work unit is one warp (fine-grained tail), and global latency is hidden by the depth of the ring instead of by
// co-resident blocks.  HBM traffic per call: read x,y + write 3 partial maps (kernel 1), read 3 maps + x,y, write the
// gradient (kernel 2) = 40 B/pixel/channel.  Per-warp partial sums are written to a buffer and reduced in fixed order.
#include "common.cuh"

#define PH_R 5           // window radius
#define PH_MAXC 6
#define PH_WARPS 1       // warps per CTA (adjacent strips of one channel / segment)
#define PH_MIN_SEG 8     // smallest segment the launcher picks (bounds the partial-sum buffer)

struct PhWin { float g[11]; };

// affine parameters may live on the device (cam_m / cam_c rows): scale = exp(m[c]), shift = c[c]
__device__ __forceinline__ void ph_affine(const float *log_scale, const float *shift, int c, float &s, float &b) {
    s = log_scale ? __expf(log_scale[c % 3]) : 1.0f;
    b = shift ? shift[c % 3] : 0.0f;
}

template <int NCOL> struct PhGeom {
    static constexpr int OUT = 32 * NCOL;            // output columns per strip
    static constexpr int LC = OUT + 2 * PH_R;        // loaded columns
    static constexpr int NL = (LC + 31) / 32;        // loads per lane per row
    static constexpr int PITCH = 32 * NL + 4;        // staged line (floats)
    static constexpr int WN = 10 + NCOL;             // register window
};

#define PH_D 4           // depth of the per-warp row ring (rows in flight: PH_D - 1)

// 4-byte asynchronous global -> shared copy, zero-filled when !pred (LDGSTS; the address is not dereferenced then)
__device__ __forceinline__ void ph_cp4(unsigned dst_saddr, const float *src, bool pred) {
    const int sz = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_saddr), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ph_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ph_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-warp streaming state shared by both kernels: NA staged input arrays (LC columns from column xl, rows [rs, re)) and NB
// per-output arrays (the lane's own NCOL pixels of the output row completed PH_R rows later) ride in one commit group per
// streamed row.  Pointers advance by W per row; column predicates are per-lane constants.
template <int NCOL, int NA, int NB>
struct PhStream {
    using Gm = PhGeom<NCOL>;
    static constexpr int SLOT = NA * Gm::PITCH + NB * 32 * NCOL;   // floats per ring slot
    const float *in[NA];      // next row to issue, at column xl + lane
    const float *out[NB ? NB : 1];  // output row of the next row to issue, at column x0 + NCOL*lane
    const float *safe;        // valid address for zero-filled (not dereferenced) copies
    unsigned base;            // shared address of the warp's ring
    int row, W, H, r0;
    unsigned colmask;         // bit i: column xl + i*32 + lane inside the image and the strip; bit 8+j: output column j inside
    __device__ __forceinline__ void init(float *ring, const float *const *src, const float *const *osrc, int W_, int H_, int rs,
                                         int r0_, int xl, int x0, int lane) {
        base = (unsigned)__cvta_generic_to_shared(ring);
        safe = src[0];
        row = rs; W = W_; H = H_; r0 = r0_;
        colmask = 0;
#pragma unroll
        for (int i = 0; i < Gm::NL; ++i) {
            const int j = i * 32 + lane, gx = xl + j;
            if (j < Gm::LC && gx >= 0 && gx < W) colmask |= 1u << i;
        }
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (x0 + NCOL * lane + j < W) colmask |= 1u << (8 + j);
#pragma unroll
        for (int a = 0; a < NA; ++a) in[a] = src[a] + (ptrdiff_t)rs * W + xl + lane;
#pragma unroll
        for (int a = 0; a < NB; ++a) out[a] = osrc[a] + (ptrdiff_t)(rs - PH_R) * W + x0 + NCOL * lane;
    }
    __device__ __forceinline__ void issue(int lane) {
        const unsigned slot = base + (unsigned)(row & (PH_D - 1)) * (SLOT * 4u) + lane * 4u;
        const bool rin = (unsigned)row < (unsigned)H;
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int i = 0; i < Gm::NL; ++i) {
                const bool ok = rin && ((colmask >> i) & 1u);
                ph_cp4(slot + (a * Gm::PITCH + i * 32) * 4u, ok ? in[a] + i * 32 : safe, ok);
            }
        const bool oin = row - PH_R >= r0 && row - PH_R < H;
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                const bool ok = oin && ((colmask >> (8 + j)) & 1u);
                ph_cp4(slot + (NA * Gm::PITCH + a * 32 * NCOL + (NCOL - 1) * lane + j) * 4u, ok ? out[a] + j : safe, ok);
            }
        ph_commit();
#pragma unroll
        for (int a = 0; a < NA; ++a) in[a] += W;
#pragma unroll
        for (int a = 0; a < NB; ++a) out[a] += W;
        ++row;
    }
};

template <int NCOL>
__device__ __forceinline__ void ph_window(float (&w)[PhGeom<NCOL>::WN], const float *line, int lane) {
    if (NCOL == 1) {
#pragma unroll
        for (int k = 0; k < 11; ++k) w[k] = line[lane + k];
    } else {
#pragma unroll
        for (int k = 0; k < PhGeom<NCOL>::WN / 2; ++k) {
            const float2 t = *reinterpret_cast<const float2 *>(line + 2 * lane + 2 * k);
            w[2 * k] = t.x; w[2 * k + 1] = t.y;
        }
    }
}

// kernel 1: per pixel SSIM partials dS/dmu1, dS/ds11, dS/ds12 + per-warp sums of |x-y| and SSIM.
// MODE 0: everything from x and y.  MODE 1: the target's window statistics (mu2 = conv(y), s22 = conv(y*y)) are read from
// maps precomputed once per target image (they do not change during an episode frame).  MODE 2: write those maps.
template <int MODE, int NCOL>
__global__ void __launch_bounds__(32 * PH_WARPS)
gsd_ssim_stats_kernel(int C, int H, int W, int seg_rows, int n_seg, int n_strip, PhWin win, const float *__restrict__ X,
                      const float *__restrict__ Y, const float *__restrict__ log_scale, const float *__restrict__ shift,
                      int affine_channels, float *__restrict__ y_mu, float *__restrict__ y_s22, float *__restrict__ dmu,
                      float *__restrict__ ds11, float *__restrict__ ds12, float *__restrict__ unit_sums /* [units][2] */) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    using Gm = PhGeom<NCOL>;
    constexpr int NQ = (MODE == 0) ? 5 : (MODE == 1 ? 3 : 2); // x, xx, xy (, y, yy)   |   MODE 2: y, yy
    constexpr int NA = (MODE == 2) ? 1 : 2;                    // staged arrays: (x,) y
    constexpr int NB = (MODE == 1) ? 2 : 0;                    // per-output arrays: y_mu, y_s22 (MODE 1 only)
    using St = PhStream<NCOL, NA, NB>;
    __shared__ __align__(16) float ring_s[PH_WARPS][PH_D][St::SLOT];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int unit = blockIdx.x * PH_WARPS + wid;
    if (unit >= C * n_seg * n_strip) return;
    const int sx = unit % n_strip, sy = (unit / n_strip) % n_seg, c = unit / (n_strip * n_seg);
    const int x0 = sx * Gm::OUT, xl = x0 - PH_R;
    const int r0 = sy * seg_rows, r1 = min(r0 + seg_rows, H);
    const size_t coff = (size_t)c * H * W;
    const float *src[NA];
    src[0] = (MODE == 2 ? Y : X) + coff;
    src[NA - 1] = Y + coff;
    const float *osrc[2] = {MODE == 1 ? y_mu + coff : nullptr, MODE == 1 ? y_s22 + coff : nullptr};
    float as = 1.f, ab = 0.f;
    if (MODE != 2 && c < affine_channels) ph_affine(log_scale, shift, c, as, ab);
    const bool affine = (as != 1.f) || (ab != 0.f);

    // vertical ring: acc[.][.][i] = partial sum of output row (current input row - 4 + i); the shift is folded into the
    // destination register of the update FMAs, so the row loop needs no unrolling and no moves
    float acc[NQ][NCOL][10];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
#pragma unroll
            for (int s = 0; s < 10; ++s) acc[q][j][s] = 0.f;
    float l1 = 0.f, ss = 0.f;
    float (*ring)[St::SLOT] = ring_s[wid];
    const int rs = r0 - PH_R, re = r1 + PH_R;  // rows streamed: [rs, re)
    St stm;
    stm.init(&ring[0][0], src, osrc, W, H, rs, r0, xl, x0, lane);

    // a landed row: affine in place on the elements this lane copied (zero padding stays zero), L1 term on owned pixels
    auto finish_row = [&](int row) {
        if (MODE == 2) return;
        const bool own_row = row >= r0 && row < r1;
        if (!own_row && !affine) return;
        float *sl = ring[row & (PH_D - 1)];
        const bool rin = (unsigned)row < (unsigned)H;
#pragma unroll
        for (int i = 0; i < Gm::NL; ++i) {
            const int j = i * 32 + lane;
            if (rin && ((stm.colmask >> i) & 1u)) {
                float xv = sl[j];
                if (affine) { xv = as * xv + ab; sl[j] = xv; }
                if (own_row && j >= PH_R && j < PH_R + Gm::OUT) l1 += fabsf(xv - sl[Gm::PITCH + j]);
            }
        }
    };

#pragma unroll
    for (int k = 0; k < PH_D - 1; ++k) stm.issue(lane);
    ph_wait<PH_D - 2>();
    finish_row(rs);
#pragma unroll 1
    for (int rr = rs; rr < re; ++rr) {
        __syncwarp();                       // row rr (finished last iteration) visible; slot of row rr-1 free
        stm.issue(lane);                    // row rr + PH_D - 1
        const float *sl = ring[rr & (PH_D - 1)];
        // rows outside the image are staged as zeros: the same arithmetic handles them
        float wx[Gm::WN], wy[Gm::WN];
        if (MODE != 2) ph_window<NCOL>(wx, sl, lane);
        ph_window<NCOL>(wy, sl + (NA - 1) * Gm::PITCH, lane);
        float done[NQ][NCOL];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            float hq[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) hq[q] = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float g = win.g[k], yy = wy[j + k];
                if (MODE == 2) {
                    hq[0] += g * yy; hq[1] += g * (yy * yy);
                } else {
                    const float xx = wx[j + k];
                    hq[0] += g * xx; hq[1] += g * (xx * xx); hq[2] += g * (xx * yy);
                    if (MODE == 0) { hq[3 % NQ] += g * yy; hq[4 % NQ] += g * (yy * yy); }
                }
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                done[q][j] = acc[q][j][0] + win.g[0] * hq[q];
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[q][j][i] = acc[q][j][i + 1] + win.g[i + 1] * hq[q];
                acc[q][j][9] = win.g[10] * hq[q];
            }
        }
        const int o = rr - PH_R;            // output row completed by this input row
        if (o >= r0) {
            float *gp0 = (MODE == 2 ? y_mu : dmu) + coff + (size_t)o * W + x0 + NCOL * lane;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                if ((stm.colmask >> (8 + j)) & 1u) {
                    if (MODE == 2) {
                        gp0[j] = done[0][j];
                        (y_s22 + (gp0 - y_mu))[j] = done[1][j];
                    } else {
                        const float mu1 = done[0][j], s11 = done[1][j], s12 = done[2][j];
                        const float mu2 = (MODE == 0) ? done[3 % NQ][j] : sl[NA * Gm::PITCH + NCOL * lane + j];
                        const float s22v = (MODE == 0) ? done[4 % NQ][j] : sl[NA * Gm::PITCH + 32 * NCOL + NCOL * lane + j];
                        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
                        const float mu1sq = mu1 * mu1, mu2sq = mu2 * mu2, mu12 = mu1 * mu2;
                        const float sig1 = s11 - mu1sq, sig2 = s22v - mu2sq, sig12 = s12 - mu12;
                        const float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2;
                        const float B1 = mu1sq + mu2sq + C1, B2 = sig1 + sig2 + C2;   // both >= C1, C2 > 0
                        const float i1 = __fdividef(1.f, B1), i2 = __fdividef(1.f, B2);
                        const float inv = i1 * i2;
                        const float S = A1 * A2 * inv;
                        const float dS_dA1 = A2 * inv, dS_dA2 = A1 * inv, dS_dB1 = -S * i1, dS_dB2 = -S * i2;
                        gp0[j] = 2.f * (mu2 * (dS_dA1 - dS_dA2) + mu1 * (dS_dB1 - dS_dB2));
                        (ds11 + (gp0 - dmu))[j] = dS_dB2;
                        (ds12 + (gp0 - dmu))[j] = 2.f * dS_dA2;
                        ss += S;
                    }
                }
            }
        }
        ph_wait<PH_D - 2>();                // row rr + 1 has landed (this lane's copies)
        finish_row(rr + 1);
    }
    ph_wait<0>();
    if (MODE != 2) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            l1 += __shfl_xor_sync(0xffffffffu, l1, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) { unit_sums[2 * (size_t)unit] = l1; unit_sums[2 * (size_t)unit + 1] = ss; }
    }
}

// fixed-order final reduction per set of 3 channels: out[3*s + 0] = loss_s, [1] = mean|x-y|, [2] = mean SSIM;
// out[3*n_sets] = sum_s set_weight[s] * loss_s
__global__ void __launch_bounds__(64)
gsd_ssim_finish_kernel(int n_sets, int blocks_per_set, const float *__restrict__ block_sums, float inv_n,
                       float w_l1, float w_ssim, float sw0, float sw1, const float *__restrict__ add, float *__restrict__ out) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    // one warp per set: lane-strided double sums, fixed butterfly -> deterministic
    __shared__ float set_loss[2];
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (s < n_sets) {
        double a = 0.0, b = 0.0;
        const float2 *bs = reinterpret_cast<const float2 *>(block_sums) + (size_t)s * blocks_per_set;
#pragma unroll 4
        for (int i = lane; i < blocks_per_set; i += 32) {
            const float2 v = bs[i];
            a += v.x; b += v.y;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            const float ml1 = (float)(a * inv_n), ms = (float)(b * inv_n);
            const float l = w_l1 * ml1 + w_ssim * (1.0f - ms);
            out[3 * s] = l; out[3 * s + 1] = ml1; out[3 * s + 2] = ms;
            set_loss[s] = l;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float total = sw0 * set_loss[0] + (n_sets > 1 ? sw1 * set_loss[1] : 0.f);
        out[3 * n_sets] = total;
        if (add) out[3 * n_sets + 1] = total + *add;
    }
}

// kernel 2: d loss / d rendered = gscale * set_weight * scale_c * ( w_l1*sign(x-y)/N - w_ssim/N * (conv(dmu) + 2x*conv(ds11) + y*conv(ds12)) )
// (same row-streaming structure; three staged maps, x and y of the completed output row ride in the same commit group)
template <int NCOL>
__global__ void __launch_bounds__(32 * PH_WARPS)
gsd_ssim_grad_kernel(int C, int H, int W, int seg_rows, int n_seg, int n_strip, PhWin win, const float *__restrict__ X,
                     const float *__restrict__ Y, const float *__restrict__ log_scale, const float *__restrict__ shift,
                     int affine_channels, const float *__restrict__ dmu, const float *__restrict__ ds11,
                     const float *__restrict__ ds12, const float *__restrict__ gscale_ptr, float sw0, float sw1, float w_l1,
                     float w_ssim, float inv_n, float *__restrict__ grad) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    using Gm = PhGeom<NCOL>;
    using St = PhStream<NCOL, 3, 2>;
    __shared__ __align__(16) float ring_s[PH_WARPS][PH_D][St::SLOT];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int unit = blockIdx.x * PH_WARPS + wid;
    if (unit >= C * n_seg * n_strip) return;
    const int sx = unit % n_strip, sy = (unit / n_strip) % n_seg, c = unit / (n_strip * n_seg);
    const int x0 = sx * Gm::OUT, xl = x0 - PH_R;
    const int r0 = sy * seg_rows, r1 = min(r0 + seg_rows, H);
    const size_t coff = (size_t)c * H * W;
    const float *src[3] = {dmu + coff, ds11 + coff, ds12 + coff};
    const float *osrc[2] = {X + coff, Y + coff};
    float as = 1.f, ab = 0.f;
    if (c < affine_channels) ph_affine(log_scale, shift, c, as, ab);
    const float gs = (gscale_ptr ? *gscale_ptr : 1.0f) * (c < 3 ? sw0 : sw1) * as * inv_n;

    float acc[3][NCOL][10];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
#pragma unroll
            for (int s = 0; s < 10; ++s) acc[q][j][s] = 0.f;
    float (*ring)[St::SLOT] = ring_s[wid];
    const int rs = r0 - PH_R, re = r1 + PH_R;
    St stm;
    stm.init(&ring[0][0], src, osrc, W, H, rs, r0, xl, x0, lane);
#pragma unroll
    for (int k = 0; k < PH_D - 1; ++k) stm.issue(lane);
#pragma unroll 1
    for (int rr = rs; rr < re; ++rr) {
        ph_wait<PH_D - 2>();                // row rr has landed (this lane's copies) ...
        __syncwarp();                       // ... and everybody's; slot of row rr-1 free
        stm.issue(lane);
        const float *sl = ring[rr & (PH_D - 1)];
        float done[3][NCOL];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float w[Gm::WN];
            ph_window<NCOL>(w, sl + q * Gm::PITCH, lane);
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                float h = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) h += win.g[k] * w[j + k];
                done[q][j] = acc[q][j][0] + win.g[0] * h;
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[q][j][i] = acc[q][j][i + 1] + win.g[i + 1] * h;
                acc[q][j][9] = win.g[10] * h;
            }
        }
        const int o = rr - PH_R;
        if (o >= r0) {
            float *gp = grad + coff + (size_t)o * W + x0 + NCOL * lane;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                if ((stm.colmask >> (8 + j)) & 1u) {
                    const float xv = as * sl[3 * Gm::PITCH + NCOL * lane + j] + ab, yv = sl[3 * Gm::PITCH + 32 * NCOL + NCOL * lane + j];
                    const float df = xv - yv;
                    const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
                    gp[j] = gs * (w_l1 * sgn - w_ssim * (done[0][j] + 2.f * xv * done[1][j] + yv * done[2][j]));
                }
            }
        }
    }
    ph_wait<0>();
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

// launch geometry: strips of 32*NCOL columns, segments of seg_rows rows; one warp per (channel, segment, strip)
struct PhPlan { int ncol, seg_rows, n_seg, n_strip, units, ctas; };

static int ph_env(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

static PhPlan ph_plan(int C, int H, int W) {
    static const int env_ncol = ph_env("GSD_PH_NCOL", 0), env_seg = ph_env("GSD_PH_SEG", 0);
    PhPlan pl;
    pl.ncol = env_ncol == 1 || env_ncol == 2 ? env_ncol : 2;
    pl.n_strip = (W + 32 * pl.ncol - 1) / (32 * pl.ncol);
    // One warp per unit (single-warp CTAs), all units resident at once.  An SM's time is that of its fullest scheduler: with w warps
    // per SM, ceil(w / 4) units of (seg + 10) streamed rows each (10 halo rows are recomputed per segment).  Pick the number of
    // row segments that minimises it among the choices that keep 7..16 warps per SM (fewer: the per-warp dependency chains are
    // exposed; measured).  640x480, 6 planes: 19 segments = 1 140 warps = 8 per SM (two per scheduler) x 36 rows; the 17 segments
    // of the first version (7 per SM: schedulers loaded 2-2-2-1, 39 rows) cost 3.6 us more per iteration.
    const int n_sm = 148;
    int best_seg = H;
    double best = 1e30;
    for (int n_seg = 1; n_seg <= (H + PH_MIN_SEG - 1) / PH_MIN_SEG; ++n_seg) {
        const int seg = (H + n_seg - 1) / n_seg;
        if ((H + seg - 1) / seg != n_seg) continue;
        const long long units = (long long)C * pl.n_strip * n_seg;
        const int per_sm = (int)((units + n_sm - 1) / n_sm);
        double cost = (double)((per_sm + 3) / 4) * (seg + 2 * PH_R);
        if (per_sm < 7) cost *= 7.0 / per_sm;
        if (per_sm > 16) cost *= per_sm / 16.0;
        if (cost < best) { best = cost; best_seg = seg; }
    }
    int seg = env_seg > 0 ? env_seg : best_seg;
    if (seg < PH_MIN_SEG) seg = PH_MIN_SEG;
    if (seg > H) seg = H;
    pl.seg_rows = seg;
    pl.n_seg = (H + seg - 1) / seg;
    pl.units = C * pl.n_seg * pl.n_strip;
    pl.ctas = (pl.units + PH_WARPS - 1) / PH_WARPS;
    return pl;
}

extern "C" int gsd_photometric_workspace_bytes(int32_t C, int32_t H, int32_t W, size_t *bytes) {
    if (C <= 0 || H <= 0 || W <= 0 || !bytes) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    size_t n = (size_t)C * H * W;
    size_t nb = (size_t)C * ((H + PH_MIN_SEG - 1) / PH_MIN_SEG) * ((W + 31) / 32);  // upper bound on the work units
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

template <int MODE>
static void ph_launch_stats(const PhPlan &pl, cudaStream_t st, int C, int H, int W, const PhWin &win, const float *x, const float *y,
                            const float *ls, const float *sh, int aff, float *y_mu, float *y_s22, float *dmu, float *ds11,
                            float *ds12, float *us) {
    if (pl.ncol == 2)
        gsd_launch((gsd_ssim_stats_kernel<MODE, 2>), dim3(pl.ctas), dim3(32 * PH_WARPS), 0, st, C, H, W, pl.seg_rows, pl.n_seg, pl.n_strip, win, x, y, ls, sh,
                                                                         aff, y_mu, y_s22, dmu, ds11, ds12, us);
    else
        gsd_launch((gsd_ssim_stats_kernel<MODE, 1>), dim3(pl.ctas), dim3(32 * PH_WARPS), 0, st, C, H, W, pl.seg_rows, pl.n_seg, pl.n_strip, win, x, y, ls, sh,
                                                                         aff, y_mu, y_s22, dmu, ds11, ds12, us);
}

// statistics pass only: per-pixel SSIM partial maps + per-warp partial sums into ws (what the gradient pass needs)
extern "C" int gsd_photometric_stats(const GsdPhotometric *p, void *stream) {
    int rc;
    if ((rc = ph_check(p))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int C = p->C, H = p->H, W = p->W;
    size_t n = (size_t)C * H * W;
    char *q = (char *)p->ws;
    float *dmu = (float *)q; q += gsd_align_up(n * 4);
    float *ds11 = (float *)q; q += gsd_align_up(n * 4);
    float *ds12 = (float *)q; q += gsd_align_up(n * 4);
    float *bs = (float *)q;
    const PhPlan pl = ph_plan(C, H, W);
    PhWin win = make_window();
    const int aff = (p->affine_log_scale || p->affine_shift) ? 3 : 0;
    if (p->y_mu && p->y_s22)
        ph_launch_stats<1>(pl, st, C, H, W, win, p->x, p->y, p->affine_log_scale, p->affine_shift, aff, (float *)p->y_mu,
                           (float *)p->y_s22, dmu, ds11, ds12, bs);
    else
        ph_launch_stats<0>(pl, st, C, H, W, win, p->x, p->y, p->affine_log_scale, p->affine_shift, aff, nullptr, nullptr, dmu, ds11,
                           ds12, bs);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// fixed-order reduction of the partial sums left by gsd_photometric_stats.  loss_out: per set {loss, mean|x-y|, mean SSIM},
// then the weighted total (3*n_sets + 1 floats); with add != NULL one more float: total + *add (e.g. the prior losses, so
// that the iteration's loss needs no further kernel).  May run on another stream than the gradient pass: it only reads ws.
extern "C" int gsd_photometric_reduce(const GsdPhotometric *p, const float *add, float *loss_out, void *stream) {
    int rc;
    if ((rc = ph_check(p))) return rc;
    if (!loss_out) { gsd_set_error("null loss_out"); return GSD_ERR_INVALID; }
    const int C = p->C, H = p->H, W = p->W;
    size_t n = (size_t)C * H * W;
    const float *bs = (const float *)((const char *)p->ws + 3 * gsd_align_up(n * 4));
    const PhPlan pl = ph_plan(C, H, W);
    const int per_set_c = C / p->n_sets;
    const int units_per_set = pl.n_seg * pl.n_strip * per_set_c;
    gsd_launch(gsd_ssim_finish_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, p->n_sets, units_per_set, bs, 1.0f / (float)((size_t)per_set_c * H * W),
                                                                 p->w_l1, p->w_ssim, p->set_weight[0], p->set_weight[1], add, loss_out);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// loss_out: per set {loss, mean|x-y|, mean SSIM}, then the weighted total  (3*n_sets + 1 floats)
extern "C" int gsd_photometric_forward(const GsdPhotometric *p, float *loss_out, void *stream) {
    int rc;
    if (!loss_out) { gsd_set_error("null loss_out"); return GSD_ERR_INVALID; }
    if ((rc = gsd_photometric_stats(p, stream))) return rc;
    return gsd_photometric_reduce(p, nullptr, loss_out, stream);
}

// window statistics of a target image, computed once and passed as GsdPhotometric.y_mu / y_s22 on later calls
extern "C" int gsd_photometric_target_stats(int32_t C, int32_t H, int32_t W, const float *y, float *y_mu, float *y_s22, void *stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !y || !y_mu || !y_s22) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    const PhPlan pl = ph_plan(C, H, W);
    PhWin win = make_window();
    ph_launch_stats<2>(pl, (cudaStream_t)stream, C, H, W, win, y, y, nullptr, nullptr, 0, y_mu, y_s22, nullptr, nullptr, nullptr,
                       nullptr);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// ---- targets as they are at rest ----------------------------------------------------------------------------------------
// The reference's dataset is PNG files: an 8-bit RGB(A) image and an 8-bit segmentation mask per camera and frame
// (train_utils.py:66-75: `torch.tensor(im).float().cuda().permute(2, 0, 1) / 255`, `torch.stack((seg, zeros, 1 - seg))`).  It converts
// to float32 on the HOST and uploads 6 float planes (7.4 MB per 640x480 camera); uploading the bytes (1.2 MB) and converting here
// gives bit-identical planes (uint8 -> float is exact, the division is IEEE) for a sixth of the PCIe traffic.
__global__ void __launch_bounds__(256)
gsd_unpack_target_u8_kernel(int n_px, int cin, const uint8_t *__restrict__ im, const uint8_t *__restrict__ seg, float *__restrict__ out) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const uint8_t *px = im + (size_t)i * cin;
    out[i] = __fdiv_rn((float)px[0], 255.0f);
    out[(size_t)n_px + i] = __fdiv_rn((float)px[1], 255.0f);
    out[2 * (size_t)n_px + i] = __fdiv_rn((float)px[2], 255.0f);
    const float sg = (float)seg[i];
    out[3 * (size_t)n_px + i] = sg;
    out[4 * (size_t)n_px + i] = 0.0f;
    out[5 * (size_t)n_px + i] = 1.0f - sg;
}

extern "C" int gsd_track_unpack_target_u8(int32_t H, int32_t W, int32_t im_channels, const uint8_t *im_hwc, const uint8_t *seg,
                                          float *target6, void *stream) {
    if (H <= 0 || W <= 0 || (im_channels != 3 && im_channels != 4) || !im_hwc || !seg || !target6) {
        gsd_set_error("invalid arguments (im_channels must be 3 or 4)");
        return GSD_ERR_INVALID;
    }
    const int n = H * W;
    gsd_launch(gsd_unpack_target_u8_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, n, im_channels, im_hwc, seg, target6);
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
    const PhPlan pl = ph_plan(C, H, W);
    PhWin win = make_window();
    const int per_set_c = C / p->n_sets;
    const int aff = (p->affine_log_scale || p->affine_shift) ? 3 : 0;
    const float inv_n = 1.0f / (float)((size_t)per_set_c * H * W);
    if (pl.ncol == 2)
        gsd_launch((gsd_ssim_grad_kernel<2>), dim3(pl.ctas), dim3(32 * PH_WARPS), 0, st, C, H, W, pl.seg_rows, pl.n_seg, pl.n_strip, win, p->x, p->y,
                                                                  p->affine_log_scale, p->affine_shift, aff, dmu, ds11, ds12,
                                                                  gscale_ptr, p->set_weight[0], p->set_weight[1], p->w_l1,
                                                                  p->w_ssim, inv_n, grad);
    else
        gsd_launch((gsd_ssim_grad_kernel<1>), dim3(pl.ctas), dim3(32 * PH_WARPS), 0, st, C, H, W, pl.seg_rows, pl.n_seg, pl.n_strip, win, p->x, p->y,
                                                                  p->affine_log_scale, p->affine_shift, aff, dmu, ds11, ds12,
                                                                  gscale_ptr, p->set_weight[0], p->set_weight[1], p->w_l1,
                                                                  p->w_ssim, inv_n, grad);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
