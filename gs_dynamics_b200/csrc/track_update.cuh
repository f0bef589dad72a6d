// track_update.cuh — per-Gaussian steady-state update (normalize backward + gradient sum + Adam on means3D / unnorm_rotations,
// train_utils.py:370-373 / train_gs.py:31-39), shared by gsd_track_update (adam.cu) and the fused per-Gaussian rasterizer
// backward + update kernel (raster_bwd.cu).
#pragma once
#include "common.cuh"

__device__ __forceinline__ float adam_1(float p, float g, float &m, float &v, float b1, float b2, float eps, float lr_bc1, float inv_sqrt_bc2) {
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    return p - lr_bc1 * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
}

// ga_m[3] / ga_q: the rasterizer's gradients w.r.t. means3D / the NORMALISED rotation, in registers
// Adam's bias-corrected step sizes of the two groups: uniform over the launch (four powf + divisions, ~400 instructions behind a
// dependent load of the step counter) — one thread per CTA computes them at kernel entry, the others read shared memory
struct GsdAdamCoef { float lrm, ism, lrr, isr; };
__device__ __forceinline__ void gsd_track_update_coef(const GsdTrackUpdate &u, GsdAdamCoef *c) {
    const float sm = *u.step_means + 1.f, sr = *u.step_rot + 1.f;
    c->lrm = u.lr_means / (1.f - powf(u.beta1, sm)); c->ism = 1.f / sqrtf(1.f - powf(u.beta2, sm));
    c->lrr = u.lr_rot / (1.f - powf(u.beta1, sr)); c->isr = 1.f / sqrtf(1.f - powf(u.beta2, sr));
}

__device__ __forceinline__ void gsd_track_update_apply(const GsdTrackUpdate &u, const GsdAdamCoef &k4, int i, const float ga_m[3], float4 ga) {
    if (i >= u.G) return;
    const float lrm = k4.lrm, ism = k4.ism, lrr = k4.lrr, isr = k4.isr;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t k = 3 * (size_t)i + c;
        float g = ga_m[c] + (u.g_means_b ? u.g_means_b[k] : 0.f);
        float m = u.m_means[k], v = u.v_means[k];
        u.means3D[k] = adam_1(u.means3D[k], g, m, v, u.beta1, u.beta2, u.eps, lrm, ism);
        u.m_means[k] = m; u.v_means[k] = v;
    }
    float4 q = reinterpret_cast<const float4 *>(u.unnorm_rotations)[i];
    if (u.g_rot_b) {
        float4 gb = reinterpret_cast<const float4 *>(u.g_rot_b)[i];
        ga.x += gb.x; ga.y += gb.y; ga.z += gb.z; ga.w += gb.w;
    }
    const float n = gsd_quat_norm(q);
    const float4 qn = make_float4(__fdiv_rn(q.x, n), __fdiv_rn(q.y, n), __fdiv_rn(q.z, n), __fdiv_rn(q.w, n));
    const float dot = qn.x * ga.x + qn.y * ga.y + qn.z * ga.z + qn.w * ga.w;
    const float gq[4] = {(ga.x - qn.x * dot) / n, (ga.y - qn.y * dot) / n, (ga.z - qn.z * dot) / n, (ga.w - qn.w * dot) / n};
    float4 m4 = reinterpret_cast<float4 *>(u.m_rot)[i], v4 = reinterpret_cast<float4 *>(u.v_rot)[i];
    float qo[4] = {q.x, q.y, q.z, q.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) qo[c] = adam_1(qo[c], gq[c], mm[c], vv[c], u.beta1, u.beta2, u.eps, lrr, isr);
    reinterpret_cast<float4 *>(u.unnorm_rotations)[i] = make_float4(qo[0], qo[1], qo[2], qo[3]);
    reinterpret_cast<float4 *>(u.m_rot)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4 *>(u.v_rot)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
}

__device__ __forceinline__ void gsd_track_update_body(const GsdTrackUpdate &u, const GsdAdamCoef &k4, int i) {
    if (i >= u.G) return;
    const float ga_m[3] = {u.g_means_a[3 * (size_t)i], u.g_means_a[3 * (size_t)i + 1], u.g_means_a[3 * (size_t)i + 2]};
    gsd_track_update_apply(u, k4, i, ga_m, reinterpret_cast<const float4 *>(u.g_rot_a)[i]);
}

// radii bookkeeping of get_loss (train_utils.py:243-245)
__device__ __forceinline__ void gsd_track_update_radii_1(const GsdTrackUpdate &u, int i) {
    if (i < u.G && u.radii) {
        const int r = u.radii[i];
        const bool s = r > 0;
        if (u.seen) u.seen[i] = s ? 1 : 0;
        if (s) u.max_2D_radius[i] = fmaxf((float)r, u.max_2D_radius[i]);
    }
}

// the last CTA to retire advances the step counters (every thread has read them by then) and re-arms the counter
__device__ __forceinline__ void gsd_track_update_advance(const GsdTrackUpdate &u) {
    if (u.block_counter) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(u.block_counter, 1u) == gridDim.x - 1) {
                *u.step_means += 1.0f;
                if (u.step_rot != u.step_means) *u.step_rot += 1.0f;
                *u.block_counter = 0u;
            }
        }
    }
}
