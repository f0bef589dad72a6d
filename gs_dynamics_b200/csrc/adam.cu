// adam.cu — multi-tensor Adam (all parameter groups of the tracker in one launch) and the small per-iteration
// bookkeeping kernel.  Replaces torch.optim.Adam's 8 per-group update sequences
// (/root/reference/src/tracking/train_utils.py:152-164, train_gs.py:38-39) and the boolean-mask indexing of
// train_utils.py:243-245 (which forces host syncs).  HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"

struct AdamTable {
    float *param[GSD_ADAM_MAX_TENSORS];
    const float *grad[GSD_ADAM_MAX_TENSORS];
    float *m[GSD_ADAM_MAX_TENSORS];
    float *v[GSD_ADAM_MAX_TENSORS];
    float *step[GSD_ADAM_MAX_TENSORS];
    float lr[GSD_ADAM_MAX_TENSORS];
    long long start[GSD_ADAM_MAX_TENSORS + 1]; // prefix of numel in elements
    int n;
    float beta1, beta2, eps;
};

__global__ void __launch_bounds__(256)
gsd_adam_kernel(AdamTable t) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    const long long total = t.start[t.n];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int k = 0;
#pragma unroll 1
        while (k + 1 < t.n && i >= t.start[k + 1]) ++k;
        const long long j = i - t.start[k];
        const float stepv = *t.step[k] + 1.0f; // the counter itself is advanced by the tail kernel
        const float g = t.grad[k][j];
        float m = t.m[k][j], v = t.v[k][j];
        m = t.beta1 * m + (1.f - t.beta1) * g;
        v = t.beta2 * v + (1.f - t.beta2) * g * g;
        t.m[k][j] = m;
        t.v[k][j] = v;
        const float bc1 = 1.f - powf(t.beta1, stepv);
        const float bc2 = 1.f - powf(t.beta2, stepv);
        const float denom = sqrtf(v) / sqrtf(bc2) + t.eps;
        t.param[k][j] -= (t.lr[k] / bc1) * (m / denom);
    }
}
__global__ void gsd_adam_advance_kernel(AdamTable t) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    int k = threadIdx.x;
    if (k < t.n) *t.step[k] += 1.0f;
}

extern "C" int gsd_adam_step(const GsdAdam *a, void *stream) {
    if (!a || a->n_tensors < 0 || a->n_tensors > GSD_ADAM_MAX_TENSORS) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    AdamTable t;
    t.n = 0; t.beta1 = a->beta1; t.beta2 = a->beta2; t.eps = a->eps;
    long long off = 0;
    // distinct step counters may be shared between tensors: advance each distinct pointer once
    AdamTable adv; adv.n = 0;
    for (int i = 0; i < a->n_tensors; ++i) {
        if (a->numel[i] <= 0) continue;
        if (!a->param[i] || !a->grad[i] || !a->exp_avg[i] || !a->exp_avg_sq[i] || !a->step[i]) { gsd_set_error("null tensor %d", i); return GSD_ERR_INVALID; }
        t.param[t.n] = a->param[i]; t.grad[t.n] = a->grad[i]; t.m[t.n] = a->exp_avg[i]; t.v[t.n] = a->exp_avg_sq[i];
        t.step[t.n] = a->step[i]; t.lr[t.n] = a->lr[i]; t.start[t.n] = off;
        off += a->numel[i];
        ++t.n;
        bool seen = false;
        for (int k = 0; k < adv.n; ++k) seen |= (adv.step[k] == a->step[i]);
        if (!seen) adv.step[adv.n++] = a->step[i];
    }
    t.start[t.n] = off;
    if (t.n == 0) return GSD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = (off + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gsd_launch(gsd_adam_kernel, dim3((unsigned)blocks), dim3(256), 0, st, t);
    GSD_LAUNCH_CHECK();
    gsd_launch(gsd_adam_advance_kernel, dim3(1), dim3(32), 0, st, adv);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

__global__ void gsd_update_radii_kernel(int G, const int32_t *__restrict__ radii, float *__restrict__ max_r, uint8_t *__restrict__ seen) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    int r = radii[i];
    bool s = r > 0;
    if (seen) seen[i] = s ? 1 : 0;
    if (s) max_r[i] = fmaxf((float)r, max_r[i]);
}

extern "C" int gsd_track_update_radii(int32_t G, const int32_t *radii, float *max_2D_radius, uint8_t *seen, void *stream) {
    if (G < 0 || (G > 0 && (!radii || !max_2D_radius))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (G == 0) return GSD_OK;
    gsd_launch(gsd_update_radii_kernel, dim3((G + 255) / 256), dim3(256), 0, (cudaStream_t)stream, G, radii, max_2D_radius, seen);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// ------------------------------------------------------------------------------------------------------
// steady-state fast path: F.normalize forward, and (normalize backward + gradient sum + Adam) for the two live groups
// ------------------------------------------------------------------------------------------------------
__global__ void gsd_normalize_rot_kernel(int G, const float4 *__restrict__ q, float4 *__restrict__ out) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    float4 v = q[i];
    float n = fmaxf(sqrtf(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w), 1e-12f); // F.normalize eps
    out[i] = make_float4(v.x / n, v.y / n, v.z / n, v.w / n);
}

extern "C" int gsd_track_normalize_rotations(int32_t G, const float *unnorm, float *rot, void *stream) {
    if (G < 0 || (G > 0 && (!unnorm || !rot))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (G == 0) return GSD_OK;
    gsd_launch(gsd_normalize_rot_kernel, dim3((G + 255) / 256), dim3(256), 0, (cudaStream_t)stream, G, (const float4 *)unnorm, (float4 *)rot);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

__device__ __forceinline__ float adam_1(float p, float g, float &m, float &v, float b1, float b2, float eps, float lr_bc1, float inv_sqrt_bc2) {
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    return p - lr_bc1 * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
}

__device__ __forceinline__ void gsd_track_update_body(const GsdTrackUpdate &u, int i) {
    if (i >= u.G) return;
    const float sm = *u.step_means + 1.f, sr = *u.step_rot + 1.f;
    const float lrm = u.lr_means / (1.f - powf(u.beta1, sm)), ism = 1.f / sqrtf(1.f - powf(u.beta2, sm));
    const float lrr = u.lr_rot / (1.f - powf(u.beta1, sr)), isr = 1.f / sqrtf(1.f - powf(u.beta2, sr));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t k = 3 * (size_t)i + c;
        float g = u.g_means_a[k] + (u.g_means_b ? u.g_means_b[k] : 0.f);
        float m = u.m_means[k], v = u.v_means[k];
        u.means3D[k] = adam_1(u.means3D[k], g, m, v, u.beta1, u.beta2, u.eps, lrm, ism);
        u.m_means[k] = m; u.v_means[k] = v;
    }
    float4 q = reinterpret_cast<const float4 *>(u.unnorm_rotations)[i];
    float4 ga = reinterpret_cast<const float4 *>(u.g_rot_a)[i];
    if (u.g_rot_b) {
        float4 gb = reinterpret_cast<const float4 *>(u.g_rot_b)[i];
        ga.x += gb.x; ga.y += gb.y; ga.z += gb.z; ga.w += gb.w;
    }
    const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    const float4 qn = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
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

// the per-Gaussian update + (optionally) the radii bookkeeping and the step advance in ONE launch: the last CTA to retire
// advances the step counters (every thread has read them by then) and re-arms the counter
__global__ void __launch_bounds__(256)
gsd_track_update_fused_kernel(GsdTrackUpdate u) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < u.G && u.radii) {
        const int r = u.radii[i];
        const bool s = r > 0;
        if (u.seen) u.seen[i] = s ? 1 : 0;
        if (s) u.max_2D_radius[i] = fmaxf((float)r, u.max_2D_radius[i]);
    }
    gsd_track_update_body(u, i);
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
__global__ void gsd_track_update_advance_kernel(float *a, float *b) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    *a += 1.0f;
    if (b != a) *b += 1.0f;
}

extern "C" int gsd_track_update(const GsdTrackUpdate *u, void *stream) {
    if (!u || u->G < 0) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (u->G == 0) return GSD_OK;
    if (!u->means3D || !u->unnorm_rotations || !u->g_means_a || !u->g_rot_a || !u->m_means || !u->v_means || !u->m_rot || !u->v_rot ||
        !u->step_means || !u->step_rot) {
        gsd_set_error("null pointer in GsdTrackUpdate");
        return GSD_ERR_INVALID;
    }
    if (u->radii && !u->max_2D_radius) { gsd_set_error("radii given without max_2D_radius"); return GSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    gsd_launch(gsd_track_update_fused_kernel, dim3((u->G + 255) / 256), dim3(256), 0, st, *u);
    GSD_LAUNCH_CHECK();
    if (!u->block_counter) {
        gsd_launch(gsd_track_update_advance_kernel, dim3(1), dim3(1), 0, st, u->step_means, u->step_rot);
        GSD_LAUNCH_CHECK();
    }
    return GSD_OK;
}
