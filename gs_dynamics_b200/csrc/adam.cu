// adam.cu — multi-tensor Adam (all parameter groups of the tracker in one launch) and the small per-iteration
// bookkeeping kernel.  Replaces torch.optim.Adam's 8 per-group update sequences
// (/root/reference/src/tracking/train_utils.py:152-164, train_gs.py:38-39) and the boolean-mask indexing of
// train_utils.py:243-245 (which forces host syncs).  HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"
#include "track_update.cuh"

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
    out[i] = gsd_quat_normalize(q[i]);
}

extern "C" int gsd_track_normalize_rotations(int32_t G, const float *unnorm, float *rot, void *stream) {
    if (G < 0 || (G > 0 && (!unnorm || !rot))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (G == 0) return GSD_OK;
    gsd_launch(gsd_normalize_rot_kernel, dim3((G + 255) / 256), dim3(256), 0, (cudaStream_t)stream, G, (const float4 *)unnorm, (float4 *)rot);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// the per-Gaussian update + (optionally) the radii bookkeeping and the step advance in ONE launch: the last CTA to retire
// advances the step counters (every thread has read them by then) and re-arms the counter
__global__ void __launch_bounds__(256)
gsd_track_update_fused_kernel(GsdTrackUpdate u) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    __shared__ GsdAdamCoef coef;
    if (threadIdx.x == 0) gsd_track_update_coef(u, &coef);
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    gsd_track_update_radii_1(u, i);
    gsd_track_update_body(u, coef, i);
    gsd_track_update_advance(u);
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
