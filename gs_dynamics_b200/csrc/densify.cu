// densify.cu — clone / split / prune of the first-frame densification as two stream-compaction kernels with the Adam-state
// surgery fused in.
//
// Replaces /root/reference/src/tracking/external.py:229-299 (densify) together with cat_params_to_optimizer (160-174),
// remove_points (177-222) and update_params_and_optimizer (145-157): there, one round is ~100 eager PyTorch kernels with boolean-
// mask indexing (a host sync each), three torch.cat per tensor and three re-allocations of all 18 per-point arrays (6 parameters
// + their two Adam moments).  Here:
//   gsd_densify_plan   one CTA: per point the class (keep / clone / split), the prune predicate of every row the round would create,
//                      and — by block-wide scans — the destination row of each of them in the reference's final order
//                      [kept originals | kept clones | kept first split copies | kept second split copies]
//   gsd_densify_apply  one thread per source point: writes its (up to four) destination rows of all 18 arrays; the split samples
//                      x + R(q) eps and the scale / 1.6 are computed on the fly, appended rows get zero Adam moments, the optional
//                      opacity reset of iterations 3000, 6000, ... (external.py:293-295) rides along
// ONE 16-byte read of the totals tells the host how large the new arrays are (the reference syncs ~10 times per round).
#include "common.cuh"

#define DN_THREADS 1024

__device__ __forceinline__ bool dn_prune(float logit_o, float smax, float thr_o, float thr_big) {
    const float o = 1.0f / (1.0f + expf(-logit_o));
    return (o < thr_o) || (thr_big > 0.f && smax > thr_big);
}

// exclusive scan over the CTA of one int per thread; returns the exclusive prefix, *total = sum (valid in all threads)
__device__ __forceinline__ int dn_scan(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    __syncthreads();
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += u;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[wid] + incl - v;
}

__global__ void __launch_bounds__(DN_THREADS)
gsd_densify_plan_kernel(GsdDensifyPlan p) {
    gsd_pdl_wait();
    __shared__ int s_warp[33];
    // running bases of the four output segments and of the split rank
    int base0 = 0, base1 = 0, base2 = 0, bases = 0;
    // pass 1: totals of segment 0 and 1 are needed before the later segments can be placed: two sweeps over the points
    int tot0 = 0, tot1 = 0, tot2 = 0, tots = 0;
    for (int sweep = 0; sweep < 2; ++sweep) {
        base0 = base1 = base2 = bases = 0;
        for (int i0 = 0; i0 < p.n; i0 += DN_THREADS) {
            const int i = i0 + threadIdx.x;
            int k0 = 0, k1 = 0, k2 = 0, sp = 0;
            if (i < p.n) {
                float g = p.grad_accum[i] / p.denom[i];
                if (isnan(g)) g = 0.f;
                const float s0 = expf(p.log_scales[3 * i]), s1 = expf(p.log_scales[3 * i + 1]), s2 = expf(p.log_scales[3 * i + 2]);
                const float smax = fmaxf(s0, fmaxf(s1, s2));
                const bool hot = p.do_densify && g >= p.grad_thresh, big = smax > p.clone_limit;
                const float lo = p.logit_opacities[i];
                const bool pr = p.do_densify && dn_prune(lo, smax, p.prune_opacity, p.prune_big);
                sp = hot && big;
                k0 = !sp && !pr;                     // the original survives unless it is split or pruned
                k1 = hot && !big && !pr;             // its clone has the same values, hence the same prune predicate
                if (sp) {                            // both split copies: scale / 1.6 in the reference's op order (external.py:270)
                    const float t0 = expf(logf(s0 / 1.6f)), t1 = expf(logf(s1 / 1.6f)), t2 = expf(logf(s2 / 1.6f));
                    k2 = !dn_prune(lo, fmaxf(t0, fmaxf(t1, t2)), p.prune_opacity, p.prune_big);
                }
            }
            int t;
            const int e0 = dn_scan(k0, s_warp, &t); const int n0 = t;
            const int e1 = dn_scan(k1, s_warp, &t); const int n1 = t;
            const int e2 = dn_scan(k2, s_warp, &t); const int n2 = t;
            const int es = dn_scan(sp, s_warp, &t); const int ns = t;
            if (sweep == 1 && i < p.n) {
                p.dst[i] = k0 ? base0 + e0 : -1;
                p.dst[p.n + i] = k1 ? tot0 + base1 + e1 : -1;
                p.dst[2 * p.n + i] = k2 ? tot0 + tot1 + base2 + e2 : -1;       // second copy: + tot2
                p.dst[3 * p.n + i] = sp ? bases + es : -1;                      // rank among ALL split candidates (sample row)
            }
            base0 += n0; base1 += n1; base2 += n2; bases += ns;
        }
        tot0 = base0; tot1 = base1; tot2 = base2; tots = bases;
    }
    if (threadIdx.x == 0) { p.totals[0] = tot0; p.totals[1] = tot1; p.totals[2] = tot2; p.totals[3] = tots; }
}

__global__ void __launch_bounds__(256)
gsd_densify_apply_kernel(GsdDensifyApply a) {
    gsd_pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int d0 = a.dst[i], d1 = a.dst[a.n + i], d2 = a.dst[2 * a.n + i], sr = a.dst[3 * a.n + i];
    const int tot2 = a.totals[2], tots = a.totals[3];
    const float reset = logf(0.01f / 0.99f);   // inverse_sigmoid(0.01), external.py:225-226,294
    // split samples: x + R(q / |q|) eps, eps ~ N(0, diag(exp(log_scales))^2) — rows sr and sr + (number of split candidates)
    float mean_a[3], mean_b[3], ls_split[3];
    if (d2 >= 0) {
        const float q0 = a.p_src[3][4 * i], q1 = a.p_src[3][4 * i + 1], q2 = a.p_src[3][4 * i + 2], q3 = a.p_src[3][4 * i + 3];
        const float inv = 1.0f / sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
        const float r = q0 * inv, x = q1 * inv, y = q2 * inv, z = q3 * inv;
        const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                               {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                               {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        float sd[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { sd[c] = expf(a.p_src[5][3 * i + c]); ls_split[c] = logf(sd[c] / 1.6f); }
        float ea[3], eb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ea[c] = a.samples[3 * (size_t)sr + c]; eb[c] = a.samples[3 * ((size_t)sr + tots) + c];
            if (!a.samples_scaled) { ea[c] *= sd[c]; eb[c] *= sd[c]; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mean_a[c] = a.p_src[0][3 * i + c] + (R[c][0] * ea[0] + R[c][1] * ea[1] + R[c][2] * ea[2]);
            mean_b[c] = a.p_src[0][3 * i + c] + (R[c][0] * eb[0] + R[c][1] * eb[1] + R[c][2] * eb[2]);
        }
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const int w = a.width[t];
        for (int c = 0; c < w; ++c) {
            float v = a.p_src[t][(size_t)i * w + c];
            float m = a.m_src[t][(size_t)i * w + c], s = a.v_src[t][(size_t)i * w + c];
            const bool is_opac = (t == 4);
            if (is_opac && a.reset_opacity) { m = 0.f; s = 0.f; }   // update_params_and_optimizer: moments restart at zero
            if (d0 >= 0) {
                a.p_dst[t][(size_t)d0 * w + c] = (is_opac && a.reset_opacity) ? reset : v;
                a.m_dst[t][(size_t)d0 * w + c] = m;
                a.v_dst[t][(size_t)d0 * w + c] = s;
            }
            if (d1 >= 0) {      // clone: same values, fresh moments (cat_params_to_optimizer)
                a.p_dst[t][(size_t)d1 * w + c] = (is_opac && a.reset_opacity) ? reset : v;
                a.m_dst[t][(size_t)d1 * w + c] = 0.f;
                a.v_dst[t][(size_t)d1 * w + c] = 0.f;
            }
            if (d2 >= 0) {
                float va = v, vb = v;
                if (t == 0) { va = mean_a[c]; vb = mean_b[c]; }
                if (t == 5) { va = ls_split[c]; vb = ls_split[c]; }
                if (is_opac && a.reset_opacity) { va = reset; vb = reset; }
                a.p_dst[t][(size_t)d2 * w + c] = va;
                a.p_dst[t][(size_t)(d2 + tot2) * w + c] = vb;
                a.m_dst[t][(size_t)d2 * w + c] = 0.f; a.m_dst[t][(size_t)(d2 + tot2) * w + c] = 0.f;
                a.v_dst[t][(size_t)d2 * w + c] = 0.f; a.v_dst[t][(size_t)(d2 + tot2) * w + c] = 0.f;
            }
        }
    }
}

extern "C" int gsd_densify_plan(const GsdDensifyPlan *p, void *stream) {
    if (!p || p->n < 0 || (p->n > 0 && (!p->grad_accum || !p->denom || !p->log_scales || !p->logit_opacities || !p->dst)) || !p->totals) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    gsd_launch(gsd_densify_plan_kernel, dim3(1), dim3(DN_THREADS), 0, (cudaStream_t)stream, *p);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

extern "C" int gsd_densify_apply(const GsdDensifyApply *a, void *stream) {
    if (!a || a->n < 0 || !a->dst || !a->totals) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    for (int t = 0; t < 6; ++t)
        if (a->n > 0 && (!a->p_src[t] || !a->m_src[t] || !a->v_src[t] || !a->p_dst[t] || !a->m_dst[t] || !a->v_dst[t] || a->width[t] <= 0)) {
            gsd_set_error("null tensor pointer (tensor %d)", t);
            return GSD_ERR_INVALID;
        }
    if (a->n == 0) return GSD_OK;
    gsd_launch(gsd_densify_apply_kernel, dim3((a->n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, *a);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
