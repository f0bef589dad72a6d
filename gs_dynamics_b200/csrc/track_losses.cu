// track_losses.cu — the physical priors of the tracking iteration, forward + backward in one pass:
//   rigid  = mean sqrt(w * |R_i^T (x_j - x_i) - prev_offset_ik|^2 + 1e-20)
//   rot    = mean sqrt(w * |rel_j - rel_i|^2 + 1e-20),   rel = q (x) prev_inv_q
//   iso    = mean sqrt(w * (sqrt(|x_j - x_i|^2 + 1e-20) - d0_ik)^2 + 1e-20)
//   floor  = mean clamp(y_fg, min=0);   bg = mean |x_bg - x0|_1 + mean |q_bg - q0|_1
// Reference: /root/reference/src/tracking/train_utils.py:198-240 (get_loss), helpers.py:79-94
// (weighted_l2_loss_v1/v2, quat_mult), external.py:25-42 (build_rotation).
//
// One thread per foreground point.  Its own gradient (x_i, q_i) is accumulated in registers over its K out-edges;
// the gradient it receives as somebody's neighbour is accumulated by walking its IN-edges (transposed CSR built once
// per episode — the kNN graph is static) and recomputing those edges.  No atomics, no [G,K,3] temporaries, fixed
// summation order (the reference materialises ~15 such temporaries and scatters with index_put atomics).
// Algorithmic bytes: 56*G*K + 72*G per call; node data is L2-resident.
#include "common.cuh"
#include <cstdlib>

struct TrackArgs {
    int Gf, K, Gb;
    int unnorm;            // q holds un-normalised quaternions: F.normalize on the fly
    unsigned div_magic; int div_shift;   // e / K == (e * div_magic) >> div_shift for every 31-bit e
    const float *x;        // [G,3] means3D
    const float *q;        // [G,4] normalised rotations
    const int32_t *fg_index; // [Gf] or null (identity)
    const float *prev_inv; // [Gf,4]
    const int32_t *nbr;    // [Gf,K] indices into the fg set
    const float *nbr_w, *nbr_d; // [Gf,K]
    const float *prev_off; // [Gf,K,3]
    const int32_t *in_ptr; // [Gf+1]
    const int32_t *in_edge;// [Gf*K] edge ids (i*K+k) grouped by neighbour
    const int32_t *bg_index; // [Gb]
    const float *bg_x0, *bg_q0; // [Gb,3], [Gb,4]
    float c_rigid, c_rot, c_iso, c_floor, c_bg; // weight / count
    float *grad_x, *grad_q; // [G,3], [G,4]
    float *block_sums;     // [nblocks][5]
};

struct Quat { float w, x, y, z; };

__device__ __forceinline__ Quat qmul(Quat a, Quat b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}
// gradient w.r.t. a of qmul(a, b) given gradient g w.r.t. the product
__device__ __forceinline__ Quat qmul_bwd_a(Quat g, Quat b) {
    Quat r;
    r.w = b.w * g.w + b.x * g.x + b.y * g.y + b.z * g.z;
    r.x = -b.x * g.w + b.w * g.x - b.z * g.y + b.y * g.z;
    r.y = -b.y * g.w + b.z * g.x + b.w * g.y - b.x * g.z;
    r.z = -b.z * g.w - b.y * g.x + b.x * g.y + b.w * g.z;
    return r;
}
__device__ __forceinline__ void rot_from_unit(Quat n, float R[3][3]) {
    float r = n.w, x = n.x, y = n.y, z = n.z;
    R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
    R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
    R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
}
__device__ __forceinline__ Quat load_q(const float *p, int i) {
    float4 v = *reinterpret_cast<const float4 *>(p + 4 * (size_t)i);
    return Quat{v.x, v.y, v.z, v.w};
}
// the current rotation of Gaussian gi as the reference's get_loss sees it: F.normalize(unnorm_rotations) (helpers.py:40)
__device__ __forceinline__ Quat load_rot(const TrackArgs &a, int gi) {
    float4 v = *reinterpret_cast<const float4 *>(a.q + 4 * (size_t)gi);
    if (a.unnorm) v = gsd_quat_normalize(v);
    return Quat{v.x, v.y, v.z, v.w};
}

struct EdgeOut { float g_off[3]; float dLde[3]; float g_rel[4]; float s1, s2, s3; };

// MUFU.RSQ: every argument below is >= 1e-20 (normal range), error 2 ulp.  The IEEE sqrtf / division sequences this replaces were
// 7 x ~15 instructions (+ slow-path calls) of a ~190-instruction edge evaluation; sqrt(a) = a * rsqrt(a), w / sqrt(a) = w * rsqrt(a).
__device__ __forceinline__ float trk_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// evaluates one edge i -> j; g_off = dL/d(x_j - x_i), dLde = dL/d(R_i^T off - prev), g_rel = dL/d rel_j (= -dL/d rel_i)
__device__ __forceinline__ void eval_edge(const TrackArgs &a, const float xi[3], const float Ri[3][3], Quat rel_i,
                                          const float xj[3], Quat rel_j, float w, float d0, const float po[3],
                                          EdgeOut &o) {
    float off[3] = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
    float e[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) e[c] = Ri[0][c] * off[0] + Ri[1][c] * off[1] + Ri[2][c] * off[2] - po[c];
    const float a1 = w * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 1e-20f;
    const float r1 = trk_rsqrt(a1);
    const float cr = a.c_rigid * w * r1;
#pragma unroll
    for (int c = 0; c < 3; ++c) o.dLde[c] = cr * e[c];
#pragma unroll
    for (int b = 0; b < 3; ++b) o.g_off[b] = Ri[b][0] * o.dLde[0] + Ri[b][1] * o.dLde[1] + Ri[b][2] * o.dLde[2];
    float d[4] = {rel_j.w - rel_i.w, rel_j.x - rel_i.x, rel_j.y - rel_i.y, rel_j.z - rel_i.z};
    const float a2 = w * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]) + 1e-20f;
    const float r2 = trk_rsqrt(a2);
    const float crot = a.c_rot * w * r2;
#pragma unroll
    for (int c = 0; c < 4; ++c) o.g_rel[c] = crot * d[c];
    const float m2 = off[0] * off[0] + off[1] * off[1] + off[2] * off[2] + 1e-20f;
    const float rm = trk_rsqrt(m2);
    const float dm = m2 * rm - d0;
    const float a3 = w * dm * dm + 1e-20f;
    const float r3 = trk_rsqrt(a3);
    const float ciso = a.c_iso * w * dm * r3 * rm;
#pragma unroll
    for (int c = 0; c < 3; ++c) o.g_off[c] += ciso * off[c];
    o.s1 = a1 * r1; o.s2 = a2 * r2; o.s3 = a3 * r3;
}

// 4 lanes per foreground point (each takes every 4th out-/in-edge, fixed-order quad reduction): 4x more warps in flight and
// 4x shorter serial gather chains than one thread per point — the kernel is latency-bound on dependent gathers.
#define TRK_SPLIT 4
// Gathers in flight per lane (out-edges / in-edges) and resident CTAs per SM, measured inside the tracking iteration at 100k
// (tools/step_ablate.py; the kernel runs on a side branch of the iteration's graph and competes with the render branch):
// (5, 3, 4 CTAs: 112 registers) 334.7 us per iteration, (5, 2, 6) 329.1, (3, 2, 7) 327.0, (2, 2, 8) 326.1, (3, 2, 8: 64 registers,
// 36 bytes of spills) 325.8 — occupancy beats loads in flight once each record is one load instruction.
#ifndef TRK_UO
#define TRK_UO 3
#endif
#ifndef TRK_UI
#define TRK_UI 2
#endif
#ifndef TRK_MIN_CTAS
#define TRK_MIN_CTAS 8
#endif
#ifndef TRK_THREADS
#define TRK_THREADS 128   // threads per CTA of the packed kernel (a multiple of 128)
#endif
#define GSD_PRIORS_CTAS_PER_SM_DEFAULT 0
#define GSD_PRIORS_CARVEOUT_DEFAULT (-1)
#define GSD_PRIORS_PIECES_DEFAULT 1
__global__ void __launch_bounds__(128)
gsd_track_fg_kernel(TrackArgs a) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    __shared__ float red[4][4];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = tid / TRK_SPLIT, sub = tid % TRK_SPLIT;
    float s_rigid = 0.f, s_rot = 0.f, s_iso = 0.f, s_floor = 0.f;
    const bool active = f < a.Gf;
    int gi = 0;
    float xi[3] = {0.f, 0.f, 0.f};
    Quat pi = {1.f, 0.f, 0.f, 0.f}, n_i = {1.f, 0.f, 0.f, 0.f};
    float inv_n = 1.f;
    float gx[3] = {0.f, 0.f, 0.f}, grel[4] = {0.f, 0.f, 0.f, 0.f};
    float Gm[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    if (active) {
        gi = a.fg_index ? a.fg_index[f] : f;
        xi[0] = a.x[3 * (size_t)gi]; xi[1] = a.x[3 * (size_t)gi + 1]; xi[2] = a.x[3 * (size_t)gi + 2];
        Quat qi = load_rot(a, gi);
        pi = load_q(a.prev_inv, f);
        Quat rel_i = qmul(qi, pi);
        float nrm = sqrtf(rel_i.w * rel_i.w + rel_i.x * rel_i.x + rel_i.y * rel_i.y + rel_i.z * rel_i.z);
        inv_n = 1.f / nrm;
        n_i = Quat{rel_i.w * inv_n, rel_i.x * inv_n, rel_i.y * inv_n, rel_i.z * inv_n};
        float Ri[3][3];
        rot_from_unit(n_i, Ri);
        // ---- out-edges: f is the centre point
        for (int k = sub; k < a.K; k += TRK_SPLIT) {
            const size_t e = (size_t)f * a.K + k;
            const int j = a.nbr[e];
            const int gj = a.fg_index ? a.fg_index[j] : j;
            float xj[3] = {a.x[3 * (size_t)gj], a.x[3 * (size_t)gj + 1], a.x[3 * (size_t)gj + 2]};
            Quat rel_j = qmul(load_rot(a, gj), load_q(a.prev_inv, j));
            float po[3] = {a.prev_off[3 * e], a.prev_off[3 * e + 1], a.prev_off[3 * e + 2]};
            EdgeOut o;
            eval_edge(a, xi, Ri, rel_i, xj, rel_j, a.nbr_w[e], a.nbr_d[e], po, o);
            float off[3] = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                gx[b] -= o.g_off[b];
#pragma unroll
                for (int c = 0; c < 3; ++c) Gm[b][c] += off[b] * o.dLde[c];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) grel[c] -= o.g_rel[c];
            s_rigid += o.s1; s_rot += o.s2; s_iso += o.s3;
        }
        // ---- in-edges: f is the neighbour of some centre i2
        const int e0 = a.in_ptr[f], e1 = a.in_ptr[f + 1];
        for (int t = e0 + sub; t < e1; t += TRK_SPLIT) {
            const int e = a.in_edge[t];
            const int i2 = e / a.K;
            const int g2 = a.fg_index ? a.fg_index[i2] : i2;
            float x2[3] = {a.x[3 * (size_t)g2], a.x[3 * (size_t)g2 + 1], a.x[3 * (size_t)g2 + 2]};
            Quat rel_2 = qmul(load_rot(a, g2), load_q(a.prev_inv, i2));
            float n2 = rsqrtf(rel_2.w * rel_2.w + rel_2.x * rel_2.x + rel_2.y * rel_2.y + rel_2.z * rel_2.z);
            Quat u2 = {rel_2.w * n2, rel_2.x * n2, rel_2.y * n2, rel_2.z * n2};
            float R2[3][3];
            rot_from_unit(u2, R2);
            float po[3] = {a.prev_off[3 * (size_t)e], a.prev_off[3 * (size_t)e + 1], a.prev_off[3 * (size_t)e + 2]};
            EdgeOut o;
            eval_edge(a, x2, R2, rel_2, xi, rel_i, a.nbr_w[e], a.nbr_d[e], po, o);
#pragma unroll
            for (int b = 0; b < 3; ++b) gx[b] += o.g_off[b];
#pragma unroll
            for (int c = 0; c < 4; ++c) grel[c] += o.g_rel[c];
        }
    }
    // fixed-order reduction over the TRK_SPLIT lanes of a point (all lanes of the warp participate)
#pragma unroll
    for (int o = 1; o < TRK_SPLIT; o <<= 1) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            gx[b] += __shfl_xor_sync(0xffffffffu, gx[b], o);
#pragma unroll
            for (int c = 0; c < 3; ++c) Gm[b][c] += __shfl_xor_sync(0xffffffffu, Gm[b][c], o);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) grel[c] += __shfl_xor_sync(0xffffffffu, grel[c], o);
    }
    if (active && sub == 0) {
        // dL/dR_i -> dL/dn_i -> dL/drel_i (through the normalisation inside build_rotation)
        {
            float r = n_i.w, x = n_i.x, y = n_i.y, z = n_i.z;
            float gn[4];
            gn[0] = 2.f * (-z * Gm[0][1] + y * Gm[0][2] + z * Gm[1][0] - x * Gm[1][2] - y * Gm[2][0] + x * Gm[2][1]);
            gn[1] = 2.f * (y * Gm[0][1] + z * Gm[0][2] + y * Gm[1][0] - 2.f * x * Gm[1][1] - r * Gm[1][2] + z * Gm[2][0] + r * Gm[2][1] - 2.f * x * Gm[2][2]);
            gn[2] = 2.f * (-2.f * y * Gm[0][0] + x * Gm[0][1] + r * Gm[0][2] + x * Gm[1][0] + z * Gm[1][2] - r * Gm[2][0] + z * Gm[2][1] - 2.f * y * Gm[2][2]);
            gn[3] = 2.f * (-2.f * z * Gm[0][0] - r * Gm[0][1] + x * Gm[0][2] + r * Gm[1][0] - 2.f * z * Gm[1][1] + y * Gm[1][2] + x * Gm[2][0] + y * Gm[2][1]);
            float dot = r * gn[0] + x * gn[1] + y * gn[2] + z * gn[3];
            grel[0] += (gn[0] - r * dot) * inv_n;
            grel[1] += (gn[1] - x * dot) * inv_n;
            grel[2] += (gn[2] - y * dot) * inv_n;
            grel[3] += (gn[3] - z * dot) * inv_n;
        }
        // floor
        if (xi[1] > 0.f) { gx[1] += a.c_floor; s_floor = xi[1]; }
        Quat gq = qmul_bwd_a(Quat{grel[0], grel[1], grel[2], grel[3]}, pi);
        a.grad_x[3 * (size_t)gi] = gx[0]; a.grad_x[3 * (size_t)gi + 1] = gx[1]; a.grad_x[3 * (size_t)gi + 2] = gx[2];
        *reinterpret_cast<float4 *>(a.grad_q + 4 * (size_t)gi) = make_float4(gq.w, gq.x, gq.y, gq.z);
    }
    // block partial sums, fixed order
    float v[4] = {s_rigid, s_rot, s_iso, s_floor};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int c = 0; c < 4; ++c) red[threadIdx.x >> 5][c] = v[c];
    __syncthreads();
    if (threadIdx.x < 4) {
        float s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
        a.block_sums[5 * (size_t)blockIdx.x + threadIdx.x] = s;
    }
    if (threadIdx.x == 4) a.block_sums[5 * (size_t)blockIdx.x + 4] = 0.f;
}

// ---- packed variant -----------------------------------------------------------------------------------------------
// Per iteration a tiny kernel writes one 32-byte record per foreground point: (x, y, z, 0 | rel = q (x) prev_inv_q); the static
// per-edge tables are packed once per timestep into 32-byte records (neighbour id, weight, rest distance, previous offset).
// Every edge evaluation is then two 32-byte-aligned 256-bit loads (edge record + node record) instead of ~10 scattered 4..16-byte
// loads and a quaternion product.  The kernel is bound by the L1 tag stage — a gather instruction costs one pass per distinct
// 128-byte line its lanes touch (ncu: L1/TEX throughput 89 %, DRAM 16 %, issue slots 27 %) — so what counts is the number of
// load INSTRUCTIONS per edge: one LDG.256 per record (two LDG.128 per record + a 64-byte node record with the rotation matrix
// precomputed measured 65-72 us whatever the occupancy or the number of loads in flight).
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const float4 *p) {   // p 32-byte aligned; read-only data
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    return r;
}

__global__ void __launch_bounds__(256)
gsd_track_node_prep_kernel(TrackArgs a, float4 *__restrict__ node) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.Gf) return;
    const int gi = a.fg_index ? a.fg_index[f] : f;
    const Quat rel = qmul(load_rot(a, gi), load_q(a.prev_inv, f));
    node[2 * (size_t)f] = make_float4(a.x[3 * (size_t)gi], a.x[3 * (size_t)gi + 1], a.x[3 * (size_t)gi + 2], 0.f);
    node[2 * (size_t)f + 1] = make_float4(rel.w, rel.x, rel.y, rel.z);
}

__global__ void __launch_bounds__(TRK_THREADS, TRK_MIN_CTAS)
gsd_track_fg_packed_kernel(TrackArgs a, const float4 *__restrict__ node, const float4 *__restrict__ edge, int vb_first, int n_vblocks) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    constexpr int NG = TRK_THREADS / 128;   // virtual blocks (128 threads = 32 points, one row of block_sums each) side by side in a CTA
    __shared__ float red_all[NG][4][4];
    float (*red)[4] = red_all[threadIdx.x >> 7];
    const int tg = threadIdx.x & 127;
    // the grid may be smaller than the number of 128-thread blocks of work (see gsd_track_losses_fwd_bwd): a CTA then walks a
    // CONTIGUOUS run of virtual blocks (consecutive blocks are neighbours on the Morton curve: the node records one block pulled
    // into L1 are the next block's neighbours too); sums are kept per virtual block (fixed order, grid-independent)
    const int n_grp = (int)gridDim.x * NG;
    const int my_grp = (int)blockIdx.x * NG + (int)(threadIdx.x >> 7);
    const int vb_per = (n_vblocks + n_grp - 1) / n_grp;
    const int vb_end = min(n_vblocks, (my_grp + 1) * vb_per);
    for (int it = 0; it < vb_per; ++it) {     // uniform trip count over the CTA (block barriers inside); surplus groups idle
    const int vb = vb_first + my_grp * vb_per + it;
    const bool vb_ok = my_grp * vb_per + it < vb_end;
    const int tid = vb * 128 + tg;
    const int f = tid / TRK_SPLIT, sub = tid % TRK_SPLIT;
    float s_rigid = 0.f, s_rot = 0.f, s_iso = 0.f, s_floor = 0.f;
    const bool active = vb_ok && f < a.Gf;
    float xi[3] = {0.f, 0.f, 0.f};
    Quat n_i = {1.f, 0.f, 0.f, 0.f};
    float inv_n = 1.f;
    float gx[3] = {0.f, 0.f, 0.f}, grel[4] = {0.f, 0.f, 0.f, 0.f};
    float Gm[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    if (active) {
        const F8 n0 = ldg256(node + 2 * (size_t)f);
        xi[0] = n0.a.x; xi[1] = n0.a.y; xi[2] = n0.a.z;
        const Quat rel_i = {n0.b.x, n0.b.y, n0.b.z, n0.b.w};
        inv_n = trk_rsqrt(rel_i.w * rel_i.w + rel_i.x * rel_i.x + rel_i.y * rel_i.y + rel_i.z * rel_i.z);
        n_i = Quat{rel_i.w * inv_n, rel_i.x * inv_n, rel_i.y * inv_n, rel_i.z * inv_n};
        float Ri[3][3];
        rot_from_unit(n_i, Ri);
        // The loads of TRK_UO / TRK_UI edges are issued together, then the arithmetic runs.  Slots past the end re-read a valid
        // record and are skipped in the arithmetic.
        for (int k0 = sub; k0 < a.K; k0 += TRK_SPLIT * TRK_UO) {
            F8 E[TRK_UO], M[TRK_UO];
#pragma unroll
            for (int u = 0; u < TRK_UO; ++u) {
                const int k = k0 + TRK_SPLIT * u;
                E[u] = ldg256(edge + 2 * ((size_t)f * a.K + (k < a.K ? k : k0)));
            }
#pragma unroll
            for (int u = 0; u < TRK_UO; ++u) M[u] = ldg256(node + 2 * (size_t)__float_as_int(E[u].a.x));
#pragma unroll
            for (int u = 0; u < TRK_UO; ++u) {
                if (k0 + TRK_SPLIT * u >= a.K) break;
                const float xj[3] = {M[u].a.x, M[u].a.y, M[u].a.z};
                const float po[3] = {E[u].a.w, E[u].b.x, E[u].b.y};
                EdgeOut o;
                eval_edge(a, xi, Ri, rel_i, xj, Quat{M[u].b.x, M[u].b.y, M[u].b.z, M[u].b.w}, E[u].a.y, E[u].a.z, po, o);
                const float off[3] = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    gx[b] -= o.g_off[b];
#pragma unroll
                    for (int c = 0; c < 3; ++c) Gm[b][c] += off[b] * o.dLde[c];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) grel[c] -= o.g_rel[c];
                s_rigid += o.s1; s_rot += o.s2; s_iso += o.s3;
            }
        }
        const int t0 = a.in_ptr[f], t1 = a.in_ptr[f + 1];
        for (int tb = t0 + sub; tb < t1; tb += TRK_SPLIT * TRK_UI) {
            int Ei[TRK_UI];
            F8 E[TRK_UI], M[TRK_UI];
#pragma unroll
            for (int u = 0; u < TRK_UI; ++u) {
                const int t = tb + TRK_SPLIT * u;
                Ei[u] = a.in_edge[t < t1 ? t : tb];
            }
#pragma unroll
            for (int u = 0; u < TRK_UI; ++u) {
                // e / K by multiplication (exact for every 31-bit e: magic = ceil(2^shift / K) < 2^32 and e K < 2^shift)
                const int i2 = (int)(((unsigned long long)(unsigned)Ei[u] * a.div_magic) >> a.div_shift);
                E[u] = ldg256(edge + 2 * (size_t)Ei[u]);
                M[u] = ldg256(node + 2 * (size_t)i2);
            }
#pragma unroll
            for (int u = 0; u < TRK_UI; ++u) {
                if (tb + TRK_SPLIT * u >= t1) break;
                const float x2[3] = {M[u].a.x, M[u].a.y, M[u].a.z};
                const Quat rel_2 = {M[u].b.x, M[u].b.y, M[u].b.z, M[u].b.w};
                const float n2 = trk_rsqrt(rel_2.w * rel_2.w + rel_2.x * rel_2.x + rel_2.y * rel_2.y + rel_2.z * rel_2.z);
                float R2[3][3];
                rot_from_unit(Quat{rel_2.w * n2, rel_2.x * n2, rel_2.y * n2, rel_2.z * n2}, R2);
                const float po[3] = {E[u].a.w, E[u].b.x, E[u].b.y};
                EdgeOut o;
                eval_edge(a, x2, R2, rel_2, xi, rel_i, E[u].a.y, E[u].a.z, po, o);
#pragma unroll
                for (int b = 0; b < 3; ++b) gx[b] += o.g_off[b];
#pragma unroll
                for (int c = 0; c < 4; ++c) grel[c] += o.g_rel[c];
            }
        }
    }
#pragma unroll
    for (int o = 1; o < TRK_SPLIT; o <<= 1) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            gx[b] += __shfl_xor_sync(0xffffffffu, gx[b], o);
#pragma unroll
            for (int c = 0; c < 3; ++c) Gm[b][c] += __shfl_xor_sync(0xffffffffu, Gm[b][c], o);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) grel[c] += __shfl_xor_sync(0xffffffffu, grel[c], o);
    }
    if (active && sub == 0) {
        const int gi = a.fg_index ? a.fg_index[f] : f;
        const float r = n_i.w, x = n_i.x, y = n_i.y, z = n_i.z;
        float gn[4];
        gn[0] = 2.f * (-z * Gm[0][1] + y * Gm[0][2] + z * Gm[1][0] - x * Gm[1][2] - y * Gm[2][0] + x * Gm[2][1]);
        gn[1] = 2.f * (y * Gm[0][1] + z * Gm[0][2] + y * Gm[1][0] - 2.f * x * Gm[1][1] - r * Gm[1][2] + z * Gm[2][0] + r * Gm[2][1] - 2.f * x * Gm[2][2]);
        gn[2] = 2.f * (-2.f * y * Gm[0][0] + x * Gm[0][1] + r * Gm[0][2] + x * Gm[1][0] + z * Gm[1][2] - r * Gm[2][0] + z * Gm[2][1] - 2.f * y * Gm[2][2]);
        gn[3] = 2.f * (-2.f * z * Gm[0][0] - r * Gm[0][1] + x * Gm[0][2] + r * Gm[1][0] - 2.f * z * Gm[1][1] + y * Gm[1][2] + x * Gm[2][0] + y * Gm[2][1]);
        const float dot = r * gn[0] + x * gn[1] + y * gn[2] + z * gn[3];
        grel[0] += (gn[0] - r * dot) * inv_n;
        grel[1] += (gn[1] - x * dot) * inv_n;
        grel[2] += (gn[2] - y * dot) * inv_n;
        grel[3] += (gn[3] - z * dot) * inv_n;
        if (xi[1] > 0.f) { gx[1] += a.c_floor; s_floor = xi[1]; }
        const Quat gq = qmul_bwd_a(Quat{grel[0], grel[1], grel[2], grel[3]}, load_q(a.prev_inv, f));
        a.grad_x[3 * (size_t)gi] = gx[0]; a.grad_x[3 * (size_t)gi + 1] = gx[1]; a.grad_x[3 * (size_t)gi + 2] = gx[2];
        *reinterpret_cast<float4 *>(a.grad_q + 4 * (size_t)gi) = make_float4(gq.w, gq.x, gq.y, gq.z);
    }
    float v[4] = {s_rigid, s_rot, s_iso, s_floor};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    if ((tg & 31) == 0)
#pragma unroll
        for (int c = 0; c < 4; ++c) red[tg >> 5][c] = v[c];
    __syncthreads();
    if (vb_ok && tg < 4) a.block_sums[5 * (size_t)vb + tg] = red[0][tg] + red[1][tg] + red[2][tg] + red[3][tg];
    if (vb_ok && tg == 4) a.block_sums[5 * (size_t)vb + 4] = 0.f;
    __syncthreads();   // red[] is reused by the next virtual block
    }
}

// packs the static per-edge tables into 32-byte records (once per timestep: prev_offset changes with the frame)
__global__ void gsd_track_pack_edges_kernel(long long n_edges, const int32_t *__restrict__ nbr, const float *__restrict__ w,
                                            const float *__restrict__ d0, const float *__restrict__ po, float4 *__restrict__ out) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    out[2 * e] = make_float4(__int_as_float(nbr[e]), w[e], d0[e], po[3 * e]);
    out[2 * e + 1] = make_float4(po[3 * e + 1], po[3 * e + 2], 0.f, 0.f);
}

extern "C" int gsd_track_pack_edges(int32_t Gf, int32_t K, const int32_t *neighbor_indices, const float *neighbor_weight,
                                    const float *neighbor_dist, const float *prev_offset, float *edge_records, void *stream) {
    if (Gf < 0 || K < 0 || (Gf > 0 && K > 0 && (!neighbor_indices || !neighbor_weight || !neighbor_dist || !prev_offset || !edge_records))) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    const long long n = (long long)Gf * K;
    if (n == 0) return GSD_OK;
    gsd_launch(gsd_track_pack_edges_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, n, neighbor_indices, neighbor_weight,
                                                                                                neighbor_dist, prev_offset, (float4 *)edge_records);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

__global__ void __launch_bounds__(128)
gsd_track_bg_kernel(TrackArgs a, int fg_blocks) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    __shared__ float red[4];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (b < a.Gb) {
        const int gi = a.bg_index[b];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float d = a.x[3 * (size_t)gi + c] - a.bg_x0[3 * (size_t)b + c];
            s += fabsf(d);
            a.grad_x[3 * (size_t)gi + c] = a.c_bg * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
        const Quat qb = load_rot(a, gi);
        const float qv[4] = {qb.w, qb.x, qb.y, qb.z};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float d = qv[c] - a.bg_q0[4 * (size_t)b + c];
            s += fabsf(d);
            a.grad_q[4 * (size_t)gi + c] = a.c_bg * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        size_t row = (size_t)fg_blocks + blockIdx.x;
        a.block_sums[5 * row + 0] = 0.f; a.block_sums[5 * row + 1] = 0.f; a.block_sums[5 * row + 2] = 0.f;
        a.block_sums[5 * row + 3] = 0.f;
        a.block_sums[5 * row + 4] = red[0] + red[1] + red[2] + red[3];
    }
}

// losses[0..4] = rigid, rot, iso, floor, bg (unweighted means); losses[5] = weighted total
__global__ void gsd_track_finish_kernel(int nrows, const float *__restrict__ block_sums, float inv_e, float inv_f,
                                        float inv_b, float w_rigid, float w_rot, float w_iso, float w_floor, float w_bg,
                                        float *__restrict__ losses) {
    gsd_pdl_wait();
    gsd_pdl_launch();
    __shared__ double r[5][128];
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < nrows; i += 128)
        for (int c = 0; c < 5; ++c) v[c] += block_sums[5 * (size_t)i + c];
    for (int c = 0; c < 5; ++c) r[c][threadIdx.x] = v[c];
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if (threadIdx.x < s)
            for (int c = 0; c < 5; ++c) r[c][threadIdx.x] += r[c][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float l0 = (float)(r[0][0] * inv_e), l1 = (float)(r[1][0] * inv_e), l2 = (float)(r[2][0] * inv_e);
        float l3 = (float)(r[3][0] * inv_f), l4 = (float)(r[4][0] * inv_b);
        losses[0] = l0; losses[1] = l1; losses[2] = l2; losses[3] = l3; losses[4] = l4;
        losses[5] = w_rigid * l0 + w_rot * l1 + w_iso * l2 + w_floor * l3 + w_bg * l4;
    }
}

extern "C" int gsd_track_losses_workspace_bytes(int32_t Gf, int32_t Gb, size_t *bytes) {
    if (Gf < 0 || Gb < 0 || !bytes) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    size_t rows = ((size_t)Gf * TRK_SPLIT + 127) / 128 + (size_t)(Gb + 127) / 128 + 1;
    *bytes = gsd_align_up(rows * 5 * 4) + gsd_align_up((size_t)(Gf > 0 ? Gf : 1) * 32);
    return GSD_OK;
}

extern "C" int gsd_track_losses_fwd_bwd(const GsdTrackLosses *t, void *stream) {
    if (!t || t->G < 0 || t->Gf < 0 || t->Gb < 0 || t->K < 0 || !t->ws || !t->losses || !t->grad_means3D || !t->grad_rotations) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    if (t->Gf > 0 && (!t->means3D || !t->rotations || !t->prev_inv_rot || (t->K > 0 && (!t->neighbor_indices || !t->neighbor_weight ||
        !t->neighbor_dist || !t->prev_offset || !t->in_ptr || !t->in_edge)))) {
        gsd_set_error("null foreground input");
        return GSD_ERR_INVALID;
    }
    if (t->Gb > 0 && (!t->bg_index || !t->init_bg_pts || !t->init_bg_rot)) { gsd_set_error("null background input"); return GSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if ((long long)t->Gf + t->Gb != t->G) {   // foreground + background cover every Gaussian (is_fg / ~is_fg): each row is written once below
        GSD_CUDA_CHECK(cudaMemsetAsync(t->grad_means3D, 0, (size_t)t->G * 3 * 4, st));
        GSD_CUDA_CHECK(cudaMemsetAsync(t->grad_rotations, 0, (size_t)t->G * 4 * 4, st));
    }
    TrackArgs a;
    a.Gf = t->Gf; a.K = t->K; a.Gb = t->Gb;
    a.unnorm = t->rotations_unnormalized;
    {
        int lg = 0;
        while ((1ll << lg) < (long long)(t->K > 0 ? t->K : 1)) ++lg;
        a.div_shift = 31 + lg;
        a.div_magic = (unsigned)(((1ull << a.div_shift) + (unsigned long long)(t->K > 0 ? t->K : 1) - 1) / (unsigned long long)(t->K > 0 ? t->K : 1));
    }
    a.x = t->means3D; a.q = t->rotations; a.fg_index = t->fg_index; a.prev_inv = t->prev_inv_rot;
    a.nbr = t->neighbor_indices; a.nbr_w = t->neighbor_weight; a.nbr_d = t->neighbor_dist; a.prev_off = t->prev_offset;
    a.in_ptr = t->in_ptr; a.in_edge = t->in_edge;
    a.bg_index = t->bg_index; a.bg_x0 = t->init_bg_pts; a.bg_q0 = t->init_bg_rot;
    const double ne = (double)t->Gf * (double)(t->K > 0 ? t->K : 1);
    a.c_rigid = t->Gf > 0 ? (float)(t->w_rigid / ne) : 0.f;
    a.c_rot = t->Gf > 0 ? (float)(t->w_rot / ne) : 0.f;
    a.c_iso = t->Gf > 0 ? (float)(t->w_iso / ne) : 0.f;
    a.c_floor = t->Gf > 0 ? t->w_floor / (float)t->Gf : 0.f;
    a.c_bg = t->Gb > 0 ? t->w_bg / (float)t->Gb : 0.f;
    a.grad_x = t->grad_means3D; a.grad_q = t->grad_rotations;
    a.block_sums = (float *)t->ws;
    const int fgb = (int)(((size_t)t->Gf * TRK_SPLIT + 127) / 128), bgb = (t->Gb + 127) / 128;
    if (fgb > 0 && t->edge_records && t->K > 0) {
        const size_t rows = (size_t)fgb + bgb + 1;
        float4 *node = (float4 *)((char *)t->ws + gsd_align_up(rows * 5 * 4));
        gsd_launch(gsd_track_node_prep_kernel, dim3((t->Gf + 255) / 256), dim3(256), 0, st, a, node);
        GSD_LAUNCH_CHECK();
        // Optional resident-CTA cap (a grid of SMs x cap CTAs walking contiguous runs of the work).  Measured inside the tracking
        // iteration, where this kernel runs on a side branch beside the render branch: every cap is slower than the plain grid
        // (cap 1 / 2 / 4 / 6 / 8: +90 / +55 / +10 / +6 / +1 us per iteration) — the kernel's cost to the iteration is its own
        // machine time, not the placement of the other branch's CTAs.  Kept as a tuning switch, off by default.
        {   // shared-memory carveout of the SMs this kernel runs on (see below); GSD_PRIORS_CARVEOUT: tuning override, -1 = driver default
            static int carve_set = -2;
            const char *ce = getenv("GSD_PRIORS_CARVEOUT");
            const int carve = ce ? atoi(ce) : GSD_PRIORS_CARVEOUT_DEFAULT;
            if (carve != carve_set) {
                GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_track_fg_packed_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                carve_set = carve;
            }
        }
        const char *cap_env = getenv("GSD_PRIORS_CTAS_PER_SM");   // tuning override (tools/step_ablate.py)
        const int cap = cap_env ? atoi(cap_env) : GSD_PRIORS_CTAS_PER_SM_DEFAULT;
        // The work is launched in `pieces` consecutive kernels of at most one wave of CTAs each: a grid with a backlog of pending
        // CTAs refills every slot a retiring CTA frees, so the render branch's 1 024-thread binning CTAs (same graph, other branch,
        // higher priority) could not be placed until the backlog was gone; between two pieces whole SMs drain and go to them.
        const char *sp_env = getenv("GSD_PRIORS_PIECES");
        const int pieces_req = sp_env ? atoi(sp_env) : GSD_PRIORS_PIECES_DEFAULT;
        const int pieces = pieces_req > 1 ? pieces_req : 1;
        const int ng = TRK_THREADS / 128;
        const int per_piece = (fgb + pieces - 1) / pieces;
        for (int pc = 0; pc < pieces; ++pc) {
            const int vb0 = pc * per_piece, nvb = (fgb - vb0 < per_piece) ? fgb - vb0 : per_piece;
            if (nvb <= 0) break;
            const int ctas = (nvb + ng - 1) / ng;
            const int grid = cap > 0 ? (ctas < 148 * cap ? ctas : 148 * cap) : ctas;
            gsd_launch(gsd_track_fg_packed_kernel, dim3(grid), dim3(TRK_THREADS), 0, st, a, node, (const float4 *)t->edge_records, vb0, nvb);
            GSD_LAUNCH_CHECK();
        }
        GSD_LAUNCH_CHECK();
    } else if (fgb > 0) {
        gsd_launch(gsd_track_fg_kernel, dim3(fgb), dim3(128), 0, st, a);
        GSD_LAUNCH_CHECK();
    }
    if (bgb > 0) { gsd_launch(gsd_track_bg_kernel, dim3(bgb), dim3(128), 0, st, a, fgb); GSD_LAUNCH_CHECK(); }
    gsd_launch(gsd_track_finish_kernel, dim3(1), dim3(128), 0, st, fgb + bgb, a.block_sums, t->Gf > 0 ? (float)(1.0 / ne) : 0.f,
                                               t->Gf > 0 ? 1.f / t->Gf : 0.f, t->Gb > 0 ? 1.f / t->Gb : 0.f, t->w_rigid,
                                               t->w_rot, t->w_iso, t->w_floor, t->w_bg, t->losses);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
