// gnn_bwd.cu — backward of the fused GNN kernels (GNN training, SURVEY.md §8f row 3: /root/reference/src/train.py:183-211
// back-propagates a 5-step unroll through DynamicsPredictor.forward, /root/reference/src/gnn/model.py:112-246, whose one-hot
// bmm gathers/scatters autograd differentiates as dense [E,N] matmuls).  Here:
//   aggregate backward  (forward: agg[r] = sum_e ReLU(A[e] + P[r,0:F] + P[send(e),F:2F]), gnn.cu)
//       gA[e]      = g[r] where the pre-activation is positive, else 0          (recomputed, nothing saved by the forward)
//       gP[r,0:F]  = sum over the receiver row of gA[e]                          (same pass, receiver CSR)
//       gP[s,F:2F] = sum over the edges SENT by s of gA[e]                       (second pass over the transposed CSR)
//   edge-input backward (forward: rel[e, off+3h+c] = state[h,recv,c] - state[h,send,c])
//       g_state[h,n,c] = sum_{e: recv=n} g_rel[e,.] - sum_{e: send=n} g_rel[e,.]
// No atomics: every output row is owned by one warp (or, for the tool rows every object is connected to, by GNN_SPLIT warps
// whose partials are summed in fixed order), so gradients are bit-reproducible.
#include "common.cuh"

#define GNN_SPLIT 128

// A work item is (row, split): light rows [0, n_light) of every batch element are one item each, the n_heavy last rows are
// GNN_SPLIT items each.  MODE 0: receiver pass (writes gA, sums into gP[:,0:F]).  MODE 1: sender pass (gathers gA rows).
template <int V, int MODE>
__global__ void __launch_bounds__(128)
gsd_gnn_aggregate_bwd_kernel(int B, int N, int cap, int n_light, int n_heavy, const int32_t *__restrict__ ptr,
                             const int32_t *__restrict__ idx /* MODE 0: senders; MODE 1: edge order by sender */,
                             const float4 *__restrict__ A, const float4 *__restrict__ P, const float4 *__restrict__ g,
                             float4 *__restrict__ gA, float4 *__restrict__ gP, float4 *__restrict__ partial) {
    constexpr int F4 = 32 * V;
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int per_b = n_light + n_heavy * GNN_SPLIT;
    if (item >= (long long)B * per_b) return;
    const int b = (int)(item / per_b), it = (int)(item % per_b);
    int r, split = -1;
    if (it < n_light) r = it;
    else { r = n_light + (it - n_light) / GNN_SPLIT; split = (it - n_light) % GNN_SPLIT; }
    const int32_t *rp = ptr + (size_t)b * (N + 1);
    int e0 = min(rp[r], cap), e1 = min(rp[r + 1], cap);
    if (split >= 0) {
        const int per = (e1 - e0 + GNN_SPLIT - 1) / GNN_SPLIT;
        e0 = e0 + split * per;
        e1 = min(e1, e0 + per);
    }
    const size_t node = (size_t)b * N + r;
    float4 acc[V], gr[V], pr[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 0) {
            gr[v] = g[node * F4 + v * 32 + lane];
            pr[v] = P[node * 2 * F4 + v * 32 + lane];
        }
    }
    for (int e = e0; e < e1; ++e) {
        const size_t ge = (size_t)b * cap + e;
        if (MODE == 0) {
            const size_t snode = (size_t)b * N + idx[ge];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 a4 = A[ge * F4 + v * 32 + lane];
                const float4 s4 = P[snode * 2 * F4 + F4 + v * 32 + lane];
                float4 o;
                o.x = (a4.x + pr[v].x + s4.x > 0.f) ? gr[v].x : 0.f;
                o.y = (a4.y + pr[v].y + s4.y > 0.f) ? gr[v].y : 0.f;
                o.z = (a4.z + pr[v].z + s4.z > 0.f) ? gr[v].z : 0.f;
                o.w = (a4.w + pr[v].w + s4.w > 0.f) ? gr[v].w : 0.f;
                gA[ge * F4 + v * 32 + lane] = o;
                acc[v].x += o.x; acc[v].y += o.y; acc[v].z += o.z; acc[v].w += o.w;
            }
        } else {
            const size_t src = (size_t)b * cap + idx[ge];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 o = gA[src * F4 + v * 32 + lane];
                acc[v].x += o.x; acc[v].y += o.y; acc[v].z += o.z; acc[v].w += o.w;
            }
        }
    }
    if (split < 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) gP[node * 2 * F4 + MODE * F4 + v * 32 + lane] = acc[v];
    } else {
        const size_t hrow = (size_t)b * n_heavy + (r - n_light);
#pragma unroll
        for (int v = 0; v < V; ++v) partial[(hrow * GNN_SPLIT + split) * F4 + v * 32 + lane] = acc[v];
    }
}

// fixed-order sum of the GNN_SPLIT partials of a heavy row into out[row, off4 : off4 + F4] (row stride stride4, in float4)
__global__ void __launch_bounds__(128)
gsd_gnn_split_finish_kernel(int N, int n_light, int n_heavy, int F4, int stride4, int off4, const float4 *__restrict__ partial,
                            float4 *__restrict__ out) {
    const int hrow = blockIdx.x, col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= F4) return;
    const int b = hrow / n_heavy, r = n_light + hrow % n_heavy;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int s = 0; s < GNN_SPLIT; ++s) {
        const float4 p = partial[((size_t)hrow * GNN_SPLIT + s) * F4 + col];
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    out[((size_t)b * N + r) * stride4 + off4 + col] = acc;
}

extern "C" int gsd_gnn_aggregate_bwd_workspace_bytes(int32_t B, int32_t n_heavy, int32_t F, size_t *bytes) {
    if (B <= 0 || n_heavy < 0 || F <= 0 || !bytes) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    *bytes = gsd_align_up((size_t)B * (n_heavy > 0 ? n_heavy : 1) * GNN_SPLIT * F * 4);
    return GSD_OK;
}

template <int V>
static int aggregate_bwd_launch(int B, int N, int cap, int n_heavy, const int32_t *row_ptr, const int32_t *senders,
                                const int32_t *col_ptr, const int32_t *order, const float *A, const float *P, const float *g,
                                void *ws, float *gA, float *gP, cudaStream_t st) {
    const int n_light = N - n_heavy;
    const long long items = (long long)B * (n_light + (long long)n_heavy * GNN_SPLIT);
    const unsigned grid = (unsigned)((items + 3) / 4);
    const int F4 = 32 * V;
    gsd_gnn_aggregate_bwd_kernel<V, 0><<<grid, 128, 0, st>>>(B, N, cap, n_light, n_heavy, row_ptr, senders, (const float4 *)A,
                                                              (const float4 *)P, (const float4 *)g, (float4 *)gA, (float4 *)gP,
                                                              (float4 *)ws);
    GSD_LAUNCH_CHECK();
    if (n_heavy > 0) {
        gsd_gnn_split_finish_kernel<<<dim3(B * n_heavy, (F4 + 127) / 128), 128, 0, st>>>(N, n_light, n_heavy, F4, 2 * F4, 0,
                                                                                         (const float4 *)ws, (float4 *)gP);
        GSD_LAUNCH_CHECK();
    }
    gsd_gnn_aggregate_bwd_kernel<V, 1><<<grid, 128, 0, st>>>(B, N, cap, n_light, n_heavy, col_ptr, order, nullptr, nullptr, nullptr,
                                                              (float4 *)gA, (float4 *)gP, (float4 *)ws);
    GSD_LAUNCH_CHECK();
    if (n_heavy > 0) {
        gsd_gnn_split_finish_kernel<<<dim3(B * n_heavy, (F4 + 127) / 128), 128, 0, st>>>(N, n_light, n_heavy, F4, 2 * F4, F4,
                                                                                         (const float4 *)ws, (float4 *)gP);
        GSD_LAUNCH_CHECK();
    }
    return GSD_OK;
}

// row_ptr/senders: receiver CSR of the forward.  col_ptr [B,N+1] / order [B,capacity]: the same edges grouped by SENDER
// (order = edge ids of element b sorted by sender, stable).  gA [B*capacity,F], gP [B*N,2F] are fully written for valid edges /
// all nodes; gA rows of unused edge slots are left untouched (callers zero or ignore them).
extern "C" int gsd_gnn_aggregate_bwd(int32_t B, int32_t N, int32_t capacity, int32_t F, int32_t n_heavy, const int32_t *row_ptr,
                                     const int32_t *senders, const int32_t *col_ptr, const int32_t *order, const float *A,
                                     const float *P, const float *g_agg, void *ws, float *gA, float *gP, void *stream) {
    if (B <= 0 || N <= 0 || capacity < 0 || n_heavy < 0 || n_heavy > N || !row_ptr || !senders || !col_ptr || !order || !A || !P ||
        !g_agg || !gA || !gP || (n_heavy > 0 && !ws)) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    if (F % 128 != 0 || F > 512) { gsd_set_error("feature width %d must be a multiple of 128 and <= 512", F); return GSD_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (F / 128) {
    case 1: return aggregate_bwd_launch<1>(B, N, capacity, n_heavy, row_ptr, senders, col_ptr, order, A, P, g_agg, ws, gA, gP, st);
    case 2: return aggregate_bwd_launch<2>(B, N, capacity, n_heavy, row_ptr, senders, col_ptr, order, A, P, g_agg, ws, gA, gP, st);
    case 3: return aggregate_bwd_launch<3>(B, N, capacity, n_heavy, row_ptr, senders, col_ptr, order, A, P, g_agg, ws, gA, gP, st);
    default: return aggregate_bwd_launch<4>(B, N, capacity, n_heavy, row_ptr, senders, col_ptr, order, A, P, g_agg, ws, gA, gP, st);
    }
}

// ------------------------------------------------------------------------------------------------------
// edge-input backward: one warp per node, lanes stride over the node's in-edges (+) and out-edges (-)
// ------------------------------------------------------------------------------------------------------
#define GNN_MAX_HIS 8

__global__ void __launch_bounds__(128)
gsd_gnn_edge_inputs_bwd_kernel(int B, int N, int cap, int n_his, int width, int off, const int32_t *__restrict__ row_ptr,
                               const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ order,
                               const float *__restrict__ g_rel, float *__restrict__ g_state) {
    const int lane = threadIdx.x & 31;
    const long long node = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (node >= (long long)B * N) return;
    const int b = (int)(node / N), n = (int)(node % N);
    const int nv = 3 * n_his;
    float acc[3 * GNN_MAX_HIS];
#pragma unroll
    for (int v = 0; v < 3 * GNN_MAX_HIS; ++v) acc[v] = 0.f;
    const int32_t *rp = row_ptr + (size_t)b * (N + 1), *cp = col_ptr + (size_t)b * (N + 1);
    for (int e = min(rp[n], cap) + lane; e < min(rp[n + 1], cap); e += 32) {
        const float *gr = g_rel + ((size_t)b * cap + e) * width + off;
#pragma unroll
        for (int v = 0; v < 3 * GNN_MAX_HIS; ++v)
            if (v < nv) acc[v] += gr[v];
    }
    for (int k = min(cp[n], cap) + lane; k < min(cp[n + 1], cap); k += 32) {
        const float *gr = g_rel + ((size_t)b * cap + order[(size_t)b * cap + k]) * width + off;
#pragma unroll
        for (int v = 0; v < 3 * GNN_MAX_HIS; ++v)
            if (v < nv) acc[v] -= gr[v];
    }
#pragma unroll
    for (int v = 0; v < 3 * GNN_MAX_HIS; ++v) {
        if (v < nv) {
            float a = acc[v];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);   // fixed butterfly: deterministic
            if (lane == 0) g_state[(((size_t)b * n_his + v / 3) * N + n) * 3 + v % 3] = a;
        }
    }
}

// g_rel [B,capacity,width] (gradient of gsd_gnn_edge_inputs' output; columns off .. off+3*n_his are the position differences)
// -> g_state [B,n_his,N,3]
extern "C" int gsd_gnn_edge_inputs_bwd(int32_t B, int32_t N, int32_t capacity, int32_t n_his, int32_t width, int32_t offset,
                                       const int32_t *row_ptr, const int32_t *col_ptr, const int32_t *order, const float *g_rel,
                                       float *g_state, void *stream) {
    if (B <= 0 || N <= 0 || capacity < 0 || n_his <= 0 || n_his > GNN_MAX_HIS || width < offset + 3 * n_his || offset < 0 || !row_ptr ||
        !col_ptr || !order || !g_rel || !g_state) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    const long long nodes = (long long)B * N;
    gsd_gnn_edge_inputs_bwd_kernel<<<(unsigned)((nodes + 3) / 4), 128, 0, (cudaStream_t)stream>>>(B, N, capacity, n_his, width, offset,
                                                                                                    row_ptr, col_ptr, order, g_rel,
                                                                                                    g_state);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
