// gnn.cu — graph kernels of the GNN particle-dynamics step (path B), sm_100a.
//
//   gsd_gnn_build_edges   radius AND top-k adjacency + tool edges -> CSR-by-receiver index lists (never N x N one-hots)
//                         replaces construct_edges_from_states, /root/reference/src/data/dataset.py:88-147
//   gsd_gnn_edge_inputs   per-edge relation features [attrs_r | attrs_s | group diff | pos_r - pos_s over history]
//                         replaces the six one-hot bmm of /root/reference/src/gnn/model.py:164-199
//   gsd_gnn_aggregate     agg[r] = sum_{e: recv(e)=r} ReLU(A[e] + Pr[r] + Ps[send(e)])   (edge MLP epilogue + segment reduce)
//                         replaces Rr.bmm / Rs.bmm / cat / Linear epilogue / Rr_t.bmm of model.py:212-229 after splitting
//                         W_rel [enc_e | h_r | h_s] = W1 enc_e + W2 h_r + W3 h_s  (the dense parts stay on cuBLAS)
//   gsd_fps / gsd_fps_radius   farthest point sampling (dgl.geometry.farthest_point_sampler, data/utils.py:50-65)
//
// All of these are gather / scatter-shaped, HBM/L2-bound integer+fp32 work: warp-per-row with float4 lanes, no tensor cores.
// Algorithmic bytes of gsd_gnn_aggregate per call: 4*E*F (A) + 8*E (indices) + 12*N*F (Pr, Ps, agg)  (SURVEY.md §8d).
#include "common.cuh"

#define GNN_MAX_SMEM_NODES 4096
#define GNN_CAND 256 // within-radius candidates per row kept in shared memory

__device__ __forceinline__ float dist2_rn(float ax, float ay, float az, float bx, float by, float bz) {
    // ((dx^2 + dy^2) + dz^2) without FMA contraction, the order torch.sum(s_diff ** 2, -1) uses for 3 elements
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ------------------------------------------------------------------------------------------------------
// edge construction.  One warp per receiver row; tools are the last n_tool nodes (dataset.py:120-124).
//   pass 0: adjacency bit rows + row counts;  (host launches a 1-block scan);  pass 1: expand bits to (recv, send) lists
// ------------------------------------------------------------------------------------------------------
struct EdgeArgs {
    int B, N, n_tool, topk, connect_all;
    const float *states;     // [B,N,3]
    const uint8_t *mask;     // [B,N] valid particle
    const uint8_t *tool_mask;// [B,N]
    const float *thresh;     // [B] adj_thresh (not squared) or null -> thresh_sq_scalar
    float thresh_sq_scalar;
    uint32_t *bits;          // [B*N][words]
    int words;
    int32_t *row_count;      // [B*N]
};

// ---- general adjacency kernel (any top-k) ----
__global__ void __launch_bounds__(256)
gsd_gnn_adjacency_general_kernel(EdgeArgs a) {
    // shared: positions [N*3] | per warp: candidates [GNN_CAND] u64 | per warp: row distances [N] | per warp: selected bits [words]
    // | node flags [N] (bit 0 valid, bit 1 tool)
    extern __shared__ float spos[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long *cand_all = reinterpret_cast<unsigned long long *>(spos + ((a.N * 3 + 1) & ~1));
    float *sd_all = reinterpret_cast<float *>(cand_all + (size_t)warps * GNN_CAND);
    uint32_t *sel_all = reinterpret_cast<uint32_t *>(sd_all + (size_t)warps * a.N);
    uint8_t *sflag = reinterpret_cast<uint8_t *>(sel_all + (size_t)warps * a.words);
    unsigned long long *cand = cand_all + (size_t)warp * GNN_CAND;
    float *sd = sd_all + (size_t)warp * a.N;
    uint32_t *selbits = sel_all + (size_t)warp * a.words;
    const int b = blockIdx.y;
    const float *st = a.states + (size_t)b * a.N * 3;
    const uint8_t *mk = a.mask + (size_t)b * a.N, *tm = a.tool_mask + (size_t)b * a.N;
#pragma unroll 4
    for (int i = threadIdx.x; i < a.N * 3; i += blockDim.x) spos[i] = st[i];
#pragma unroll 4
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) sflag[i] = (mk[i] ? 1 : 0) | (tm[i] ? 2 : 0);
    __syncthreads();
    const int n_obj = a.N - a.n_tool;
    const float thr = a.thresh ? __fmul_rn(a.thresh[b], a.thresh[b]) : a.thresh_sq_scalar;
    for (int r = blockIdx.x * warps + warp; r < a.N; r += gridDim.x * warps) {
    const float rx = spos[3 * r], ry = spos[3 * r + 1], rz = spos[3 * r + 2];
    const bool r_valid = (sflag[r] & 1) != 0, r_tool = (sflag[r] & 2) != 0;
    uint32_t *row = a.bits + ((size_t)b * a.N + r) * a.words;
    __syncwarp(); // previous row's readers of sd / selbits / cand are done

    // ---- distances of this row, cached per warp (masked pairs -> 1e10 like the reference's dis[mask] = 1e10)
    int cnt = 0; // object columns within the radius
    for (int c = lane; c < a.N; c += 32) {
        float d = dist2_rn(rx, ry, rz, spos[3 * c], spos[3 * c + 1], spos[3 * c + 2]);
        const int fc = sflag[c];
        if (!(r_valid && (fc & 1)) || (r_tool && (fc & 2))) d = 1e10f;
        sd[c] = d;
        cnt += (c < n_obj && __fsub_rn(d, thr) < 0.f) ? 1 : 0;
    }
    for (int w = lane; w < a.words; w += 32) selbits[w] = 0u;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __syncwarp();

    // ---- object block: adj = (d < r^2) AND (column among the k nearest of the row).  Every column nearer than a
    // within-radius column is itself within the radius, so only within-radius columns need ranking; and once at least k
    // columns lie strictly inside a smaller bound t, every column at or beyond t is outside the k nearest.  So: shrink t by
    // bisection until a small candidate set (>= k columns) lies inside it, compact that set into shared memory, rank the
    // candidates against each other (ties -> lowest index) and set the bits of those ranked below k.  If bisection cannot
    // separate (hundreds of equal distances, or k larger than the buffer) fall back to k rounds of warp arg-min.
    if (r < n_obj && cnt > 0) {
        const int k_eff = min(a.topk, a.N);
        const int want = min(GNN_CAND, max(64, 2 * k_eff)); // candidate-set size the bisection aims below
        float t = thr;
        bool exact_thr = true, fallback = false;
        if (cnt > want) {
            fallback = true;
            if (k_eff <= want) {
                float lo = 0.f, hi = thr;
                for (int it = 0; it < 48; ++it) {
                    const float mid = 0.5f * (lo + hi);
                    if (!(mid > lo && mid < hi)) break;
                    int c_mid = 0;
                    for (int c = lane; c < n_obj; c += 32) c_mid += sd[c] < mid ? 1 : 0;
                    c_mid = __reduce_add_sync(0xffffffffu, c_mid);
                    if (c_mid > want) hi = mid;
                    else if (c_mid < k_eff) lo = mid;
                    else { t = mid; exact_thr = false; fallback = false; break; }
                }
            }
        }
        if (!fallback) {
            int n_cand = 0;
            for (int c0 = 0; c0 < n_obj; c0 += 32) {
                const int c = c0 + lane;
                const float d = c < n_obj ? sd[c] : 3.0e38f;
                const bool in = exact_thr ? (__fsub_rn(d, thr) < 0.f) : (d < t);
                const unsigned m = __ballot_sync(0xffffffffu, in);
                const int pos = n_cand + __popc(m & ((1u << lane) - 1u));
                if (in) cand[pos] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)c; // d >= 0: bit order = value order
                n_cand += __popc(m);
            }
            __syncwarp();
            for (int q = lane; q < n_cand; q += 32) {
                const unsigned long long me = cand[q];
                int rank = 0;
                for (int j = 0; j < n_cand; ++j) rank += cand[j] < me ? 1 : 0;
                if (rank < k_eff) {
                    const unsigned c = (unsigned)(me & 0xffffffffull);
                    atomicOr(&selbits[c >> 5], 1u << (c & 31));
                }
            }
        } else {
            for (int it = 0; it < min(k_eff, n_obj); ++it) {
                float best = 3.0e38f;
                int best_c = 0x7fffffff;
                for (int c = lane; c < n_obj; c += 32) { // column c is always visited by lane c % 32: selbits[c/32] bit lane
                    if ((selbits[c >> 5] >> lane) & 1u) continue;
                    const float d = sd[c];
                    if (d < best) { best = d; best_c = c; }
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
                    if (ob < best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
                }
                if (best_c == 0x7fffffff) break;
                if ((best_c & 31) == lane) selbits[best_c >> 5] |= 1u << lane;
                __syncwarp();
            }
        }
        __syncwarp();
    }

    // ---- adjacency bits
    int count = 0;
    for (int w = 0; w < a.words; ++w) {
        const int c = w * 32 + lane;
        bool on = false;
        if (c < a.N) {
            const bool c_valid = (sflag[c] & 1) != 0, c_tool = (sflag[c] & 2) != 0;
            on = __fsub_rn(sd[c], thr) < 0.f;
            if (r < n_obj && c < n_obj) on = on && ((selbits[w] >> lane) & 1u);
            if (a.connect_all) {
                if (r_tool && c_valid) on = true;
                if (c_tool && r_valid) on = true;
                if (r_tool && c_tool) on = false;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) row[w] = m;
        count += __popc(m);
    }
    if (lane == 0) a.row_count[(size_t)b * a.N + r] = count;
    } // rows
}

// ---- fused edge builder for top-k <= KL (every configuration of the reference: topk 5 / 6 / 10): ONE launch -----------------------
// A CTA owns a contiguous block of receiver rows; one warp per row, ONE pass over the columns: lane l visits columns c = 32 i + l,
// so the warp ballot of "within the radius" IS bit-word i of the row.  The (distance, column) keys of the within-radius object
// columns are compacted into a small per-warp buffer (ballot prefix); when it fills, and at the end of the row, every lane inserts
// ITS share of the buffer into a sorted register list of the KL smallest keys it has seen (all lanes busy; inserting inside the
// column loop ran the ~50-instruction insertion whenever ANY lane had a hit, i.e. in ~90 % of the iterations).  The row's k nearest
// are then k rounds of a warp arg-min over the lanes' list heads (two redux.sync each); the final bit-words are word-wise logic on
// the radius words, the selected words and the per-graph valid / tool words and STAY IN SHARED MEMORY.  The CTA publishes its
// edge count, sums the counts of the CTAs before it (decoupled look-back: all CTAs publish at about the same time, each reads its
// predecessors' words in parallel), and expands its rows' bit-words straight into the receiver / sender lists: no N x N bit matrix
// in HBM, no scan launch, no expand launch.
#define GNN_KEYBUF 128   // keys per warp between two flushes
__device__ __forceinline__ int gnn_ld_acquire(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gnn_st_release(int32_t *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int KL>
__global__ void __launch_bounds__(256)
gsd_gnn_edges_fused_kernel(EdgeArgs a, int rows_per_cta, int cap, int32_t *__restrict__ row_ptr, int32_t *__restrict__ n_edges,
                           int32_t *__restrict__ recv, int32_t *__restrict__ send, int32_t *agg /* [B][gridDim.x], zeroed */) {
    // shared: per warp key buffer | positions [3 * 32 words] | valid words | tool words | per warp selected words |
    //         row words [rows_per_cta][words] | row counts, row offsets [rows_per_cta] | flags [32 words]
    extern __shared__ unsigned long long skeys[];
    __shared__ int s_red[8], s_prefix;
    gsd_pdl_trigger();
    gsd_pdl_wait();
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npad = a.words * 32;
    float *spos = reinterpret_cast<float *>(skeys + (size_t)warps * GNN_KEYBUF);
    uint32_t *validw = reinterpret_cast<uint32_t *>(spos + 3 * npad);
    uint32_t *toolw = validw + a.words;
    uint32_t *selw = toolw + a.words + (size_t)warp * a.words;
    uint32_t *roww = toolw + a.words + (size_t)warps * a.words;
    int *rcount = reinterpret_cast<int *>(roww + (size_t)rows_per_cta * a.words);
    int *roff = rcount + rows_per_cta;
    uint8_t *sflag = reinterpret_cast<uint8_t *>(roff + rows_per_cta);
    unsigned long long *keys = skeys + (size_t)warp * GNN_KEYBUF;
    const int b = blockIdx.y;
    const float *st = a.states + (size_t)b * a.N * 3;
    const uint8_t *mk = a.mask + (size_t)b * a.N, *tm = a.tool_mask + (size_t)b * a.N;
    // staging: positions by 4-byte cp.async (all in flight at once; padding columns far away), flags by plain loads, then the
    // valid / tool bit-words
    for (int i = threadIdx.x; i < 3 * npad; i += blockDim.x) {
        if (i < a.N * 3) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(spos + i)), "l"(st + i) : "memory");
        else spos[i] = 1e18f;
    }
#pragma unroll 8
    for (int i = threadIdx.x; i < npad; i += blockDim.x) sflag[i] = i < a.N ? ((mk[i] ? 1 : 0) | (tm[i] ? 2 : 0)) : 0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int f = sflag[i];
        const unsigned v = __ballot_sync(0xffffffffu, f & 1), t = __ballot_sync(0xffffffffu, f & 2);
        if (lane == 0) { validw[i >> 5] = v; toolw[i >> 5] = t; }
    }
    __syncthreads();
    const int n_obj = a.N - a.n_tool;
    const float thr = a.thresh ? __fmul_rn(a.thresh[b], a.thresh[b]) : a.thresh_sq_scalar;
    const int k_eff = min(a.topk, a.N);   // the launcher guarantees k_eff <= KL
    const unsigned lt_mask = (1u << lane) - 1u;
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(row0 + rows_per_cta, a.N);
    for (int r = row0 + warp; r < row1; r += warps) {
        const float rx = spos[3 * r], ry = spos[3 * r + 1], rz = spos[3 * r + 2];
        const bool r_valid = (validw[r >> 5] >> (r & 31)) & 1u, r_tool = (toolw[r >> 5] >> (r & 31)) & 1u, r_obj = r < n_obj;
        uint32_t *rw = roww + (size_t)(r - row0) * a.words;
        __syncwarp();   // the previous row's readers of selw / keys are done
        for (int w = lane; w < a.words; w += 32) selw[w] = 0u;
        unsigned long long best[KL];   // ascending; key = distance bits << 32 | column (d >= 0: bit order = value order, ties -> lowest column)
#pragma unroll
        for (int j = 0; j < KL; ++j) best[j] = ~0ull;
        int n_keys = 0;
        auto flush = [&]() {
            __syncwarp();
            for (int q = lane; q < n_keys; q += 32) {
                const unsigned long long key = keys[q];
                if (key < best[KL - 1]) {
                    best[KL - 1] = key;
#pragma unroll
                    for (int j = KL - 1; j >= 1; --j) {
                        const unsigned long long lo = best[j - 1], hi = best[j];
                        const bool sw = hi < lo;
                        best[j - 1] = sw ? hi : lo;
                        best[j] = sw ? lo : hi;
                    }
                }
            }
            n_keys = 0;
            __syncwarp();
        };
        for (int w = 0; w < a.words; ++w) {
            const int c = w * 32 + lane;
            // masked pairs: invalid row or column, tool-tool  (dis[mask] = 1e10, dataset.py:106-111)
            const unsigned bad = r_valid ? (~validw[w] | (r_tool ? toolw[w] : 0u)) : 0xffffffffu;
            float d = dist2_rn(rx, ry, rz, spos[3 * c], spos[3 * c + 1], spos[3 * c + 2]);
            if ((bad >> lane) & 1u) d = 1e10f;
            const bool in = __fsub_rn(d, thr) < 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (lane == 0) rw[w] = m;
            if (r_obj && m != 0u && w * 32 < n_obj) {
                const bool cand = in && c < n_obj;
                const unsigned mc = __ballot_sync(0xffffffffu, cand);
                if (cand) keys[n_keys + __popc(mc & lt_mask)] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)c;
                n_keys += __popc(mc);
                if (n_keys > GNN_KEYBUF - 32) flush();
            }
        }
        if (r_obj) {
            flush();
            for (int it = 0; it < k_eff; ++it) {
                const unsigned hd = (unsigned)(best[0] >> 32), hc = (unsigned)best[0];
                const unsigned md = __reduce_min_sync(0xffffffffu, hd);
                if (md == 0xffffffffu) break;                                   // no within-radius object column left
                const unsigned mc = __reduce_min_sync(0xffffffffu, hd == md ? hc : 0xffffffffu);
                if (hd == md && hc == mc) {                                     // exactly one lane owns column mc
                    atomicOr(&selw[mc >> 5], 1u << (mc & 31));
#pragma unroll
                    for (int j = 0; j < KL - 1; ++j) best[j] = best[j + 1];
                    best[KL - 1] = ~0ull;
                }
            }
        }
        __syncwarp();
        // ---- adjacency bit-words (same order of rules as the per-bit form of the general kernel)
        int count = 0;
        for (int w = lane; w < a.words; w += 32) {
            const int c0 = w * 32;
            const unsigned objm = c0 + 32 <= n_obj ? 0xffffffffu : (c0 >= n_obj ? 0u : ((1u << (n_obj - c0)) - 1u));
            unsigned on = rw[w];
            if (r_obj) on = (on & ~objm) | (on & objm & selw[w]);
            if (a.connect_all) {
                if (r_tool) on |= validw[w];
                if (r_valid) on |= toolw[w];
                if (r_tool) on &= ~toolw[w];
            }
            rw[w] = on;
            count += __popc(on);
        }
        count = __reduce_add_sync(0xffffffffu, count);
        if (lane == 0) rcount[r - row0] = count;
    }
    __syncthreads();
    // ---- this CTA's edge count -> published; exclusive prefix over the CTAs before it (look-back) and over its own rows
    const int nrows = max(row1 - row0, 0);
    int part = 0;
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) part += rcount[i];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    int total = 0;
    for (int w = 0; w < warps; ++w) total += s_red[w];
    int32_t *ag = agg + (size_t)b * gridDim.x;
    if (threadIdx.x == 0) gnn_st_release(ag + blockIdx.x, (total << 1) | 1);
    int before = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += blockDim.x) {
        int v;
        while (((v = gnn_ld_acquire(ag + j)) & 1) == 0) __nanosleep(20);
        before += v >> 1;
    }
    before = __reduce_add_sync(0xffffffffu, before);
    __syncthreads();                       // s_red readers above are done
    if (lane == 0) s_red[warp] = before;
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int w = 0; w < warps; ++w) off += s_red[w];
        s_prefix = off;
        for (int i = 0; i < nrows; ++i) { roff[i] = off; off += rcount[i]; }
    }
    __syncthreads();
    int32_t *rp = row_ptr + (size_t)b * (a.N + 1);
    int32_t *rv = recv + (size_t)b * cap, *sd = send + (size_t)b * cap;
    for (int r = row0 + warp; r < row1; r += warps) {
        const uint32_t *rw = roww + (size_t)(r - row0) * a.words;
        int off = roff[r - row0];
        if (lane == 0) rp[r] = off;
        for (int w0 = 0; w0 < a.words; w0 += 32) {
            const int w = w0 + lane;
            uint32_t m = w < a.words ? rw[w] : 0u;
            const int n = __popc(m);
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            int e = off + incl - n;
            while (m) {                     // ascending columns inside the word = adj.nonzero() order
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                if (e < cap) { rv[e] = r; sd[e] = w * 32 + bit; }
                ++e;
            }
            off += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    if (blockIdx.x == gridDim.x - 1) {     // totals and the -1 padding of the unused slots
        const int all = s_prefix + total;
        if (threadIdx.x == 0) { rp[a.N] = all; n_edges[b] = all; }
        for (int e = all + threadIdx.x; e < cap; e += blockDim.x) { rv[e] = -1; sd[e] = -1; }
    }
}

// one warp per row: expand adjacency bits to receiver / sender lists (row-major = adj.nonzero() order).
// Edges of batch element b occupy [b*cap, b*cap + n_edges[b]); the remaining slots get receiver = sender = -1.
// The exclusive scan of the row counts is folded in: every CTA sums the counts of the rows before its first one (<= N ints from
// L2, 128 threads) instead of a separate single-CTA scan launch; the warp of row r also publishes row_ptr[r] (row N: the total).
__global__ void __launch_bounds__(128)
gsd_gnn_expand_kernel(int B, int N, int words, int cap, const uint32_t *__restrict__ bits, const int32_t *__restrict__ row_count,
                      int32_t *__restrict__ row_ptr, int32_t *__restrict__ n_edges, int32_t *__restrict__ recv,
                      int32_t *__restrict__ send) {
    gsd_pdl_wait();
    __shared__ int s_part[4], s_cnt[4];
    const int b = blockIdx.y;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * warps, r = r0 + warp;
    const int32_t *rc = row_count + (size_t)b * N;
    int part = 0;
    const int lim = min(r0, N);
#pragma unroll 8
    for (int i = threadIdx.x; i < lim; i += blockDim.x) part += rc[i];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) { s_part[warp] = part; s_cnt[warp] = r < N ? rc[r] : 0; }
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warps; ++w) off += s_part[w] + (w < warp ? s_cnt[w] : 0);
    if (r > N) return;
    int32_t *rp = row_ptr + (size_t)b * (N + 1);
    if (lane == 0) rp[r] = off;
    if (r == N) { // padding
        if (lane == 0) n_edges[b] = off;
        for (int e = off + lane; e < cap; e += 32) {
            recv[(size_t)b * cap + e] = -1;
            send[(size_t)b * cap + e] = -1;
        }
        return;
    }
    // the row's bit-words: lane l holds words l, l + 32, ... (N <= 4096: at most 4), broadcast one at a time by shuffle
    const uint32_t *row = bits + ((size_t)b * N + r) * words;
    uint32_t wreg[GNN_MAX_SMEM_NODES / 1024];
#pragma unroll
    for (int j = 0; j < GNN_MAX_SMEM_NODES / 1024; ++j) wreg[j] = j * 32 + lane < words ? row[j * 32 + lane] : 0u;
#pragma unroll
    for (int j = 0; j < GNN_MAX_SMEM_NODES / 1024; ++j) {
        if (j * 32 >= words) break;
        for (int t = 0; t < 32 && j * 32 + t < words; ++t) {
            const uint32_t m = __shfl_sync(0xffffffffu, wreg[j], t);
            if ((m >> lane) & 1u) {
                const int e = off + __popc(m & ((1u << lane) - 1u));
                if (e < cap) {
                    recv[(size_t)b * cap + e] = r;
                    send[(size_t)b * cap + e] = (j * 32 + t) * 32 + lane;
                }
            }
            off += __popc(m);
        }
    }
}

extern "C" int gsd_gnn_edges_workspace_bytes(int32_t B, int32_t N, size_t *bytes) {
    if (B <= 0 || N <= 0 || !bytes) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    size_t words = (size_t)(N + 31) / 32;
    *bytes = gsd_align_up((size_t)B * N * words * 4) + gsd_align_up((size_t)B * N * 4);
    return GSD_OK;
}

extern "C" int gsd_gnn_build_edges(const GsdGnnEdges *g, void *stream) {
    if (!g || g->B <= 0 || g->N <= 0 || g->n_tool < 0 || g->n_tool > g->N || g->capacity < 0 || !g->states || !g->mask ||
        !g->tool_mask || !g->ws || !g->row_ptr || !g->n_edges || !g->receivers || !g->senders) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    if (g->N > GNN_MAX_SMEM_NODES) { gsd_set_error("N=%d exceeds %d nodes per graph", g->N, GNN_MAX_SMEM_NODES); return GSD_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    EdgeArgs a;
    a.B = g->B; a.N = g->N; a.n_tool = g->n_tool; a.topk = g->topk; a.connect_all = g->connect_all;
    a.states = g->states; a.mask = g->mask; a.tool_mask = g->tool_mask; a.thresh = g->adj_thresh; a.thresh_sq_scalar = g->adj_thresh_sq_scalar;
    a.words = (g->N + 31) / 32;
    a.bits = (uint32_t *)g->ws;
    a.row_count = (int32_t *)((char *)g->ws + gsd_align_up((size_t)g->B * g->N * a.words * 4));
    const int warps = 4, adj_warps = 8;
    // persistent CTAs (8 rows in flight each, 2 CTAs per SM at N = 2000): one wave, the position table is staged once per CTA
    int ctas = (g->N + adj_warps - 1) / adj_warps;
    const int max_ctas = (2 * 148 + g->B - 1) / g->B;
    if (ctas > max_ctas) ctas = max_ctas;
    dim3 grid(ctas, g->B);
    size_t smem = (size_t)((g->N * 3 + 1) & ~1) * 4 + (size_t)adj_warps * GNN_CAND * 8 + (size_t)adj_warps * g->N * 4 +
                  (size_t)adj_warps * a.words * 4 + (size_t)g->N;
    const int k_eff = g->topk < g->N ? g->topk : g->N;
    const int rows_per_cta = (g->N + ctas - 1) / ctas;
    if (k_eff <= 16 && (size_t)rows_per_cta * a.words * 4 <= 32 * 1024 && !getenv("GSD_GNN_ADJ_GENERAL")) {
        // fused builder: key buffers + 13 bytes per (padded) node + (2 + 1 per warp + 1 per row) words per 32 nodes (39 KB at N = 2001)
        const size_t npad = (size_t)a.words * 32;
        const size_t smem_fast = (size_t)adj_warps * GNN_KEYBUF * 8 + npad * 12 + (size_t)(2 + adj_warps + rows_per_cta) * a.words * 4 +
                                 (size_t)rows_per_cta * 8 + npad;
        static bool attr_fast = false;
        if (!attr_fast) {
            const int max_fast = 8 * GNN_KEYBUF * 8 + GNN_MAX_SMEM_NODES * 13 + (2 + 8) * (GNN_MAX_SMEM_NODES / 32) * 4 + 32 * 1024 + 8 * 1024;
            GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_gnn_edges_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_fast));
            GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_gnn_edges_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_fast));
            attr_fast = true;
        }
        if (smem_fast > (size_t)(8 * GNN_KEYBUF * 8 + GNN_MAX_SMEM_NODES * 13 + (2 + 8) * (GNN_MAX_SMEM_NODES / 32) * 4 + 32 * 1024 + 8 * 1024)) {
            gsd_set_error("edge builder: shared-memory budget exceeded");
            return GSD_ERR_UNSUPPORTED;
        }
        int32_t *agg = a.row_count;       // per-CTA edge counts + ready flag (the row-count array is not needed on this path)
        GSD_CUDA_CHECK(cudaMemsetAsync(agg, 0, (size_t)g->B * ctas * 4, st));
        if (k_eff <= 8)
            gsd_launch(gsd_gnn_edges_fused_kernel<8>, grid, dim3(adj_warps * 32), smem_fast, st, a, rows_per_cta, g->capacity, g->row_ptr, g->n_edges,
                       g->receivers, g->senders, agg);
        else
            gsd_launch(gsd_gnn_edges_fused_kernel<16>, grid, dim3(adj_warps * 32), smem_fast, st, a, rows_per_cta, g->capacity, g->row_ptr, g->n_edges,
                       g->receivers, g->senders, agg);
        GSD_LAUNCH_CHECK();
        return GSD_OK;
    }
    {
        static bool attr_set = false;
        if (!attr_set) {
            GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_gnn_adjacency_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                GNN_MAX_SMEM_NODES * 12 + 8 + 8 * GNN_CAND * 8 + 8 * GNN_MAX_SMEM_NODES * 4 +
                                                    8 * (GNN_MAX_SMEM_NODES / 32) * 4 + GNN_MAX_SMEM_NODES));
            attr_set = true;
        }
        gsd_gnn_adjacency_general_kernel<<<grid, adj_warps * 32, smem, st>>>(a);
    }
    GSD_LAUNCH_CHECK();
    dim3 grid2((g->N + 1 + warps - 1) / warps, g->B);
    gsd_launch(gsd_gnn_expand_kernel, grid2, dim3(warps * 32), 0, st, g->B, g->N, a.words, g->capacity, (const uint32_t *)a.bits,
               (const int32_t *)a.row_count, g->row_ptr, g->n_edges, g->receivers, g->senders);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// ------------------------------------------------------------------------------------------------------
// per-edge relation inputs: [attrs_r(A) | attrs_s(A) | sum|g_r - g_s| (1) | (pos_r - pos_s) for each history frame (3*n_his)]
// node-major state layout: state_t [B,N,n_his*3] is what the reference builds at model.py:132; here state is [B,n_his,N,3].
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gsd_gnn_edge_inputs_kernel(int B, int N, int cap, int n_his, int attr_dim, int n_inst, int n_p, const float *__restrict__ state,
                           const float *__restrict__ attrs, const float *__restrict__ p_instance, const int32_t *__restrict__ recv,
                           const int32_t *__restrict__ send, float *__restrict__ out) {
    gsd_pdl_trigger();
    gsd_pdl_wait();
    const int width = 2 * attr_dim + 1 + 3 * n_his;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)B * cap) return;
    const int b = (int)(e / cap);
    float *o = out + e * width;
    const int r = recv[e], s = send[e];
    if (r < 0) {
        for (int k = 0; k < width; ++k) o[k] = 0.f;
        return;
    }
    const float *at = attrs + (size_t)b * N * attr_dim;
    for (int k = 0; k < attr_dim; ++k) {
        o[k] = at[(size_t)r * attr_dim + k];
        o[attr_dim + k] = at[(size_t)s * attr_dim + k];
    }
    float gd = 0.f;
    const float *pi = p_instance + (size_t)b * n_p * n_inst;
    for (int k = 0; k < n_inst; ++k) {
        float gr = r < n_p ? pi[(size_t)r * n_inst + k] : 0.f;
        float gs = s < n_p ? pi[(size_t)s * n_inst + k] : 0.f;
        gd += fabsf(gr - gs);
    }
    o[2 * attr_dim] = gd;
    for (int h = 0; h < n_his; ++h) {
        const float *sp = state + ((size_t)b * n_his + h) * N * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[2 * attr_dim + 1 + 3 * h + c] = sp[(size_t)r * 3 + c] - sp[(size_t)s * 3 + c];
    }
}

extern "C" int gsd_gnn_edge_inputs(int32_t B, int32_t N, int32_t capacity, int32_t n_his, int32_t attr_dim, int32_t n_instance,
                                   int32_t n_p, const float *state, const float *attrs, const float *p_instance,
                                   const int32_t *receivers, const int32_t *senders, float *rel_inputs, void *stream) {
    if (B <= 0 || N <= 0 || capacity < 0 || n_his <= 0 || attr_dim < 0 || n_instance < 0 || !state || !attrs || !receivers || !senders || !rel_inputs) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    if (capacity == 0) return GSD_OK;
    long long total = (long long)B * capacity;
    gsd_launch(gsd_gnn_edge_inputs_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, B, N, capacity, n_his,
               attr_dim, n_instance, n_p, state, attrs, p_instance, receivers, senders, rel_inputs);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// ------------------------------------------------------------------------------------------------------
// fused edge epilogue + segment reduce.  Nodes are globally indexed (b*N + n), edges of element b live at [b*cap, ...).
//   agg[node] = sum over the node's incoming edges of ReLU(A[e] + P[node][0:F] + P[b*N + send(e)][F:2F])
// One warp per (receiver) row for ordinary rows; rows with many edges (the tool rows: every object is a sender) are split
// over GNN_SPLIT CTAs writing partials that a second kernel sums in fixed order.
// ------------------------------------------------------------------------------------------------------
#define GNN_SPLIT 32

// ONE launch: CTAs [0, rows_grid) take the light rows (one warp per receiver), the others the heavy (tool) rows, GNN_SPLIT CTAs
// per row; the last of them to finish (ticket counter) adds the partials up in FIXED order — deterministic, no second launch.
template <int VEC_PER_LANE> // F = 128 * VEC_PER_LANE
__global__ void __launch_bounds__(128)
gsd_gnn_aggregate_kernel(int B, int N, int cap, int n_light, int n_heavy, int rows_grid, const int32_t *__restrict__ row_ptr,
                         const int32_t *__restrict__ send, const float4 *__restrict__ A, const float4 *__restrict__ P,
                         float4 *__restrict__ agg, float4 *__restrict__ partial /* [B*n_heavy][GNN_SPLIT][F4] */, int32_t *__restrict__ tickets) {
    gsd_pdl_trigger();
    gsd_pdl_wait();
    constexpr int F4 = 32 * VEC_PER_LANE; // float4 per feature row
    if ((int)blockIdx.x < rows_grid) {
        const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const long long row = (long long)blockIdx.x * warps + warp;
        if (row >= (long long)B * n_light) return;
        const int b = (int)(row / n_light), r = (int)(row % n_light);
        const int32_t *rp = row_ptr + (size_t)b * (N + 1);
        const int e0 = min(rp[r], cap), e1 = min(rp[r + 1], cap);
        const size_t node = (size_t)b * N + r;
        float4 pr[VEC_PER_LANE], acc[VEC_PER_LANE];
#pragma unroll
        for (int v = 0; v < VEC_PER_LANE; ++v) {
            pr[v] = P[node * 2 * F4 + v * 32 + lane];
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // The sender indices of up to 32 edges are fetched with ONE coalesced load and broadcast by shuffle (a dependent index load
        // per edge serialised the gathers: 24 % of the DRAM rate, long_scoreboard 30 warps per issue); two edges' rows in flight.
        for (int eb = e0; eb < e1; eb += 32) {
            const int ne = min(32, e1 - eb);
            const int my_s = lane < ne ? send[(size_t)b * cap + eb + lane] : 0;
            for (int j = 0; j < ne; j += 2) {
                const bool two = j + 1 < ne;
                const size_t ge0 = (size_t)b * cap + eb + j, ge1 = ge0 + (two ? 1 : 0);
                const size_t sn0 = (size_t)b * N + __shfl_sync(0xffffffffu, my_s, j);
                const size_t sn1 = (size_t)b * N + __shfl_sync(0xffffffffu, my_s, two ? j + 1 : j);
                float4 a0[VEC_PER_LANE], s0[VEC_PER_LANE], a1[VEC_PER_LANE], s1[VEC_PER_LANE];
#pragma unroll
                for (int v = 0; v < VEC_PER_LANE; ++v) {
                    a0[v] = A[ge0 * F4 + v * 32 + lane];
                    s0[v] = P[sn0 * 2 * F4 + F4 + v * 32 + lane];
                    a1[v] = A[ge1 * F4 + v * 32 + lane];
                    s1[v] = P[sn1 * 2 * F4 + F4 + v * 32 + lane];
                }
#pragma unroll
                for (int v = 0; v < VEC_PER_LANE; ++v) {      // edge j, then edge j + 1: the same summation order as one edge at a time
                    acc[v].x += fmaxf(a0[v].x + pr[v].x + s0[v].x, 0.f);
                    acc[v].y += fmaxf(a0[v].y + pr[v].y + s0[v].y, 0.f);
                    acc[v].z += fmaxf(a0[v].z + pr[v].z + s0[v].z, 0.f);
                    acc[v].w += fmaxf(a0[v].w + pr[v].w + s0[v].w, 0.f);
                    if (two) {
                        acc[v].x += fmaxf(a1[v].x + pr[v].x + s1[v].x, 0.f);
                        acc[v].y += fmaxf(a1[v].y + pr[v].y + s1[v].y, 0.f);
                        acc[v].z += fmaxf(a1[v].z + pr[v].z + s1[v].z, 0.f);
                        acc[v].w += fmaxf(a1[v].w + pr[v].w + s1[v].w, 0.f);
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VEC_PER_LANE; ++v) agg[node * F4 + v * 32 + lane] = acc[v];
        return;
    }
    // ---- heavy rows
    __shared__ int s_last;
    const int hidx = (int)blockIdx.x - rows_grid;
    const int hrow = hidx / GNN_SPLIT, split = hidx % GNN_SPLIT; // hrow in [0, B*n_heavy)
    const int b = hrow / n_heavy, r = n_light + hrow % n_heavy;
    const int32_t *rp = row_ptr + (size_t)b * (N + 1);
    const int e0 = min(rp[r], cap), e1 = min(rp[r + 1], cap);
    const int n = e1 - e0;
    const int per = (n + GNN_SPLIT - 1) / GNN_SPLIT;
    const int s0 = e0 + split * per, s1 = min(e1, s0 + per);
    const size_t node = (size_t)b * N + r;
    for (int col = threadIdx.x; col < F4; col += blockDim.x) {
        const float4 pr = P[node * 2 * F4 + col];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = s0; e < s1; e += 4) { // 4 independent gathers in flight
            float4 a4[4], s4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ee = min(e + u, s1 - 1);
                const size_t ge = (size_t)b * cap + ee;
                const size_t snode = (size_t)b * N + send[ge];
                a4[u] = A[ge * F4 + col];
                s4[u] = P[snode * 2 * F4 + F4 + col];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (e + u < s1) {
                    acc.x += fmaxf(a4[u].x + pr.x + s4[u].x, 0.f);
                    acc.y += fmaxf(a4[u].y + pr.y + s4[u].y, 0.f);
                    acc.z += fmaxf(a4[u].z + pr.z + s4[u].z, 0.f);
                    acc.w += fmaxf(a4[u].w + pr.w + s4[u].w, 0.f);
                }
            }
        }
        partial[((size_t)hrow * GNN_SPLIT + split) * F4 + col] = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = (atomicAdd(&tickets[hrow], 1) == GNN_SPLIT - 1);
        if (s_last) tickets[hrow] = 0;      // self-cleaning: the counters are zero again when the kernel ends (no memset per call)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int col = threadIdx.x; col < F4; col += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < GNN_SPLIT; s += 8) {   // loads of 8 partials in flight, added in index order
            float4 p[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) p[u] = __ldcg(&partial[((size_t)hrow * GNN_SPLIT + s + u) * F4 + col]);
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += p[u].x; acc.y += p[u].y; acc.z += p[u].z; acc.w += p[u].w; }
        }
        agg[((size_t)b * N + r) * F4 + col] = acc;
    }
}

extern "C" int gsd_gnn_aggregate_workspace_bytes(int32_t B, int32_t n_heavy, int32_t F, size_t *bytes) {
    if (B <= 0 || n_heavy < 0 || F <= 0 || !bytes) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    const size_t rows = (size_t)B * (n_heavy > 0 ? n_heavy : 1);
    *bytes = gsd_align_up(rows * GNN_SPLIT * F * 4) + gsd_align_up(rows * 4);   // partial sums, then the ticket counters
    return GSD_OK;
}

// rows [0, N - n_heavy) : warp per row;  rows [N - n_heavy, N) (tool nodes) : split over GNN_SPLIT CTAs
extern "C" int gsd_gnn_aggregate(int32_t B, int32_t N, int32_t capacity, int32_t F, int32_t n_heavy, const int32_t *row_ptr,
                                 const int32_t *senders, const float *A, const float *P, void *ws, float *agg, void *stream) {
    if (B <= 0 || N <= 0 || capacity < 0 || n_heavy < 0 || n_heavy > N || !row_ptr || !senders || !A || !P || !agg || (n_heavy > 0 && !ws)) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    if (F % 128 != 0 || F > 512) { gsd_set_error("feature width %d must be a multiple of 128 and <= 512", F); return GSD_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_light = N - n_heavy;
    const int warps = 4;
    const long long rows = (long long)B * n_light;
    const int rows_grid = (int)((rows + warps - 1) / warps);
    const int heavy_grid = B * n_heavy * GNN_SPLIT;
    if (rows_grid + heavy_grid == 0) return GSD_OK;
    int32_t *tickets = nullptr;
    if (n_heavy > 0) {
        tickets = (int32_t *)((char *)ws + gsd_align_up((size_t)B * n_heavy * GNN_SPLIT * F * 4));   // zero on entry, zero on exit
    }
    const dim3 grid((unsigned)(rows_grid + heavy_grid));
#define GSD_AGG_ARGS B, N, capacity, n_light, n_heavy, rows_grid, row_ptr, senders, (const float4 *)A, (const float4 *)P, (float4 *)agg, (float4 *)ws, tickets
    switch (F / 128) {
    case 1: gsd_launch(gsd_gnn_aggregate_kernel<1>, grid, dim3(128), 0, st, GSD_AGG_ARGS); break;
    case 2: gsd_launch(gsd_gnn_aggregate_kernel<2>, grid, dim3(128), 0, st, GSD_AGG_ARGS); break;
    case 3: gsd_launch(gsd_gnn_aggregate_kernel<3>, grid, dim3(128), 0, st, GSD_AGG_ARGS); break;
    default: gsd_launch(gsd_gnn_aggregate_kernel<4>, grid, dim3(128), 0, st, GSD_AGG_ARGS); break;
    }
#undef GSD_AGG_ARGS
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// ------------------------------------------------------------------------------------------------------
// farthest point sampling: one CTA per batch element, distances in shared memory, first-max tie break.
//   radius <= 0 : classic FPS to npoints (dgl.geometry.farthest_point_sampler, squared distances)
//   radius  > 0 : radius-terminated FPS of fps_rad_idx_torch (data/utils.py:50-65, euclidean distances): stops when the
//                 largest distance to the picked set is <= radius; count[b] receives the number of picks
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
gsd_fps_kernel(int N, int npoints, float radius, const float *__restrict__ pos, const int64_t *__restrict__ start_idx,
               int64_t *__restrict__ out, int32_t *__restrict__ count) {
    extern __shared__ float sdist[]; // [N]
    __shared__ float rv[32];
    __shared__ int ri[32];
    __shared__ int s_cur;
    const int b = blockIdx.x;
    const float *p = pos + (size_t)b * N * 3;
    int64_t *o = out + (size_t)b * npoints;
    const bool rad = radius > 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sdist[i] = 3.0e38f;
    if (threadIdx.x == 0) {
        long long s = start_idx ? start_idx[b] : 0;
        s_cur = (int)(s < 0 ? 0 : (s >= N ? N - 1 : s));
    }
    __syncthreads();
    int picked = 0;
    for (int it = 0; it < npoints; ++it) {
        const int cur = s_cur;
        if (threadIdx.x == 0) o[it] = cur;
        picked = it + 1;
        const float cx = p[3 * cur], cy = p[3 * cur + 1], cz = p[3 * cur + 2];
        float best = -1.f;
        int best_i = 0x7fffffff;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            float d = dist2_rn(p[3 * i], p[3 * i + 1], p[3 * i + 2], cx, cy, cz);
            if (rad) d = sqrtf(d);
            float m = fminf(sdist[i], d);
            sdist[i] = m;
            if (m > best) { best = m; best_i = i; }
        }
#pragma unroll
        for (int of = 16; of >= 1; of >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, of);
            int oi = __shfl_xor_sync(0xffffffffu, best_i, of);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
        }
        if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = best; ri[threadIdx.x >> 5] = best_i; }
        __syncthreads();
        if (threadIdx.x < 32) {
            best = threadIdx.x < (blockDim.x >> 5) ? rv[threadIdx.x] : -1.f;
            best_i = threadIdx.x < (blockDim.x >> 5) ? ri[threadIdx.x] : 0x7fffffff;
#pragma unroll
            for (int of = 16; of >= 1; of >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, of);
                int oi = __shfl_xor_sync(0xffffffffu, best_i, of);
                if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
            }
            if (threadIdx.x == 0) {
                rv[0] = best;
                s_cur = best_i;
            }
        }
        __syncthreads();
        if (rad && !(rv[0] > radius)) break;
        __syncthreads();
    }
    if (threadIdx.x == 0 && count) count[b] = picked;
    for (int i = picked + threadIdx.x; i < npoints; i += blockDim.x) o[i] = -1;
}

extern "C" int gsd_fps(int32_t B, int32_t N, int32_t npoints, float radius, const float *pos, const int64_t *start_idx,
                       int64_t *out_idx, int32_t *count, void *stream) {
    if (B <= 0 || N <= 0 || npoints <= 0 || !pos || !out_idx) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if ((size_t)N * 4 > 200 * 1024) { gsd_set_error("N=%d too large for the shared-memory FPS kernel", N); return GSD_ERR_UNSUPPORTED; }
    if (radius <= 0.f && npoints > N) { gsd_set_error("npoints > N"); return GSD_ERR_INVALID; }
    static bool attr_set = false;
    if (!attr_set) {
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    int threads = N >= 1024 ? 1024 : ((N + 31) / 32) * 32;
    gsd_fps_kernel<<<B, threads, (size_t)N * 4, (cudaStream_t)stream>>>(N, npoints, radius, pos, start_idx, out_idx, count);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// ------------------------------------------------------------------------------------------------------
// error-compensated TF32 operands for the dense layers: fp32 accuracy on the tensor cores from ONE TF32 GEMM over a
// 3x longer K.  With t = relu?(x + add), hi = t rounded to 10 explicit mantissa bits and lo = t - hi, the activation row
// becomes [lo | hi | hi] and the weight row [w_hi | w_lo | w_hi]:  lo.w_hi + hi.w_lo + hi.w_hi = t.w up to the dropped
// lo.w_lo term (2^-22 relative).  The small products come first in K so they are summed before the accumulator is large.
// One pass also applies the previous layer's ReLU / residual add and optionally writes t itself.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float t) { return __uint_as_float((__float_as_uint(t) + 0x1000u) & 0xffffe000u); }

__global__ void __launch_bounds__(256)
gsd_tf32_pack_kernel(long long rows, int F4, int relu, int weight_layout, const float4 *__restrict__ x, const float4 *__restrict__ add,
                     float4 *__restrict__ full, float4 *__restrict__ out) {
    const long long n4 = rows * F4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = x[i];
        if (add) { const float4 a = add[i]; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (full) full[i] = v;
        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        const long long r = i / F4;
        const int c = (int)(i - r * F4);
        float4 *o = out + r * 3 * F4 + c;
        if (weight_layout) { o[0] = h; o[F4] = l; o[2 * F4] = h; }
        else { o[0] = l; o[F4] = h; o[2 * F4] = h; }
    }
}

extern "C" int gsd_tf32_pack(int64_t rows, int32_t F, int32_t relu, int32_t weight_layout, const float *x, const float *add, float *full,
                             float *out, void *stream) {
    if (rows < 0 || F <= 0 || (F % 4) != 0 || (rows > 0 && (!x || !out))) { gsd_set_error("invalid arguments (F must be a multiple of 4)"); return GSD_ERR_INVALID; }
    if (rows == 0) return GSD_OK;
    long long n4 = rows * (F / 4);
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gsd_tf32_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, F / 4, relu, weight_layout, (const float4 *)x, (const float4 *)add,
                                                                              (float4 *)full, (float4 *)out);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

// ------------------------------------------------------------------------------------------------------
// rollout glue: the host-side tensor plumbing around one autoregressive step as two launches.
//   gsd_gnn_rollout_pre   action row of the tool nodes <- eef_delta (dynamics_module.py:110-111); the particle encoder's input rows
//                         [attrs | state over history | motion over history | action] (model.py:132-160 for state_dim = 3), and a
//                         contiguous copy of the current positions for the edge builder (dynamics_module.py:127)
//   gsd_gnn_rollout_post  pred_pos = state[-1] + clamp(pred_motion) (model.py:241-244), the tool node advanced by eef_delta and
//                         the history shifted by one frame (dynamics_module.py:145-158), one thread per (batch element, node)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gsd_gnn_rollout_pre_kernel(int B, int N, int n_obj, int n_his, int attr_dim, int state_dim, int motion, int has_action,
                           const float *__restrict__ states, const float *__restrict__ attrs, float *__restrict__ action,
                           const float *__restrict__ eef_delta, int delta_stride, float *__restrict__ p_inputs, float *__restrict__ cur) {
    gsd_pdl_trigger();
    gsd_pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N, n = i - b * N;
    const int sw = state_dim * n_his;                       // state_dim 3: xyz per frame; 1: z only (model.py:144-148); 0: none
    const int width = attr_dim + sw + (motion ? 3 * (n_his - 1) : 0) + (has_action ? 3 : 0);
    float *o = p_inputs + (size_t)i * width;
    for (int k = 0; k < attr_dim; ++k) o[k] = attrs[(size_t)i * attr_dim + k];
    o += attr_dim;
    float prev[3] = {0.f, 0.f, 0.f};
    for (int h = 0; h < n_his; ++h) {
        const float *s = states + (((size_t)b * n_his + h) * N + n) * 3;
        const float x = s[0], y = s[1], z = s[2];
        if (state_dim == 3) { o[3 * h] = x; o[3 * h + 1] = y; o[3 * h + 2] = z; }
        else if (state_dim == 1) o[h] = z;
        if (motion && h > 0) {
            float *m = o + sw + 3 * (h - 1);
            m[0] = x - prev[0]; m[1] = y - prev[1]; m[2] = z - prev[2];
        }
        prev[0] = x; prev[1] = y; prev[2] = z;
    }
    cur[(size_t)i * 3] = prev[0]; cur[(size_t)i * 3 + 1] = prev[1]; cur[(size_t)i * 3 + 2] = prev[2];
    if (has_action) {
        float *a = action + (size_t)i * 3;
        float ax = a[0], ay = a[1], az = a[2];
        if (n >= n_obj) {
            const float *d = eef_delta + (size_t)b * delta_stride;
            ax = d[0]; ay = d[1]; az = d[2];
            a[0] = ax; a[1] = ay; a[2] = az;
        }
        float *q = o + sw + (motion ? 3 * (n_his - 1) : 0);
        q[0] = ax; q[1] = ay; q[2] = az;
    }
}

__global__ void __launch_bounds__(256)
gsd_gnn_rollout_post_kernel(int B, int N, int n_obj, int n_his, float *__restrict__ states, const float *__restrict__ motion,
                            const float *__restrict__ eef_delta, int delta_stride, float clampv, float *__restrict__ pred) {
    gsd_pdl_trigger();
    gsd_pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N, n = i - b * N;
    float *last = states + (((size_t)b * n_his + (n_his - 1)) * N + n) * 3;
    float nx = last[0], ny = last[1], nz = last[2];
    if (n < n_obj) {
        const float *m = motion + ((size_t)b * n_obj + n) * 3;
        nx += fminf(fmaxf(m[0], -clampv), clampv); ny += fminf(fmaxf(m[1], -clampv), clampv); nz += fminf(fmaxf(m[2], -clampv), clampv);
        float *p = pred + ((size_t)b * n_obj + n) * 3;
        p[0] = nx; p[1] = ny; p[2] = nz;
    } else {
        const float *d = eef_delta + (size_t)b * delta_stride;
        nx += d[0]; ny += d[1]; nz += d[2];
    }
    for (int h = 0; h + 1 < n_his; ++h) {      // this thread owns every history entry of its node: no cross-thread hazard
        float *dst = states + (((size_t)b * n_his + h) * N + n) * 3;
        const float *src = dst + (size_t)N * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
    last[0] = nx; last[1] = ny; last[2] = nz;
}

extern "C" int gsd_gnn_rollout_pre(int32_t B, int32_t N, int32_t n_obj, int32_t n_his, int32_t attr_dim, int32_t state_dim, int32_t motion, int32_t has_action,
                                   const float *states, const float *attrs, float *action, const float *eef_delta, int32_t delta_stride,
                                   float *p_inputs, float *cur, void *stream) {
    if (B <= 0 || N <= 0 || n_obj < 0 || n_obj > N || n_his <= 0 || attr_dim < 0 || !states || (attr_dim > 0 && !attrs) || !p_inputs || !cur ||
        (has_action && (!action || !eef_delta)) || (state_dim != 0 && state_dim != 1 && state_dim != 3)) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    gsd_launch(gsd_gnn_rollout_pre_kernel, dim3((B * N + 255) / 256), dim3(256), 0, (cudaStream_t)stream, B, N, n_obj, n_his, attr_dim, state_dim,
               motion, has_action, states, attrs, action, eef_delta, delta_stride, p_inputs, cur);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

extern "C" int gsd_gnn_rollout_post(int32_t B, int32_t N, int32_t n_obj, int32_t n_his, float *states, const float *motion,
                                    const float *eef_delta, int32_t delta_stride, float clampv, float *pred, void *stream) {
    if (B <= 0 || N <= 0 || n_obj < 0 || n_obj > N || n_his <= 0 || !states || (n_obj > 0 && (!motion || !pred)) || (n_obj < N && !eef_delta)) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    gsd_launch(gsd_gnn_rollout_post_kernel, dim3((B * N + 255) / 256), dim3(256), 0, (cudaStream_t)stream, B, N, n_obj, n_his, states, motion,
               eef_delta, delta_stride, clampv, pred);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
