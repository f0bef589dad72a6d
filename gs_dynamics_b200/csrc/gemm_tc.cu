// gemm_tc.cu — the dense layers of the GNN step (path B) on the 5th-generation tensor cores, fp32-accurate.
//
// Replaces the nn.Linear layers of /root/reference/src/gnn/model.py:16-22 (Encoder), 36-47 (Propagator), 58-67
// (ParticlePredictor) as DynamicsPredictor.forward composes them (model.py:202-241).  SURVEY.md fact 4: the GNN step is
// dominated by these GEMMs (M = 20 000 edge rows or 2 001 node rows, K = 512, N = 512 / 1024).
//
// out[M, N] = act( A[M, K] . W[N, K]^T + bias[N] + res1[M, N] + res2[M, N] ),  act = ReLU or identity, everything fp32.
//
// One warp-specialised, persistent sm_100a kernel per layer (round 1 used cuBLAS over a K-concatenated, 3x larger packed copy of
// the activations written by a separate pack kernel):
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) brings a 128 x 32 fp32 tile of A and the matching tiles of W_hi / W_lo
//     (the weight's TF32 head and remainder, split once per weight update) into a shared-memory ring;
//   * two groups of four converter warps (alternating k-blocks) read the A tile from shared memory, split it in registers —
//     a_hi = tf32(a), a_lo = tf32(a - a_hi) — and store both halves into TENSOR MEMORY (tcgen05.st): no packed operand in HBM, no
//     extra launch, and the MMAs read A from TMEM instead of shared memory.  That matters: with all operands in shared memory
//     the kernel was measured shared-memory-bandwidth bound (operand reads of 12 K=8 MMAs + converter traffic + TMA writes =
//     192 KB per 128x128x32 block = 1536 cycles at 128 B/clk against 768 cycles of tensor work);
//   * one thread issues tcgen05.mma kind::tf32 with A from TMEM and B from shared memory: per 8-wide k-step
//     D_small += a_lo.w_hi, D_small += a_hi.w_lo, D_big += a_hi.w_hi — two fp32 accumulators in TMEM, summed in the epilogue
//     (error vs float64 at K = 512: ~5e-7 of the |a||w| bound, fp32-SIMT level);
//   * four epilogue warps read the accumulators back with tcgen05.ld and apply bias / residual(s) / ReLU on the way to HBM; at
//     64-wide tiles two accumulator sets alternate, so the epilogue of a tile overlaps the main loop of the next one.
// No tensor-core GEMM library is involved; SASS shows UTCHMMA-class (UTC*MMA), LDTM and UTMALDG (profiles/r2_sass_gemm_tc.txt).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 32, UMMA_K = 8;
constexpr uint32_t A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;   // 16 KB, one 128-byte swizzle row per matrix row

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem_ptr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) & 0x3ffff) >> 4);        // start address  [0, 14)
    d |= (uint64_t)0 << 16;                                      // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                            // stride byte offset [32, 46)
    d |= (uint64_t)1 << 46;                                      // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                      // layout: SWIZZLE_128B
    return d;
}

// D[tmem] (+)= A[smem] . B[smem]^T, TF32 inputs, fp32 accumulation
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with the A operand in tensor memory (128 lanes x 8 columns of TF32)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
        "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
        "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])),
        "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])),
        "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
// one lane of a converged warp (the tcgen05.mma / commit issuer); everything around it stays warp-uniform so that descriptors
// and tensor-memory addresses live in uniform registers — inside an `if (lane == 0)` region the compiler has to move every
// operand of every MMA through an ELECT / R2UR.BROADCAST retry loop, which was measured at ~80-110 cycles per tcgen05.mma issue,
// more than the 64 cycles a 128x128x8 TF32 product executes
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// round-to-nearest TF32 with two integer instructions (cvt.rna.tf32.f32 issues at a quarter of the ALU rate, and the converter warps
// are the pipeline's bottleneck): add half an ulp of the 10-bit mantissa, clear the 13 low bits (ties away from zero, like .rna)
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// 32 consecutive accumulator columns of this thread's row (TMEM lane); the caller waits (tmem_ld_wait) before using v
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BLOCK_N>
struct GemmCfg {
    static constexpr uint32_t B_TILE_BYTES = BLOCK_N * BLOCK_K * 4;
    static constexpr uint32_t STAGE_BYTES = A_TILE_BYTES + 2 * B_TILE_BYTES;         // a (raw fp32) | w_hi | w_lo
    static constexpr int STAGES = 4;
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024;             // + slack to align the ring to 1024 B
    // tensor memory (512 columns): [0, 256) accumulator sets of (small | big) x BLOCK_N columns — two sets at BLOCK_N = 64 (the
    // epilogue of a tile overlaps the next main loop), one at 128; [256, 512) the A ring: per stage a_hi | a_lo (2 x 32 columns)
    static constexpr int ACC_BUFS = 256 / (2 * BLOCK_N);
    static constexpr uint32_t A_COLS0 = 256;
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(ACC_BUFS >= 1 && STAGES * STAGE_BYTES <= 200 * 1024, "tile configuration");
    static_assert(A_COLS0 + STAGES * 2 * BLOCK_K <= 512, "tensor memory budget");
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
    // the same with N = 2 * BLOCK_N: one instruction against the stacked [w_hi ; w_lo] tiles (adjacent in a stage)
    static constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BLOCK_N) >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
};

// warp 0: TMA producer; warp 1: TMEM owner + MMA issuer; warps 2-5 and 6-9: two converter groups (alternate k-blocks);
// warps 10-17: epilogue (a warp may only read the TMEM lane quarter warp_id % 4; warps 10-13 take the left half of the
// tile's columns, 14-17 the right half — the accumulators are not double-buffered at 128-wide tiles, so the epilogue is exposed)
constexpr int GEMM_THREADS = 576, PAIR_THREADS = 448, EPI_WARPS = 8;
// optional pipeline trace of CTA 0 (GSD_GEMM_DBG=1): globaltimer stamps per role and k-block, printed by the launcher
__device__ long long g_gemm_trace[6 * 64];
__device__ int g_gemm_trace_on;
#define GEMM_TRACE(slot, idx) do { if (trace_on && (idx) < 64) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_gemm_trace[(slot) * 64 + (idx)] = t_; } } while (0)
constexpr int CVT_GROUPS = 2;

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gsd_gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_wh,
                       const __grid_constant__ CUtensorMap tm_wl, int M, int N, int K, const float *__restrict__ bias,
                       const float *__restrict__ res1, const float *__restrict__ res2, int relu, float *__restrict__ out, long long ldo, int dbg) {
    // dbg (GSD_GEMM_MODE, timing experiments only — results are wrong): 1 converters skip the TMEM stores, 2 converters skip all
    // work, 4 only the big product, 8 one accumulator, 32 W tiles loaded for the first ring round only, 64 same for A
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES, ACC_BUFS = Cfg::ACC_BUFS;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_raw[STAGES], full_cvt[STAGES], empty[STAGES], tmem_full[ACC_BUFS], tmem_empty[ACC_BUFS];
    __shared__ uint32_t tmem_base_slot;
    gsd_pdl_trigger();           // the next kernel of the chain may set itself up while this one runs (it waits before reading)
    const bool trace_on = g_gemm_trace_on && blockIdx.x == 0;
    if (threadIdx.x == 0) GEMM_TRACE(5, 0);
    uint8_t *ring = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = K / BLOCK_K;
    const int tiles_n = N / BLOCK_N, tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
    const int n_tiles = tiles_m * tiles_n;
    auto sA = [&](int s) { return ring + (size_t)s * Cfg::STAGE_BYTES; };
    auto sWh = [&](int s) { return ring + (size_t)s * Cfg::STAGE_BYTES + A_TILE_BYTES; };
    auto sWl = [&](int s) { return ring + (size_t)s * Cfg::STAGE_BYTES + A_TILE_BYTES + Cfg::B_TILE_BYTES; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_raw[s], 1); mbar_init(&full_cvt[s], 4); mbar_init(&empty[s], 1); }
        for (int b = 0; b < ACC_BUFS; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS); }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_wl) : "memory");
    }
    if (warp == 1) {   // tensor-memory allocation: one warp, power-of-two columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // the allocation is the whole tensor memory (512 columns), so its base is column 0 of lane 0: a compile-time constant keeps
    // every tensor-memory address in uniform registers
    if (tmem_base_slot != 0u) __trap();
    constexpr uint32_t tmem_base = 0u;
    gsd_pdl_wait();              // everything above (barriers, tensor memory, descriptor prefetch) overlapped the previous kernel
    if (threadIdx.x == 0) GEMM_TRACE(5, 1);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    GEMM_TRACE(0, kb + (tile != (int)blockIdx.x) * 16);
                    const bool first_round = tile == (int)blockIdx.x && kb < STAGES;
                    const bool ld_w = !(dbg & 32) || first_round, ld_a = !(dbg & 64) || first_round;
                    if (!ld_w && !ld_a) { mbar_arrive_cta(&full_raw[stage]); }
                    else {
                        mbar_expect_tx(&full_raw[stage], (ld_a ? A_TILE_BYTES : 0) + (ld_w ? 2 * Cfg::B_TILE_BYTES : 0));
                        if (ld_a) tma_load_2d(sA(stage), &tm_a, kb * BLOCK_K, m0, &full_raw[stage]);      // rows beyond M arrive as zeros
                        if (ld_w) {
                            tma_load_2d(sWh(stage), &tm_wh, kb * BLOCK_K, n0, &full_raw[stage]);
                            tma_load_2d(sWl(stage), &tm_wl, kb * BLOCK_K, n0, &full_raw[stage]);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop (uniform control flow), one elected lane issues =====
        {
            int stage = 0;
            uint32_t phase = 0;
            int ti = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
                const int ab = ti % ACC_BUFS;
                // two accumulators per tile: the 2^-11-scaled correction terms are summed among themselves (d_small) before they meet
                // the large sum (d_big).  With ONE accumulator the result carried a systematic bias (the tensor core's fp32
                // accumulation truncates; 192 accumulating MMAs per output instead of 64): 1.4e-6 instead of 5e-7 of the |a||w|
                // bound, and 1.2e-4 instead of 5e-5 on the model's predicted motion — outside the stated parity tolerance.
                // Layout [big | small]: w_hi and w_lo tiles are adjacent in a stage, so ONE N = 2 * BLOCK_N instruction computes
                // a_hi . [w_hi ; w_lo]^T into both (2 instead of 3 instructions per k-step; an instruction costs ~100 cycles up to N = 128).
                const uint32_t d_big = tmem_base + (uint32_t)(ab * 2 * BLOCK_N), d_small = d_big + BLOCK_N;
                mbar_wait(&tmem_empty[ab], ((ti / ACC_BUFS) & 1) ^ 1);   // the epilogue has drained this accumulator set
                // The wait for k-block kb + 1 is issued BETWEEN the third and the last k-step of k-block kb: the barrier poll, the
                // fence and the descriptor arithmetic then run while the tensor pipe still has queued work instead of after it drained.
                mbar_wait(&full_cvt[stage], phase);       // TMA landed (w_hi, w_lo in shared memory) and a converter group has written a_hi / a_lo to TMEM
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    if (lane == 0) GEMM_TRACE(3, kb + (tile != (int)blockIdx.x) * 16);
                    const uint32_t a_hi = tmem_base + Cfg::A_COLS0 + (uint32_t)(stage * 2 * BLOCK_K), a_lo = a_hi + BLOCK_K;
                    const uint64_t w_hi = umma_desc_sw128(sWh(stage));
                    const uint32_t acc0 = (uint32_t)(kb != 0);
                    auto issue = [&](int k) {
                        const uint64_t ko = (uint64_t)((k * UMMA_K * 4) >> 4);   // 32 bytes per k-step inside the 128-byte swizzle row
                        const uint32_t kc = (uint32_t)(k * UMMA_K);              // 8 TMEM columns per k-step
                        umma_tf32_ts(d_big, a_hi + kc, w_hi + ko, Cfg::IDESC2, k ? 1u : acc0);   // big += a_hi.w_hi, small += a_hi.w_lo
                        if (!(dbg & 4)) umma_tf32_ts(d_small, a_lo + kc, w_hi + ko, Cfg::IDESC, 1u);   // small += a_lo.w_hi
                    };
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K - 1; ++k) issue(k);
                    }
                    __syncwarp();
                    int nstage = stage + 1;
                    uint32_t nphase = phase;
                    if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
                    const bool more = kb + 1 < kblocks;
                    if (more) { mbar_wait(&full_cvt[nstage], nphase); tc_fence_after(); }
                    if (elect_one()) {
                        issue(BLOCK_K / UMMA_K - 1);
                        umma_commit(&empty[stage]);       // the stage returns to the producer when these MMAs have read it
                        if (!more) umma_commit(&tmem_full[ab]);   // accumulators complete
                    }
                    __syncwarp();
                    if (lane == 0 && tile == (int)blockIdx.x) GEMM_TRACE(5, 16 + kb);
                    stage = nstage; phase = nphase;
                }
            }
        }
    } else if (warp < 2 + 4 * CVT_GROUPS) {
        // ===== converters: group g splits every CVT_GROUPS-th k-block in shared memory =====
        const int g = (warp - 2) / 4;
        const int ct = threadIdx.x - 64 - g * 128;
        const int cq = warp & 3;                          // TMEM lane quarter this warp may access
        const int crow = cq * 32 + lane;                  // = the matrix row of the tile it converts
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                if (it % CVT_GROUPS == g) {
                    mbar_wait(&full_raw[stage], phase);
                    if (ct == 0) GEMM_TRACE(1, kb + (tile != (int)blockIdx.x) * 16);
                    // this thread's matrix row: 8 x 16-byte chunks, chunk c stored at position c ^ (row % 8) (128-byte swizzle)
                    const float4 *rowp = reinterpret_cast<const float4 *>(sA(stage) + (size_t)crow * 128);
                    float hi[32], lo[32];
                    if (!(dbg & 2)) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = rowp[c ^ (crow & 7)];
                        hi[4 * c] = round_tf32(v.x); hi[4 * c + 1] = round_tf32(v.y); hi[4 * c + 2] = round_tf32(v.z); hi[4 * c + 3] = round_tf32(v.w);
                        lo[4 * c] = round_tf32(v.x - hi[4 * c]); lo[4 * c + 1] = round_tf32(v.y - hi[4 * c + 1]);
                        lo[4 * c + 2] = round_tf32(v.z - hi[4 * c + 2]); lo[4 * c + 3] = round_tf32(v.w - hi[4 * c + 3]);
                    }
                    }
                    const uint32_t ta = tmem_base + ((uint32_t)(cq * 32) << 16) + Cfg::A_COLS0 + (uint32_t)(stage * 2 * BLOCK_K);
                    if (!(dbg & 3)) {
                        tmem_st32(ta, hi);
                        tmem_st32(ta + BLOCK_K, lo);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    } else if (!(dbg & 2)) {
                        if (hi[0] + lo[31] == 123.456f) tmem_st32(ta, hi);     // keep the loads and the split alive
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(&full_cvt[stage]);
                    if (ct == 0) GEMM_TRACE(2, kb + (tile != (int)blockIdx.x) * 16);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: 4 warps = 128 accumulator rows; overlaps the next tile's main loop when there are two accumulator pairs =====
        const int quarter = warp & 3;                     // a warp may only read its own quarter of the TMEM lanes
        const int row_in_tile = quarter * 32 + lane;
        const int chalf = (warp - (2 + 4 * CVT_GROUPS)) >> 2;          // 0: columns [0, BLOCK_N / 2), 1: the rest
        int ti = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
            const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
            const int ab = ti % ACC_BUFS;
            mbar_wait(&tmem_full[ab], (ti / ACC_BUFS) & 1);
            tc_fence_after();
            if (threadIdx.x == 320) GEMM_TRACE(4, (tile != (int)blockIdx.x));
            const int row = m0 + row_in_tile;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * 2 * BLOCK_N);
            // this thread's BLOCK_N / 2 output columns are pulled into registers first and the accumulators handed back at once:
            // the bias / residual loads and the stores below overlap the NEXT tile's main loop (the accumulators are not
            // double-buffered at 128-wide tiles, so everything before the hand-back is exposed time)
            constexpr int EC = BLOCK_N / 2;
            float v[EC];
#pragma unroll
            for (int c = 0; c < EC; c += 32) {
                float vs[32];
                tmem_ld32(lane_addr + (uint32_t)(chalf * EC + c), v + c);
                tmem_ld32(lane_addr + (uint32_t)(BLOCK_N + chalf * EC + c), vs);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[c + j] += vs[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&tmem_empty[ab]);
            if (row < M) {
                const int col = n0 + chalf * EC;
                float *o = out + (size_t)row * ldo + col;
#pragma unroll
                for (int j = 0; j < EC; j += 4) {
                    float4 r = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (bias) {
                        const float4 bq = __ldg(reinterpret_cast<const float4 *>(bias + col + j));
                        r.x += bq.x; r.y += bq.y; r.z += bq.z; r.w += bq.w;
                    }
                    if (res1) {
                        const float4 q = __ldg(reinterpret_cast<const float4 *>(res1 + (size_t)row * ldo + col + j));
                        r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
                    }
                    if (res2) {
                        const float4 q = __ldg(reinterpret_cast<const float4 *>(res2 + (size_t)row * ldo + col + j));
                        r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
                    }
                    if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                    *reinterpret_cast<float4 *>(o + j) = r;
                }
            }
            if (threadIdx.x == 320) GEMM_TRACE(4, 2 + (tile != (int)blockIdx.x));
        }
    }
    // ----- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
    if (threadIdx.x == 32) GEMM_TRACE(5, 2);
}

// ---- CTA-pair variant (cta_group::2) for the edge-row layers ---------------------------------------------------------------------
// Two CTAs of a cluster (one TPC) work on one 256 x 128 tile.  Each CTA converts ITS 128 rows of A into its own tensor memory and
// stages HALF of the W tiles (64 of the 128 weight rows); the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2
// (M = 256): one instruction drives both SMs' tensor cores, each reading its A rows from its TMEM and both B halves.  Per CTA
// and k-block that is 32 KB instead of 48 KB through L2 -> shared memory (the single-CTA kernel sits on the L2 -> SM path:
// 48 KB per 768 tensor cycles = 62 B/clk/SM against ~43 B/clk/SM measured chip-wide), half the B-operand shared-memory reads, and
// half as many MMA instructions per flop (the per-instruction cost measured ~115 cycles for a 64-cycle 128x128x8 TF32 product).
// Synchronisation: full_a (local: own A tile landed) -> own converters; full_w (leader's: both CTAs' TMA loads complete_tx on it,
// cp.async.bulk.tensor .cta_group::2) and full_cvt (leader's: 8 converter warps of both CTAs arrive, remotely for the peer)
// -> MMA thread; tcgen05.commit .multicast::cluster frees the stage / publishes the accumulators in BOTH CTAs; tmem_empty
// (leader's: 8 epilogue warps of both CTAs).
constexpr int PAIR_N = 128, PAIR_HALF_N = 64;
constexpr uint32_t PAIR_B_BYTES = PAIR_HALF_N * BLOCK_K * 4;                       // 8 KB: this CTA's half of a W tile
constexpr uint32_t PAIR_STAGE_BYTES = A_TILE_BYTES + 2 * PAIR_B_BYTES;             // 32 KB
constexpr int PAIR_STAGES = 4;                                                     // tensor-memory A ring: 4 x (32 + 32) columns
constexpr uint32_t PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + 1024;
constexpr uint32_t PAIR_A_COLS0 = 256, PAIR_TMEM_COLS = 512;                       // [0,128) small | [128,256) big | [256,512) A ring
constexpr uint32_t PAIR_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PAIR_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier of either CTA of the pair (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
gsd_gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_wh,
                            const __grid_constant__ CUtensorMap tm_wl, int M, int N, int K, const float *__restrict__ bias,
                            const float *__restrict__ res1, const float *__restrict__ res2, int relu, float *__restrict__ out, long long ldo) {
    constexpr int STAGES = PAIR_STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_a[STAGES], full_w[STAGES], full_cvt[STAGES], empty[STAGES], tmem_full, tmem_empty;
    __shared__ uint32_t tmem_base_slot;
    gsd_pdl_trigger();
    uint8_t *ring = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int kblocks = K / BLOCK_K;
    const int tiles_n = N / PAIR_N, tiles_m = (M + 255) / 256;
    const int n_tiles = tiles_m * tiles_n;
    auto sA = [&](int s) { return ring + (size_t)s * PAIR_STAGE_BYTES; };
    auto sWh = [&](int s) { return ring + (size_t)s * PAIR_STAGE_BYTES + A_TILE_BYTES; };
    auto sWl = [&](int s) { return ring + (size_t)s * PAIR_STAGE_BYTES + A_TILE_BYTES + PAIR_B_BYTES; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_a[s], 1); mbar_init(&full_w[s], 1); mbar_init(&full_cvt[s], 8); mbar_init(&empty[s], 1); }
        mbar_init(&tmem_full, 1);
        mbar_init(&tmem_empty, 8);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_wl) : "memory");
    }
    if (warp == 1) {   // tensor memory of the pair: one warp per CTA
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(PAIR_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers are initialised before anything is signalled across the pair
    tc_fence_after();
    if (tmem_base_slot != 0u) __trap();   // the whole tensor memory was allocated: base = column 0 (keeps addresses in uniform registers)
    constexpr uint32_t tmem_base = 0u;
    gsd_pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A rows -> local barrier; own half of W -> the leader's barrier =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = (tile / tiles_n) * 256 + (int)rank * BLOCK_M, n0 = (tile % tiles_n) * PAIR_N + (int)rank * PAIR_HALF_N;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full_a[stage], A_TILE_BYTES);
                    tma_load_2d(sA(stage), &tm_a, kb * BLOCK_K, m0, &full_a[stage]);        // rows beyond M arrive as zeros
                    if (rank == 0) mbar_expect_tx(&full_w[stage], 4 * PAIR_B_BYTES);        // w_hi, w_lo halves of both CTAs
                    const uint32_t fw = mapa_u32(smem_u32(&full_w[stage]), 0);
                    tma_load_2d_pair(sWh(stage), &tm_wh, kb * BLOCK_K, n0, fw);
                    tma_load_2d_pair(sWl(stage), &tm_wl, kb * BLOCK_K, n0, fw);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp 1 of the leader CTA (uniform control flow, one elected lane issues) =====
        if (rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int ti = 0;
            const uint32_t d_small = tmem_base, d_big = tmem_base + PAIR_N;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
                mbar_wait(&tmem_empty, (ti & 1) ^ 1);     // both CTAs' epilogues have drained the accumulators
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_w[stage], phase);      // both halves of w_hi / w_lo landed (in both CTAs)
                    mbar_wait(&full_cvt[stage], phase);    // both CTAs' converters have written a_hi / a_lo to their tensor memory
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + PAIR_A_COLS0 + (uint32_t)(stage * 2 * BLOCK_K), a_lo = a_hi + BLOCK_K;
                    const uint64_t w_hi = umma_desc_sw128(sWh(stage)), w_lo = umma_desc_sw128(sWl(stage));
                    const uint32_t acc0 = (uint32_t)(kb != 0);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t ko = (uint64_t)((k * UMMA_K * 4) >> 4);
                            const uint32_t kc = (uint32_t)(k * UMMA_K);
                            umma_tf32_ts_pair(d_small, a_lo + kc, w_hi + ko, PAIR_IDESC, k ? 1u : acc0);
                            umma_tf32_ts_pair(d_small, a_hi + kc, w_lo + ko, PAIR_IDESC, 1u);
                            umma_tf32_ts_pair(d_big, a_hi + kc, w_hi + ko, PAIR_IDESC, k ? 1u : acc0);
                        }
                        umma_commit_pair(&empty[stage]);
                        if (kb == kblocks - 1) umma_commit_pair(&tmem_full);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 2 + 4 * CVT_GROUPS) {
        // ===== converters (both CTAs): group g splits every CVT_GROUPS-th k-block of this CTA's A rows =====
        const int g = (warp - 2) / 4;
        const int cq = warp & 3;
        const int crow = cq * 32 + lane;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                if (it % CVT_GROUPS == g) {
                    mbar_wait(&full_a[stage], phase);
                    const float4 *rowp = reinterpret_cast<const float4 *>(sA(stage) + (size_t)crow * 128);
                    float hi[32], lo[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = rowp[c ^ (crow & 7)];
                        hi[4 * c] = round_tf32(v.x); hi[4 * c + 1] = round_tf32(v.y); hi[4 * c + 2] = round_tf32(v.z); hi[4 * c + 3] = round_tf32(v.w);
                        lo[4 * c] = round_tf32(v.x - hi[4 * c]); lo[4 * c + 1] = round_tf32(v.y - hi[4 * c + 1]);
                        lo[4 * c + 2] = round_tf32(v.z - hi[4 * c + 2]); lo[4 * c + 3] = round_tf32(v.w - hi[4 * c + 3]);
                    }
                    const uint32_t ta = tmem_base + ((uint32_t)(cq * 32) << 16) + PAIR_A_COLS0 + (uint32_t)(stage * 2 * BLOCK_K);
                    tmem_st32(ta, hi);
                    tmem_st32(ta + BLOCK_K, lo);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&full_cvt[stage]), 0));
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue (both CTAs): this CTA's 128 rows of the tile =====
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        int ti = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
            const int m0 = (tile / tiles_n) * 256 + (int)rank * BLOCK_M, n0 = (tile % tiles_n) * PAIR_N;
            mbar_wait(&tmem_full, ti & 1);
            tc_fence_after();
            const int row = m0 + row_in_tile;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < PAIR_N; c0 += 32) {
                float vb[32];
                {
                    float vs[32];
                    tmem_ld32(lane_addr + (uint32_t)c0, vs);
                    tmem_ld32(lane_addr + (uint32_t)(PAIR_N + c0), vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) vb[j] += vs[j];
                }
                if (row < M) {
                    const int col = n0 + c0;
                    float *o = out + (size_t)row * ldo + col;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 r = make_float4(vb[j], vb[j + 1], vb[j + 2], vb[j + 3]);
                        if (bias) {
                            const float4 bq = __ldg(reinterpret_cast<const float4 *>(bias + col + j));
                            r.x += bq.x; r.y += bq.y; r.z += bq.z; r.w += bq.w;
                        }
                        if (res1) {
                            const float4 q = __ldg(reinterpret_cast<const float4 *>(res1 + (size_t)row * ldo + col + j));
                            r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
                        }
                        if (res2) {
                            const float4 q = __ldg(reinterpret_cast<const float4 *>(res2 + (size_t)row * ldo + col + j));
                            r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
                        }
                        if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                        *reinterpret_cast<float4 *>(o + j) = r;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tmem_empty), 0));
        }
    }
    // ----- teardown: neither CTA may leave while the other can still signal its barriers or read its shared memory
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(PAIR_TMEM_COLS) : "memory");
    }
}

// ---- host: tensor maps ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// row-major fp32 matrix [rows, cols] with row stride ld (elements); box = box_rows x 32 columns, 128-byte swizzle, zero fill
int make_map(CUtensorMap *map, const float *ptr, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { gsd_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GSD_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gsd_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GSD_ERR_CUDA; }
    return GSD_OK;
}

template <int BLOCK_N>
int launch_gemm(long long M, int N, int K, const float *A, long long lda, const float *W_hi, const float *W_lo, const float *bias,
                const float *res1, const float *res2, int relu, float *out, long long ldo, cudaStream_t st) {
    using Cfg = GemmCfg<BLOCK_N>;
    static bool attr_done = false;
    if (!attr_done) {
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_gemm_tf32x3_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_done = true;
    }
    CUtensorMap ta, th, tl;
    int rc;
    if ((rc = make_map(&ta, A, M, K, lda, BLOCK_M))) return rc;
    if ((rc = make_map(&th, W_hi, N, K, K, BLOCK_N))) return rc;
    if ((rc = make_map(&tl, W_lo, N, K, K, BLOCK_N))) return rc;
    const int tiles = (int)((M + BLOCK_M - 1) / BLOCK_M) * (N / BLOCK_N);
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool trace = getenv("GSD_GEMM_DBG") != nullptr;
    if (trace) { const int one = 1; cudaMemcpyToSymbol(g_gemm_trace_on, &one, 4); }
    static const int dbg_mode = getenv("GSD_GEMM_MODE") ? atoi(getenv("GSD_GEMM_MODE")) : 0;
    gsd_launch(gsd_gemm_tf32x3_kernel<BLOCK_N>, dim3(tiles < sms ? tiles : sms), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, ta, th, tl, (int)M, N, K,
               bias, res1, res2, relu, out, ldo, dbg_mode);
    GSD_LAUNCH_CHECK();
    if (trace) {
        long long h[6 * 64];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_gemm_trace, sizeof(h));
        const long long t0 = h[5 * 64];
        fprintf(stderr, "GEMM BN=%d M=%lld N=%d grid=%d: kernel start 0, setup done %lld, end %lld (ns)\n  kb: tma_issue full_raw_seen cvt_done mma_issue\n", BLOCK_N, M, N,
                tiles < sms ? tiles : sms, h[5 * 64 + 1] - t0, h[5 * 64 + 2] - t0);
        for (int i = 0; i < 32; ++i) fprintf(stderr, "  %2d: %6lld %6lld %6lld %6lld\n", i, h[i] - t0, h[64 + i] - t0, h[128 + i] - t0, h[192 + i] - t0);
        fprintf(stderr, "  MMA warp, first tile: kb: issue_start  issued_and_committed\n");
        for (int i = 0; i < 16; ++i) fprintf(stderr, "  %2d: %6lld %6lld\n", i, h[192 + i] - t0, h[5 * 64 + 16 + i] - t0);
        fprintf(stderr, "  epilogue start/end first tile: %lld %lld  last tile: %lld %lld\n", h[256] - t0, h[258] - t0, h[257] - t0, h[259] - t0);
    }
    return GSD_OK;
}

// ---- small helpers: weight split, and the two layer shapes that are not tensor-core work (K <= 32, N <= 8) ---------------
__global__ void gsd_tf32_split_kernel(long long n, const float *__restrict__ w, float *__restrict__ hi, float *__restrict__ lo) {
    gsd_pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = w[i], h = round_tf32(v);
    hi[i] = h;
    lo[i] = round_tf32(v - h);
}

// out[m, n] = act(bias[n] + sum_k x[m, k] W[n, k]), K <= 32 (the first layer of the encoders: K = 5 / 14).  W^T and 32 rows of x
// (transposed) are staged in shared memory; a thread owns a 4-row x 4-column register tile: per k one conflict-free LDS.128 of
// weights and one broadcast LDS.128 of inputs feed 16 FMAs (the first version read 5 shared-memory wavefronts per 4 FMAs and
// was shared-memory bound: 41 us for the 20 000 edge rows).
constexpr int SK_ROWS = 32;
__global__ void __launch_bounds__(256)
gsd_linear_small_k_kernel(long long M, int N, int K, const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ bias,
                          int relu, float *__restrict__ out) {
    gsd_pdl_trigger();
    gsd_pdl_wait();
    extern __shared__ float sk_smem[];          // [K][N] transposed weights, then [K][SK_ROWS] transposed inputs
    float *sWt = sk_smem, *sxT = sk_smem + (size_t)K * N;
    // persistent CTAs: the transposed weights (strided, uncoalesced reads) are staged ONCE per CTA, then it walks its row blocks
    // (staging them per 32-row block cost as much as the block's arithmetic: 23 us for the 20 000 edge rows)
    for (int i = threadIdx.x; i < N * K; i += 256) { const int k = i / N, n = i % N; sWt[i] = __ldg(W + (size_t)n * K + k); }   // conflict-free stores
    const int n4 = N / 4;
    const int rg = threadIdx.x / 32;             // 8 row groups of 4 rows
    const long long blocks = (M + SK_ROWS - 1) / SK_ROWS;
    for (long long blk = blockIdx.x; blk < blocks; blk += gridDim.x) {
        const long long m0 = blk * SK_ROWS;
        __syncthreads();                         // the previous block's readers of sxT are done (and sWt is staged)
        for (int i = threadIdx.x; i < SK_ROWS * K; i += 256) { const int r = i / K, k = i % K; const long long m = m0 + r; sxT[k * SK_ROWS + r] = m < M ? x[m * K + k] : 0.f; }
        __syncthreads();
        for (int c = threadIdx.x % 32; c < n4; c += 32) {
            const float4 b = bias ? __ldg(reinterpret_cast<const float4 *>(bias) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 acc[4] = {b, b, b, b};
            for (int k = 0; k < K; ++k) {
                const float4 w = *reinterpret_cast<const float4 *>(sWt + (size_t)k * N + 4 * c);
                const float4 xv = *reinterpret_cast<const float4 *>(sxT + k * SK_ROWS + 4 * rg);
                const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc[r].x = fmaf(xr[r], w.x, acc[r].x); acc[r].y = fmaf(xr[r], w.y, acc[r].y);
                    acc[r].z = fmaf(xr[r], w.z, acc[r].z); acc[r].w = fmaf(xr[r], w.w, acc[r].w);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const long long m = m0 + 4 * rg + r;
                if (m >= M) break;
                float4 v = acc[r];
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                reinterpret_cast<float4 *>(out + m * N)[c] = v;
            }
        }
    }
}

// out[m, n] = bias[n] + sum_k x[m, k] W[n, k], N <= 8 (the 512 -> 3 motion head): one warp per row
__global__ void __launch_bounds__(256)
gsd_linear_small_n_kernel(long long M, int N, int K, const float *__restrict__ x, long long ldx, const float *__restrict__ W,
                          const float *__restrict__ bias, float *__restrict__ out) {
    gsd_pdl_trigger();
    gsd_pdl_wait();
    const long long m = (long long)blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
        const float4 xv = *reinterpret_cast<const float4 *>(x + m * ldx + k);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            if (n < N) {
                const float4 w = __ldg(reinterpret_cast<const float4 *>(W + (size_t)n * K + k));
                acc[n] += xv.x * w.x + xv.y * w.y + xv.z * w.z + xv.w * w.w;
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    if (lane < N) {
        float v = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) v = (lane == n) ? acc[n] : v;
        out[m * N + lane] = v + (bias ? bias[lane] : 0.f);
    }
}

}  // namespace

int launch_gemm_pair(long long M, int N, int K, const float *A, long long lda, const float *W_hi, const float *W_lo, const float *bias,
                     const float *res1, const float *res2, int relu, float *out, long long ldo, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_gemm_tf32x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
        attr_done = true;
    }
    CUtensorMap ta, th, tl;
    int rc;
    if ((rc = make_map(&ta, A, M, K, lda, BLOCK_M))) return rc;
    if ((rc = make_map(&th, W_hi, N, K, K, PAIR_HALF_N))) return rc;
    if ((rc = make_map(&tl, W_lo, N, K, K, PAIR_HALF_N))) return rc;
    const int tiles = (int)((M + 255) / 256) * (N / PAIR_N);
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int pairs = tiles < sms / 2 ? tiles : sms / 2;
    gsd_launch(gsd_gemm_tf32x3_pair_kernel, dim3(2 * pairs), dim3(PAIR_THREADS), PAIR_SMEM_BYTES, st, ta, th, tl, (int)M, N, K, bias, res1, res2, relu,
               out, ldo);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

extern "C" int gsd_linear_tf32x3(int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *W_hi, const float *W_lo,
                                 const float *bias, const float *res1, const float *res2, int32_t relu, float *out, int64_t ldo, void *stream) {
    if (M < 0 || N <= 0 || K <= 0 || (M > 0 && (!A || !W_hi || !W_lo || !out))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (K % BLOCK_K != 0 || N % 64 != 0 || lda % 4 != 0 || ldo % 4 != 0 || lda < K || ldo < N) {
        gsd_set_error("gsd_linear_tf32x3 needs K %% 32 == 0, N %% 64 == 0 and 16-byte aligned rows (K=%d N=%d lda=%lld ldo=%lld)", K, N, (long long)lda,
                      (long long)ldo);
        return GSD_ERR_UNSUPPORTED;
    }
    if (((uintptr_t)A | (uintptr_t)W_hi | (uintptr_t)W_lo | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)res1 | (uintptr_t)res2) & 15) {
        gsd_set_error("pointers must be 16-byte aligned");
        return GSD_ERR_INVALID;
    }
    if (M == 0) return GSD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // wide tiles when there are enough rows to fill the machine with them, narrow tiles for the node-level layers
    const long long tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
    const char *force = getenv("GSD_GEMM_BN");
    const int fbn = force ? atoi(force) : 0;
    // GSD_GEMM_PAIR=1: CTA pairs (cta_group::2) on 256 x 128 tiles
    const char *fp = getenv("GSD_GEMM_PAIR");
    const int pair_mode = fp ? atoi(fp) : -1;
    if (N % PAIR_N == 0 && fbn == 0 && pair_mode == 1)   // measured slower than the single-CTA kernel: opt-in only
        return launch_gemm_pair(M, N, K, A, lda, W_hi, W_lo, bias, res1, res2, relu, out, ldo, st);
    if (fbn == 64) return launch_gemm<64>(M, N, K, A, lda, W_hi, W_lo, bias, res1, res2, relu, out, ldo, st);
    // tile width: fewer waves win (a 128-wide tile costs ~1.15x a 64-wide one per k-block: both sit on the per-instruction floor)
    const long long w64 = (tiles_m * (N / 64) + 147) / 148, w128 = N % 128 == 0 ? (tiles_m * (N / 128) + 147) / 148 : (1LL << 40);
    if (N % 128 == 0 && (fbn == 128 || (fbn == 0 && 115 * w128 < 100 * w64)))
        return launch_gemm<128>(M, N, K, A, lda, W_hi, W_lo, bias, res1, res2, relu, out, ldo, st);
    return launch_gemm<64>(M, N, K, A, lda, W_hi, W_lo, bias, res1, res2, relu, out, ldo, st);
}

extern "C" int gsd_tf32_split(int64_t n, const float *w, float *w_hi, float *w_lo, void *stream) {
    if (n < 0 || (n > 0 && (!w || !w_hi || !w_lo))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (n == 0) return GSD_OK;
    gsd_tf32_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, w, w_hi, w_lo);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

extern "C" int gsd_linear_small(int64_t M, int32_t N, int32_t K, const float *x, int64_t ldx, const float *W, const float *bias, int32_t relu,
                                float *out, void *stream) {
    if (M < 0 || N <= 0 || K <= 0 || (M > 0 && (!x || !W || !out))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (M == 0) return GSD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 32 && N % 4 == 0 && ldx == K && (size_t)(N + SK_ROWS) * K * 4 <= 96 * 1024) {
        const size_t smem = (size_t)(N + SK_ROWS) * K * 4;
        static bool attr_done = false;
        if (!attr_done) {
            GSD_CUDA_CHECK(cudaFuncSetAttribute(gsd_linear_small_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_done = true;
        }
        const long long blocks = (M + SK_ROWS - 1) / SK_ROWS;
        const unsigned grid = (unsigned)(blocks < 2 * 148 ? blocks : 2 * 148);      // two CTAs per SM (30 KB of shared memory each at N = 512, K = 14)
        gsd_launch(gsd_linear_small_k_kernel, dim3(grid), dim3(256), smem, st, (long long)M, N, K, x, W, bias, relu, out);
    } else if (N <= 8 && K % 4 == 0 && ldx % 4 == 0 && !relu) {
        gsd_launch(gsd_linear_small_n_kernel, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, (long long)M, N, K, x, (long long)ldx, W, bias, out);
    } else {
        gsd_set_error("gsd_linear_small handles K <= 32 (N %% 4 == 0, dense rows) or N <= 8 (K %% 4 == 0, no activation)");
        return GSD_ERR_UNSUPPORTED;
    }
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
