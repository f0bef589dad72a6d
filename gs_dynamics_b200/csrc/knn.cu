// knn.cu — exact k nearest neighbours of every point of a cloud among the other points, sm_100a.
//
//   gsd_knn   replaces o3d_knn, /root/reference/src/tracking/helpers.py:97-115 (Open3D KD-tree on the CPU, one Python
//             iteration per point; called for k = 3 at initialisation, train_utils.py:113, and k = 20 after the first frame,
//             train_utils.py:359) — "minutes at 100k points" in SURVEY.md §8 A11.
//
// Brute force in double precision: Open3D searches in float64, and the tracking priors use exp(-2000 d^2) of these
// distances, so the arithmetic is kept in fp64 (B200: 64 fp64 lanes per SM; 1e10 pairs at n = 100k ≈ 6e10 fp64 ops ≈ ms).
// One thread per query; candidates stream through shared memory in tiles; each thread keeps its k best (distance, index)
// sorted in local memory — after the first tiles almost every candidate fails the single compare against the current
// k-th distance, so insertions are rare.  Output order: ascending distance, ties by lower index.
#include "common.cuh"

#define KNN_TILE 512
#define KNN_MAX_K 64

__global__ void __launch_bounds__(128)
gsd_knn_kernel(int n, int k, const float *__restrict__ pts, double *__restrict__ sq_dist, int32_t *__restrict__ idx) {
    __shared__ double sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int qi = min(i, n - 1);
    const double qx = (double)pts[3 * qi], qy = (double)pts[3 * qi + 1], qz = (double)pts[3 * qi + 2];
    double bd[KNN_MAX_K];
    int bi[KNN_MAX_K];
    for (int j = 0; j < k; ++j) { bd[j] = 1.0e300; bi[j] = -1; }
    double worst = 1.0e300;
    for (int c0 = 0; c0 < n; c0 += KNN_TILE) {
        const int nc = min(KNN_TILE, n - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < nc; t += blockDim.x) {
            sx[t] = (double)pts[3 * (c0 + t)]; sy[t] = (double)pts[3 * (c0 + t) + 1]; sz[t] = (double)pts[3 * (c0 + t) + 2];
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < nc; ++t) {
            const double dx = qx - sx[t], dy = qy - sy[t], dz = qz - sz[t];
            const double d = dx * dx + dy * dy + dz * dz;
            if (d < worst && c0 + t != qi) { // candidates arrive in index order: a tie never displaces an earlier index
                int p = k - 1;
                while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                bd[p] = d;
                bi[p] = c0 + t;
                worst = bd[k - 1];
            }
        }
    }
    if (i < n) {
        for (int j = 0; j < k; ++j) { sq_dist[(size_t)i * k + j] = bd[j]; idx[(size_t)i * k + j] = bi[j]; }
    }
}

extern "C" int gsd_knn(int32_t n, int32_t k, const float *pts, double *sq_dist, int32_t *idx, void *stream) {
    if (n < 0 || k <= 0 || (n > 0 && (!pts || !sq_dist || !idx))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (k > KNN_MAX_K) { gsd_set_error("k=%d exceeds %d", k, KNN_MAX_K); return GSD_ERR_UNSUPPORTED; }
    if (n > 0 && k > n - 1) { gsd_set_error("k=%d needs at least k+1 points (n=%d)", k, n); return GSD_ERR_INVALID; }
    if (n == 0) return GSD_OK;
    gsd_knn_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, k, pts, sq_dist, idx);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
