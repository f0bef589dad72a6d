// baseline/upstream_structure/raster_upstream.cu — BENCHMARK BASELINE, NOT PRODUCT CODE.
//
// A clearly-labelled STRUCTURAL RE-CREATION of the upstream rasterizer the reference depends on
// (JonathonLuiten/diff-gaussian-rasterization-w-depth, a fork of graphdeco-inria/diff-gaussian-rasterization; its source is NOT
// under /root/reference — README.md:26-35 clones it — so the real thing cannot be built or timed here).  SURVEY.md §8(c)/(d) and
// BASELINE.md §3 allow this stand-in for the north-star's "reference CUDA rasterizer on 1 GPU" baseline: the published
// algorithm with its published execution structure —
//     preprocess (one thread per Gaussian)  ->  CUB InclusiveSum of tiles touched  ->  D2H read of num_rendered (host sync)
//     -> duplicateWithKeys (tile << 32 | depth bits)  ->  CUB DeviceRadixSort::SortPairs over 32 + log2(tiles) bits
//     -> identifyTileRanges  ->  render: ONE 16x16 CTA PER TILE, 256-wide shared-memory staging, early exit by block vote
//     backward: one CTA per tile walking its list back to front, NINE float atomicAdd per (pixel, Gaussian) pair,
//     then the per-Gaussian cov2D / projection / cov3D chain rule.
// Written from the algorithm description, compiled for sm_100a with the same flags as the product.  Nothing
// under gs_dynamics_b200/ imports, links or calls this file; bench.py --impl upstream_structure and tests/ do.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#define BX 16
#define BY 16
#define BSZ (BX * BY)

struct GsuGeom {   // carved from one caller-allocated buffer
    float *depths; float2 *xy; float *cov3D; float4 *conic_o; uint32_t *tiles; uint32_t *offsets; int *radii_i; void *scan_tmp; size_t scan_bytes;
};
static size_t al(size_t x) { return (x + 255) / 256 * 256; }
static size_t carve_geom(int P, char *base, GsuGeom *g) {
    size_t off = 0;
    auto take = [&](size_t b) { char *r = base ? base + off : nullptr; off += al(b); return r; };
    size_t n = P > 0 ? P : 1;
    g->depths = (float *)take(n * 4); g->xy = (float2 *)take(n * 8); g->cov3D = (float *)take(n * 24);
    g->conic_o = (float4 *)take(n * 16); g->tiles = (uint32_t *)take(n * 4); g->offsets = (uint32_t *)take(n * 4);
    g->radii_i = (int *)take(n * 4);
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    g->scan_bytes = tb; g->scan_tmp = take(tb);
    return off;
}
struct GsuBin { uint64_t *keys_u, *keys; uint32_t *vals_u, *vals; void *sort_tmp; size_t sort_bytes; };
static size_t carve_bin(int64_t R, char *base, GsuBin *b) {
    size_t off = 0;
    auto take = [&](size_t bb) { char *r = base ? base + off : nullptr; off += al(bb); return r; };
    size_t n = R > 0 ? R : 1;
    b->keys_u = (uint64_t *)take(n * 8); b->keys = (uint64_t *)take(n * 8); b->vals_u = (uint32_t *)take(n * 4); b->vals = (uint32_t *)take(n * 4);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    b->sort_bytes = tb; b->sort_tmp = take(tb);
    return off;
}
struct GsuImg { uint2 *ranges; uint32_t *n_contrib; float *final_T; };
static size_t carve_img(int W, int H, char *base, GsuImg *im) {
    size_t off = 0;
    auto take = [&](size_t bb) { char *r = base ? base + off : nullptr; off += al(bb); return r; };
    const int tiles = ((W + BX - 1) / BX) * ((H + BY - 1) / BY);
    im->ranges = (uint2 *)take((size_t)tiles * 8); im->n_contrib = (uint32_t *)take((size_t)W * H * 4); im->final_T = (float *)take((size_t)W * H * 4);
    return off;
}

__device__ __forceinline__ float3 mulpt(const float *m, float3 p) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13], m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ void rot_of(const float *q, float R[3][3]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
    R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
    R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
}
// M = J * Wview (2x3) for a view-space point t
__device__ __forceinline__ void proj_jac(const float *V, float3 t, float fx, float fy, float tanx, float tany, float M[2][3], float *txo, float *tyo) {
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float tx = fminf(limx, fmaxf(-limx, t.x / t.z)) * t.z, ty = fminf(limy, fmaxf(-limy, t.y / t.z)) * t.z;
    const float J00 = fx / t.z, J02 = -(fx * tx) / (t.z * t.z), J11 = fy / t.z, J12 = -(fy * ty) / (t.z * t.z);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        M[0][b] = J00 * V[b * 4 + 0] + J02 * V[b * 4 + 2];
        M[1][b] = J11 * V[b * 4 + 1] + J12 * V[b * 4 + 2];
    }
    *txo = tx; *tyo = ty;
}

__global__ void gsu_preprocess(int P, const float *means3D, const float *scales, float mod, const float *rots, const float *opac,
                               const float *V, const float *Pm, int W, int H, float tanx, float tany, float fx, float fy,
                               int *radii, GsuGeom g, int gx, int gy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    radii[i] = 0; g.tiles[i] = 0;
    const float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    const float3 t = mulpt(V, p);
    if (t.z <= 0.2f) return;
    const float hx = Pm[0] * p.x + Pm[4] * p.y + Pm[8] * p.z + Pm[12], hy = Pm[1] * p.x + Pm[5] * p.y + Pm[9] * p.z + Pm[13];
    const float hw = Pm[3] * p.x + Pm[7] * p.y + Pm[11] * p.z + Pm[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    float R[3][3];
    rot_of(rots + 4 * i, R);
    const float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
    float S[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[a][b] = R[a][0] * s[0] * s[0] * R[b][0] + R[a][1] * s[1] * s[1] * R[b][1] + R[a][2] * s[2] * s[2] * R[b][2];
    float *c3 = g.cov3D + 6 * i;
    c3[0] = S[0][0]; c3[1] = S[0][1]; c3[2] = S[0][2]; c3[3] = S[1][1]; c3[4] = S[1][2]; c3[5] = S[2][2];
    float M[2][3], tx, ty;
    proj_jac(V, t, fx, fy, tanx, tany, M, &tx, &ty);
    float MS[2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) MS[a][b] = M[a][0] * S[0][b] + M[a][1] * S[1][b] + M[a][2] * S[2][b];
    const float ca = MS[0][0] * M[0][0] + MS[0][1] * M[0][1] + MS[0][2] * M[0][2] + 0.3f;
    const float cb = MS[0][0] * M[1][0] + MS[0][1] * M[1][1] + MS[0][2] * M[1][2];
    const float cc = MS[1][0] * M[1][0] + MS[1][1] * M[1][1] + MS[1][2] * M[1][2] + 0.3f;
    const float det = ca * cc - cb * cb;
    if (det == 0.0f) return;
    const float di = 1.f / det;
    const float mid = 0.5f * (ca + cc);
    const float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det)), l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float rad = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
    const float px = ((hx * pw + 1.0f) * W - 1.0f) * 0.5f, py = ((hy * pw + 1.0f) * H - 1.0f) * 0.5f;
    const int minx = min(gx, max(0, (int)((px - rad) / BX))), miny = min(gy, max(0, (int)((py - rad) / BY)));
    const int maxx = min(gx, max(0, (int)((px + rad + BX - 1) / BX))), maxy = min(gy, max(0, (int)((py + rad + BY - 1) / BY)));
    if ((maxx - minx) * (maxy - miny) == 0) return;
    g.depths[i] = t.z;
    radii[i] = (int)rad;
    g.radii_i[i] = (int)rad;
    g.xy[i] = make_float2(px, py);
    g.conic_o[i] = make_float4(cc * di, -cb * di, ca * di, opac[i]);
    g.tiles[i] = (uint32_t)((maxx - minx) * (maxy - miny));
}

__global__ void gsu_duplicate(int P, const float2 *xy, const float *depths, const uint32_t *offsets, const int *radii, uint64_t *keys, uint32_t *vals, int gx, int gy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || radii[i] <= 0) return;
    uint32_t off = (i == 0) ? 0 : offsets[i - 1];
    const float px = xy[i].x, py = xy[i].y, rad = (float)radii[i];
    const int minx = min(gx, max(0, (int)((px - rad) / BX))), miny = min(gy, max(0, (int)((py - rad) / BY)));
    const int maxx = min(gx, max(0, (int)((px + rad + BX - 1) / BX))), maxy = min(gy, max(0, (int)((py + rad + BY - 1) / BY)));
    for (int y = miny; y < maxy; ++y)
        for (int x = minx; x < maxx; ++x) {
            keys[off] = ((uint64_t)(y * gx + x) << 32) | (uint64_t)__float_as_uint(depths[i]);
            vals[off] = (uint32_t)i;
            ++off;
        }
}

__global__ void gsu_ranges(int L, const uint64_t *keys, uint2 *ranges) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) ranges[cur].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) { ranges[prev].y = idx; ranges[cur].x = idx; }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

__global__ void __launch_bounds__(BSZ)
gsu_render_fwd(const uint2 *ranges, const uint32_t *point_list, int W, int H, const float2 *xy, const float *colors, const float *depths,
               const float4 *conic_o, float *final_T, uint32_t *n_contrib, const float *bg, float *out_color, float *out_depth) {
    const int gxn = (W + BX - 1) / BX;
    const int pxi = blockIdx.x * BX + threadIdx.x, pyi = blockIdx.y * BY + threadIdx.y;
    const int pix_id = W * pyi + pxi;
    const float pxf = (float)pxi, pyf = (float)pyi;
    const bool inside = pxi < W && pyi < H;
    bool done = !inside;
    const uint2 range = ranges[blockIdx.y * gxn + blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BSZ - 1) / BSZ;
    int toDo = (int)(range.y - range.x);
    __shared__ int c_id[BSZ];
    __shared__ float2 c_xy[BSZ];
    __shared__ float4 c_co[BSZ];
    const int rank = threadIdx.y * BX + threadIdx.x;
    float T = 1.0f, C[3] = {0.f, 0.f, 0.f}, D = 0.f;
    uint32_t contributor = 0, last_contributor = 0;
    for (int i = 0; i < rounds; ++i, toDo -= BSZ) {
        if (__syncthreads_count(done) == BSZ) break;
        const int progress = i * BSZ + rank;
        if ((int)range.x + progress < (int)range.y) {
            const int id = (int)point_list[range.x + progress];
            c_id[rank] = id; c_xy[rank] = xy[id]; c_co[rank] = conic_o[id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BSZ, toDo); ++j) {
            ++contributor;
            const float2 c = c_xy[j];
            const float dx = c.x - pxf, dy = c.y - pyf;
            const float4 co = c_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, co.w * __expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const int id = c_id[j];
            const float w = alpha * T;
            C[0] += colors[id * 3] * w; C[1] += colors[id * 3 + 1] * w; C[2] += colors[id * 3 + 2] * w;
            D += depths[id] * w;
            T = test_T;
            last_contributor = contributor;
        }
    }
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        for (int ch = 0; ch < 3; ++ch) out_color[ch * H * W + pix_id] = C[ch] + T * bg[ch];
        out_depth[pix_id] = D;
    }
}

__global__ void __launch_bounds__(BSZ)
gsu_render_bwd(const uint2 *ranges, const uint32_t *point_list, int W, int H, const float *bg, const float2 *xy, const float4 *conic_o,
               const float *colors, const float *final_T, const uint32_t *n_contrib, const float *dL_dpix, float3 *dL_dmean2D,
               float4 *dL_dconic, float *dL_dopac, float *dL_dcol) {
    const int gxn = (W + BX - 1) / BX;
    const int pxi = blockIdx.x * BX + threadIdx.x, pyi = blockIdx.y * BY + threadIdx.y;
    const int pix_id = W * pyi + pxi;
    const float pxf = (float)pxi, pyf = (float)pyi;
    const bool inside = pxi < W && pyi < H;
    const uint2 range = ranges[blockIdx.y * gxn + blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BSZ - 1) / BSZ;
    bool done = !inside;
    int toDo = (int)(range.y - range.x);
    __shared__ int c_id[BSZ];
    __shared__ float2 c_xy[BSZ];
    __shared__ float4 c_co[BSZ];
    __shared__ float c_col[3 * BSZ];
    const int rank = threadIdx.y * BX + threadIdx.x;
    const float T_final = inside ? final_T[pix_id] : 0.f;
    float T = T_final;
    uint32_t contributor = (uint32_t)toDo;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
    float accum[3] = {0.f, 0.f, 0.f}, dLp[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
    if (inside)
        for (int ch = 0; ch < 3; ++ch) dLp[ch] = dL_dpix[ch * H * W + pix_id];
    float last_alpha = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    for (int i = 0; i < rounds; ++i, toDo -= BSZ) {
        __syncthreads();
        const int progress = i * BSZ + rank;
        if ((int)range.x + progress < (int)range.y) {
            const int id = (int)point_list[range.y - progress - 1];
            c_id[rank] = id; c_xy[rank] = xy[id]; c_co[rank] = conic_o[id];
            for (int ch = 0; ch < 3; ++ch) c_col[ch * BSZ + rank] = colors[id * 3 + ch];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BSZ, toDo); ++j) {
            --contributor;
            if ((int)contributor >= last_contributor) continue;
            const float2 c = c_xy[j];
            const float dx = c.x - pxf, dy = c.y - pyf;
            const float4 co = c_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float G = __expf(power);
            const float alpha = fminf(0.99f, co.w * G);
            if (alpha < 1.0f / 255.0f) continue;
            T = T / (1.f - alpha);
            const float dch = alpha * T;
            float dL_dalpha = 0.f;
            const int gid = c_id[j];
            for (int ch = 0; ch < 3; ++ch) {
                const float cc = c_col[ch * BSZ + j];
                accum[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum[ch];
                last_color[ch] = cc;
                dL_dalpha += (cc - accum[ch]) * dLp[ch];
                atomicAdd(&dL_dcol[gid * 3 + ch], dch * dLp[ch]);
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            float bgdot = 0.f;
            for (int ch = 0; ch < 3; ++ch) bgdot += bg[ch] * dLp[ch];
            dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co.x - gdy * co.y, dG_ddely = -gdy * co.z - gdx * co.y;
            atomicAdd(&dL_dmean2D[gid].x, dL_dG * dG_ddelx * ddelx_dx);
            atomicAdd(&dL_dmean2D[gid].y, dL_dG * dG_ddely * ddely_dy);
            atomicAdd(&dL_dconic[gid].x, -0.5f * gdx * dx * dL_dG);
            atomicAdd(&dL_dconic[gid].y, -0.5f * gdx * dy * dL_dG);
            atomicAdd(&dL_dconic[gid].w, -0.5f * gdy * dy * dL_dG);
            atomicAdd(&dL_dopac[gid], G * dL_dalpha);
        }
    }
}

// per-Gaussian backward, upstream's two kernels: (1) conic -> cov2D -> cov3D + mean (through J), (2) NDC mean2D -> mean3D, cov3D -> scale / quaternion
__global__ void gsu_cov2d_bwd(int P, const float *means3D, const int *radii, const float *cov3D, const float *V, float fx, float fy,
                              float tanx, float tany, const float4 *dL_dconic, float *dL_dmeans, float *dL_dcov) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || radii[i] <= 0) return;
    const float *c3 = cov3D + 6 * i;
    const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    const float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    const float3 t = mulpt(V, p);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float txtz = t.x / t.z, tytz = t.y / t.z;
    const float xg = (txtz < -limx || txtz > limx) ? 0.f : 1.f, yg = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    float M[2][3], tx, ty;
    proj_jac(V, t, fx, fy, tanx, tany, M, &tx, &ty);
    float MS[2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) MS[a][b] = M[a][0] * S[0][b] + M[a][1] * S[1][b] + M[a][2] * S[2][b];
    const float a_ = MS[0][0] * M[0][0] + MS[0][1] * M[0][1] + MS[0][2] * M[0][2] + 0.3f;
    const float b_ = MS[0][0] * M[1][0] + MS[0][1] * M[1][1] + MS[0][2] * M[1][2];
    const float c_ = MS[1][0] * M[1][0] + MS[1][1] * M[1][1] + MS[1][2] * M[1][2] + 0.3f;
    const float denom = a_ * c_ - b_ * b_;
    const float d2 = 1.0f / (denom * denom + 0.0000001f);
    const float4 dc = dL_dconic[i];
    float da = 0.f, db = 0.f, dcc = 0.f;
    if (d2 != 0.f) {
        da = d2 * (-c_ * c_ * dc.x + 2.f * b_ * c_ * dc.y + (denom - a_ * c_) * dc.w);
        dcc = d2 * (-a_ * a_ * dc.w + 2.f * a_ * b_ * dc.y + (denom - a_ * c_) * dc.x);
        db = d2 * 2.f * (b_ * c_ * dc.x - (denom + 2.f * b_ * b_) * dc.y + a_ * b_ * dc.w);
    }
    // dL/dcov3D (6 unique entries; off-diagonals carry both symmetric halves)
    float *o = dL_dcov + 6 * i;
    o[0] = M[0][0] * M[0][0] * da + M[0][0] * M[1][0] * db + M[1][0] * M[1][0] * dcc;
    o[3] = M[0][1] * M[0][1] * da + M[0][1] * M[1][1] * db + M[1][1] * M[1][1] * dcc;
    o[5] = M[0][2] * M[0][2] * da + M[0][2] * M[1][2] * db + M[1][2] * M[1][2] * dcc;
    o[1] = 2.f * M[0][0] * M[0][1] * da + (M[0][0] * M[1][1] + M[0][1] * M[1][0]) * db + 2.f * M[1][0] * M[1][1] * dcc;
    o[2] = 2.f * M[0][0] * M[0][2] * da + (M[0][0] * M[1][2] + M[0][2] * M[1][0]) * db + 2.f * M[1][0] * M[1][2] * dcc;
    o[4] = 2.f * M[0][2] * M[0][1] * da + (M[0][1] * M[1][2] + M[0][2] * M[1][1]) * db + 2.f * M[1][1] * M[1][2] * dcc;
    float dM[2][3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        dM[0][b] = 2.f * MS[0][b] * da + MS[1][b] * db;
        dM[1][b] = 2.f * MS[1][b] * dcc + MS[0][b] * db;
    }
    const float dJ00 = dM[0][0] * V[0] + dM[0][1] * V[4] + dM[0][2] * V[8];
    const float dJ02 = dM[0][0] * V[2] + dM[0][1] * V[6] + dM[0][2] * V[10];
    const float dJ11 = dM[1][0] * V[1] + dM[1][1] * V[5] + dM[1][2] * V[9];
    const float dJ12 = dM[1][0] * V[2] + dM[1][1] * V[6] + dM[1][2] * V[10];
    const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xg * -fx * tz2 * dJ02, dty = yg * -fy * tz2 * dJ12;
    const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * tx) * tz3 * dJ02 + (2.f * fy * ty) * tz3 * dJ12;
    dL_dmeans[3 * i] = V[0] * dtx + V[1] * dty + V[2] * dtz;
    dL_dmeans[3 * i + 1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
    dL_dmeans[3 * i + 2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
}

__global__ void gsu_pre_bwd(int P, const float *means3D, const int *radii, const float *scales, const float *rots, float mod, const float *Pm,
                            const float3 *dL_dmean2D, float *dL_dmeans, const float *dL_dcov, float *dL_dscale, float *dL_drot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || radii[i] <= 0) return;
    const float3 m = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    const float hw = Pm[3] * m.x + Pm[7] * m.y + Pm[11] * m.z + Pm[15];
    const float mw = 1.0f / (hw + 0.0000001f);
    const float mul1 = (Pm[0] * m.x + Pm[4] * m.y + Pm[8] * m.z + Pm[12]) * mw * mw;
    const float mul2 = (Pm[1] * m.x + Pm[5] * m.y + Pm[9] * m.z + Pm[13]) * mw * mw;
    const float gxm = dL_dmean2D[i].x, gym = dL_dmean2D[i].y;
    dL_dmeans[3 * i] += (Pm[0] * mw - Pm[3] * mul1) * gxm + (Pm[1] * mw - Pm[3] * mul2) * gym;
    dL_dmeans[3 * i + 1] += (Pm[4] * mw - Pm[7] * mul1) * gxm + (Pm[5] * mw - Pm[7] * mul2) * gym;
    dL_dmeans[3 * i + 2] += (Pm[8] * mw - Pm[11] * mul1) * gxm + (Pm[9] * mw - Pm[11] * mul2) * gym;
    // cov3D = R diag(s^2) R^T
    float R[3][3];
    rot_of(rots + 4 * i, R);
    const float *q = rots + 4 * i;
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
    const float *d = dL_dcov + 6 * i;
    const float Gs[3][3] = {{d[0], 0.5f * d[1], 0.5f * d[2]}, {0.5f * d[1], d[3], 0.5f * d[4]}, {0.5f * d[2], 0.5f * d[4], d[5]}};
    float GR[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) GR[a][b] = Gs[a][0] * R[0][b] + Gs[a][1] * R[1][b] + Gs[a][2] * R[2][b];
    float g[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float rgr = R[0][k] * GR[0][k] + R[1][k] * GR[1][k] + R[2][k] * GR[2][k];
        dL_dscale[3 * i + k] = 2.f * s[k] * rgr * mod;
#pragma unroll
        for (int a = 0; a < 3; ++a) g[a][k] = 2.f * GR[a][k] * s[k] * s[k];
    }
    dL_drot[4 * i] = 2.f * (-z * g[0][1] + y * g[0][2] + z * g[1][0] - x * g[1][2] - y * g[2][0] + x * g[2][1]);
    dL_drot[4 * i + 1] = 2.f * (y * g[0][1] + z * g[0][2] + y * g[1][0] - 2.f * x * g[1][1] - r * g[1][2] + z * g[2][0] + r * g[2][1] - 2.f * x * g[2][2]);
    dL_drot[4 * i + 2] = 2.f * (-2.f * y * g[0][0] + x * g[0][1] + r * g[0][2] + x * g[1][0] + z * g[1][2] - r * g[2][0] + z * g[2][1] - 2.f * y * g[2][2]);
    dL_drot[4 * i + 3] = 2.f * (-2.f * z * g[0][0] - r * g[0][1] + x * g[0][2] + r * g[1][0] - 2.f * z * g[1][1] + y * g[1][2] + x * g[2][0] + y * g[2][1]);
}

// ---------------------------------------------------------------------------------------------------------------------
struct GsuArgs {
    int32_t P, W, H;
    float tanfovx, tanfovy, scale_modifier;
    const float *means3D, *colors, *opacities, *scales, *rotations, *viewmatrix, *projmatrix, *bg;
    float *out_color, *out_depth;
    int32_t *radii;
    void *geom, *binning, *img;
};
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return -(int64_t)e_ - 1000; } while (0)

extern "C" void gsu_buffer_bytes(int32_t P, int32_t W, int32_t H, int64_t R, size_t out[3]) {
    GsuGeom g; GsuBin b; GsuImg im;
    out[0] = carve_geom(P, nullptr, &g); out[1] = carve_bin(R, nullptr, &b); out[2] = carve_img(W, H, nullptr, &im);
}

// phase 1: preprocess + inclusive scan + the D2H read of num_rendered (blocking, as upstream's forward does)
extern "C" int64_t gsu_forward_count(const GsuArgs *a, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GsuGeom g; carve_geom(a->P, (char *)a->geom, &g);
    const int gx = (a->W + BX - 1) / BX, gy = (a->H + BY - 1) / BY;
    const float fx = a->W / (2.f * a->tanfovx), fy = a->H / (2.f * a->tanfovy);
    if (a->P == 0) return 0;
    gsu_preprocess<<<(a->P + 255) / 256, 256, 0, st>>>(a->P, a->means3D, a->scales, a->scale_modifier, a->rotations, a->opacities, a->viewmatrix,
                                                        a->projmatrix, a->W, a->H, a->tanfovx, a->tanfovy, fx, fy, a->radii, g, gx, gy);
    CK(cub::DeviceScan::InclusiveSum(g.scan_tmp, g.scan_bytes, g.tiles, g.offsets, a->P, st));
    uint32_t R = 0;
    CK(cudaMemcpyAsync(&R, g.offsets + a->P - 1, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (int64_t)R;
}

// phase 2: keys, radix sort, ranges, blend
extern "C" int64_t gsu_forward_render(const GsuArgs *a, int64_t R, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GsuGeom g; carve_geom(a->P, (char *)a->geom, &g);
    GsuBin b; carve_bin(R, (char *)a->binning, &b);
    GsuImg im; carve_img(a->W, a->H, (char *)a->img, &im);
    const int gx = (a->W + BX - 1) / BX, gy = (a->H + BY - 1) / BY;
    if (a->P > 0 && R > 0) {
        gsu_duplicate<<<(a->P + 255) / 256, 256, 0, st>>>(a->P, g.xy, g.depths, g.offsets, g.radii_i, b.keys_u, b.vals_u, gx, gy);
        int bit = 0;
        for (uint32_t n = (uint32_t)(gx * gy); n > 0; n >>= 1) ++bit;   // bits needed for the tile id
        CK(cub::DeviceRadixSort::SortPairs(b.sort_tmp, b.sort_bytes, b.keys_u, b.keys, b.vals_u, b.vals, (int)R, 0, 32 + bit, st));
    }
    CK(cudaMemsetAsync(im.ranges, 0, (size_t)gx * gy * 8, st));
    if (R > 0) gsu_ranges<<<(int)((R + 255) / 256), 256, 0, st>>>((int)R, b.keys, im.ranges);
    gsu_render_fwd<<<dim3(gx, gy), dim3(BX, BY), 0, st>>>(im.ranges, b.vals, a->W, a->H, g.xy, a->colors, g.depths, g.conic_o, im.final_T,
                                                          im.n_contrib, a->bg, a->out_color, a->out_depth);
    CK(cudaGetLastError());
    return 0;
}

struct GsuGrads {
    const float *dL_dpix;
    float *dL_dmean2D /* [P,3] */, *dL_dconic /* [P,4] */, *dL_dopacity /* [P] */, *dL_dcolor /* [P,3] */, *dL_dmean3D /* [P,3] */,
        *dL_dcov3D /* [P,6] */, *dL_dscale /* [P,3] */, *dL_drot /* [P,4] */;   // all zero-initialised by the caller (torch.zeros), as upstream does
};
extern "C" int64_t gsu_backward(const GsuArgs *a, const GsuGrads *gr, int64_t R, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GsuGeom g; carve_geom(a->P, (char *)a->geom, &g);
    GsuBin b; carve_bin(R, (char *)a->binning, &b);
    GsuImg im; carve_img(a->W, a->H, (char *)a->img, &im);
    const int gx = (a->W + BX - 1) / BX, gy = (a->H + BY - 1) / BY;
    const float fx = a->W / (2.f * a->tanfovx), fy = a->H / (2.f * a->tanfovy);
    if (a->P == 0) return 0;
    gsu_render_bwd<<<dim3(gx, gy), dim3(BX, BY), 0, st>>>(im.ranges, b.vals, a->W, a->H, a->bg, g.xy, g.conic_o, a->colors, im.final_T, im.n_contrib,
                                                          gr->dL_dpix, (float3 *)gr->dL_dmean2D, (float4 *)gr->dL_dconic, gr->dL_dopacity, gr->dL_dcolor);
    gsu_cov2d_bwd<<<(a->P + 255) / 256, 256, 0, st>>>(a->P, a->means3D, g.radii_i, g.cov3D, a->viewmatrix, fx, fy, a->tanfovx, a->tanfovy,
                                                       (const float4 *)gr->dL_dconic, gr->dL_dmean3D, gr->dL_dcov3D);
    gsu_pre_bwd<<<(a->P + 255) / 256, 256, 0, st>>>(a->P, a->means3D, g.radii_i, a->scales, a->rotations, a->scale_modifier, a->projmatrix,
                                                     (const float3 *)gr->dL_dmean2D, gr->dL_dmean3D, gr->dL_dcov3D, gr->dL_dscale, gr->dL_drot);
    CK(cudaGetLastError());
    return 0;
}
