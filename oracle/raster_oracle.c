/*
 * oracle/raster_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32 arithmetic, OpenMP over tiles) of the tile-based
 * differentiable 3D-Gaussian rasterizer-with-depth that the reference calls through
 *   diff_gaussian_rasterization.GaussianRasterizer
 *   (/root/reference/src/tracking/train_utils.py:178,192; src/tracking/helpers.py:10-45;
 *    src/render/renderer.py:18-23).
 *
 * The algorithm itself lives in a third-party dependency that is NOT under /root/reference:
 *   JonathonLuiten/diff-gaussian-rasterization-w-depth, unpinned HEAD
 *   (/root/reference/README.md:26-35; .SUBMODULES.json lists no submodules).
 * It is restated here from its published algorithm: preprocess -> per-tile
 * (tile|depth-bits) stable sort -> front-to-back alpha blending with the exact constants
 * 0.2 / 1.3*tanfov / +0.3 / ceil(3*sqrt(lambda)) / 0.99 / 1/255 / 1e-4, and its backward
 * (back-to-front with T reconstructed by division, straight-through 0.99 clamp, 1e-7 guards,
 * depth output non-differentiable).
 *
 * PARITY UNPINNED: the reference ships no tests / golden images for this path; this file is
 * cross-checked against an independently written float64 autograd oracle
 * (oracle/raster_torch.py) in tests/test_oracle_raster.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

typedef struct {
    int G, W, H, CH;
    const float *means3D;   /* [G,3] */
    const float *colors;    /* [G,CH] */
    const float *opacities; /* [G]   */
    const float *scales;    /* [G,3] */
    const float *rotations; /* [G,4] (r,x,y,z), used as given (no renormalisation) */
    const float *viewmatrix;/* [16] = w2c^T flattened row-major (column-major w2c) */
    const float *projmatrix;/* [16] = (P*w2c)^T flattened */
    const float *bg;        /* [CH] */
    float tanfovx, tanfovy, scale_modifier;
} OracleIn;

/* per-Gaussian preprocess state kept for backward / debugging */
typedef struct {
    float *xy;        /* [G,2] pixel centre */
    float *conic_o;   /* [G,4] conic a,b,c + opacity */
    float *depth;     /* [G] */
    float *cov3D;     /* [G,6] */
    int   *radii;     /* [G] */
    int   *rect;      /* [G,4] minx,miny,maxx,maxy (tile units, max exclusive) */
    uint32_t *tiles_touched; /* [G] */
} OracleGeom;

static inline void xform4x3(const float *m, const float *p, float *o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static inline void xform4x4(const float *m, const float *p, float *o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* Sigma = R diag((mod*s)^2) R^T, stored as (xx,xy,xz,yy,yz,zz) */
static void cov3d_from_scale_rot(const float *s, float mod, const float *q, float *c6) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[3][3] = {
        {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
        {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
        {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    float sx = mod * s[0], sy = mod * s[1], sz = mod * s[2];
    /* M = S * R^T (rows of M = scaled columns of R); Sigma = M^T M */
    float M[3][3];
    for (int j = 0; j < 3; ++j) {
        M[0][j] = sx * R[j][0];
        M[1][j] = sy * R[j][1];
        M[2][j] = sz * R[j][2];
    }
    float S[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            S[i][j] = M[0][i] * M[0][j] + M[1][i] * M[1][j] + M[2][i] * M[2][j];
    c6[0] = S[0][0]; c6[1] = S[0][1]; c6[2] = S[0][2];
    c6[3] = S[1][1]; c6[4] = S[1][2]; c6[5] = S[2][2];
}

/* EWA projection: returns cov2D (a,b,c) with the +0.3 low-pass; also M = J*Rw (2x3) */
static void cov2d_project(const float *mean, float fx, float fy, float tfx, float tfy,
                          const float *c6, const float *V, float *abc, float Mo[2][3],
                          float *t_out, int *xin, int *yin) {
    float t[3];
    xform4x3(V, mean, t);
    float limx = 1.3f * tfx, limy = 1.3f * tfy;
    float txtz = t[0] / t[2], tytz = t[1] / t[2];
    *xin = !(txtz < -limx || txtz > limx);
    *yin = !(tytz < -limy || tytz > limy);
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    float J[2][3] = {{fx / t[2], 0.f, -(fx * t[0]) / (t[2] * t[2])},
                     {0.f, fy / t[2], -(fy * t[1]) / (t[2] * t[2])}};
    /* Rw[i][j] = w2c[i][j] = V[j*4+i] */
    float M[2][3];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j)
            M[i][j] = J[i][0] * V[j * 4 + 0] + J[i][1] * V[j * 4 + 1] + J[i][2] * V[j * 4 + 2];
    float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float MS[2][3];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j)
            MS[i][j] = M[i][0] * S[0][j] + M[i][1] * S[1][j] + M[i][2] * S[2][j];
    abc[0] = MS[0][0] * M[0][0] + MS[0][1] * M[0][1] + MS[0][2] * M[0][2] + 0.3f;
    abc[1] = MS[0][0] * M[1][0] + MS[0][1] * M[1][1] + MS[0][2] * M[1][2];
    abc[2] = MS[1][0] * M[1][0] + MS[1][1] * M[1][1] + MS[1][2] * M[1][2] + 0.3f;
    if (Mo) memcpy(Mo, M, sizeof(M));
    if (t_out) { t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2]; }
}

static void preprocess(const OracleIn *in, OracleGeom *g) {
    const int G = in->G, W = in->W, H = in->H;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float fx = W / (2.0f * in->tanfovx), fy = H / (2.0f * in->tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < G; ++i) {
        g->radii[i] = 0;
        g->tiles_touched[i] = 0;
        g->rect[4 * i] = g->rect[4 * i + 1] = g->rect[4 * i + 2] = g->rect[4 * i + 3] = 0;
        const float *p = in->means3D + 3 * i;
        float pv[3];
        xform4x3(in->viewmatrix, p, pv);
        if (pv[2] <= 0.2f) continue; /* near cull */
        float ph[4];
        xform4x4(in->projmatrix, p, ph);
        float pw = 1.0f / (ph[3] + 0.0000001f);
        float ndcx = ph[0] * pw, ndcy = ph[1] * pw;
        float *c6 = g->cov3D + 6 * i;
        cov3d_from_scale_rot(in->scales + 3 * i, in->scale_modifier, in->rotations + 4 * i, c6);
        float abc[3];
        int xi, yi;
        cov2d_project(p, fx, fy, in->tanfovx, in->tanfovy, c6, in->viewmatrix, abc, NULL, NULL, &xi, &yi);
        float det = abc[0] * abc[2] - abc[1] * abc[1];
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = {abc[2] * det_inv, -abc[1] * det_inv, abc[0] * det_inv};
        float mid = 0.5f * (abc[0] + abc[2]);
        float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float rad = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
        float px = ((ndcx + 1.0f) * W - 1.0f) * 0.5f;
        float py = ((ndcy + 1.0f) * H - 1.0f) * 0.5f;
        int minx = (int)((px - rad) / TILE), miny = (int)((py - rad) / TILE);
        int maxx = (int)((px + rad + TILE - 1) / TILE), maxy = (int)((py + rad + TILE - 1) / TILE);
        minx = minx < 0 ? 0 : (minx > gx ? gx : minx);
        miny = miny < 0 ? 0 : (miny > gy ? gy : miny);
        maxx = maxx < 0 ? 0 : (maxx > gx ? gx : maxx);
        maxy = maxy < 0 ? 0 : (maxy > gy ? gy : maxy);
        if ((maxx - minx) * (maxy - miny) == 0) continue;
        g->depth[i] = pv[2];
        g->radii[i] = (int)rad;
        g->xy[2 * i] = px; g->xy[2 * i + 1] = py;
        g->conic_o[4 * i] = conic[0]; g->conic_o[4 * i + 1] = conic[1];
        g->conic_o[4 * i + 2] = conic[2]; g->conic_o[4 * i + 3] = in->opacities[i];
        g->rect[4 * i] = minx; g->rect[4 * i + 1] = miny; g->rect[4 * i + 2] = maxx; g->rect[4 * i + 3] = maxy;
        g->tiles_touched[i] = (uint32_t)((maxx - minx) * (maxy - miny));
    }
}

typedef struct { uint64_t key; uint32_t id; } KV;
static int kv_cmp(const void *a, const void *b) {
    const KV *x = (const KV *)a, *y = (const KV *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id); /* stable: ties by Gaussian index */
}

/* builds sorted instance list + per-tile ranges; returns R */
static int64_t bin_and_sort(const OracleIn *in, const OracleGeom *g, KV **list_out, int64_t **ranges_out) {
    const int G = in->G;
    const int gx = (in->W + TILE - 1) / TILE, gy = (in->H + TILE - 1) / TILE;
    int64_t R = 0;
    for (int i = 0; i < G; ++i) R += g->tiles_touched[i];
    KV *list = (KV *)malloc(sizeof(KV) * (size_t)(R > 0 ? R : 1));
    int64_t off = 0;
    for (int i = 0; i < G; ++i) {
        if (g->radii[i] <= 0) continue;
        uint32_t dbits;
        memcpy(&dbits, &g->depth[i], 4);
        for (int y = g->rect[4 * i + 1]; y < g->rect[4 * i + 3]; ++y)
            for (int x = g->rect[4 * i]; x < g->rect[4 * i + 2]; ++x) {
                list[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                list[off].id = (uint32_t)i;
                ++off;
            }
    }
    qsort(list, (size_t)R, sizeof(KV), kv_cmp);
    int64_t *ranges = (int64_t *)calloc((size_t)gx * gy * 2, sizeof(int64_t));
    for (int64_t j = 0; j < R; ++j) {
        uint32_t t = (uint32_t)(list[j].key >> 32);
        if (j == 0 || t != (uint32_t)(list[j - 1].key >> 32)) ranges[2 * t] = j;
        if (j == R - 1 || t != (uint32_t)(list[j + 1].key >> 32)) ranges[2 * t + 1] = j + 1;
    }
    *list_out = list;
    *ranges_out = ranges;
    return R;
}

#define MAXCH 8

/*
 * Forward. Outputs: out_color [CH,H,W], out_depth [H,W], radii [G], final_T [H,W], n_contrib [H,W].
 * Optional debug outputs (may be NULL): xy [G,2], conic_o [G,4], depth [G], tiles_touched [G].
 * Returns R (number of tile instances) or <0 on error.
 */
int64_t gsd_oracle_raster_forward(const OracleIn *in, float *out_color, float *out_depth, int *radii,
                                  float *final_T, int *n_contrib, float *dbg_xy, float *dbg_conic_o,
                                  float *dbg_depth, uint32_t *dbg_tiles) {
    const int G = in->G, W = in->W, H = in->H, CH = in->CH;
    if (CH > MAXCH || G < 0) return -1;
    OracleGeom g;
    g.xy = (float *)calloc((size_t)G * 2 + 1, 4);
    g.conic_o = (float *)calloc((size_t)G * 4 + 1, 4);
    g.depth = (float *)calloc((size_t)G + 1, 4);
    g.cov3D = (float *)calloc((size_t)G * 6 + 1, 4);
    g.radii = radii;
    g.rect = (int *)calloc((size_t)G * 4 + 1, 4);
    g.tiles_touched = (uint32_t *)calloc((size_t)G + 1, 4);
    preprocess(in, &g);
    KV *list; int64_t *ranges;
    int64_t R = bin_and_sort(in, &g, &list, &ranges);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        int64_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                float T = 1.0f, C[MAXCH] = {0}, D = 0.f;
                int contributor = 0, last = 0;
                for (int64_t j = s; j < e; ++j) {
                    ++contributor;
                    uint32_t id = list[j].id;
                    float dx = g.xy[2 * id] - (float)px, dy = g.xy[2 * id + 1] - (float)py;
                    const float *co = g.conic_o + 4 * id;
                    float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break; /* done */
                    for (int c = 0; c < CH; ++c) C[c] += in->colors[(size_t)id * CH + c] * alpha * T;
                    D += g.depth[id] * alpha * T;
                    T = test_T;
                    last = contributor;
                }
                size_t pid = (size_t)py * W + px;
                final_T[pid] = T;
                n_contrib[pid] = last;
                for (int c = 0; c < CH; ++c) out_color[(size_t)c * H * W + pid] = C[c] + T * in->bg[c];
                out_depth[pid] = D;
            }
    }
    if (dbg_xy) memcpy(dbg_xy, g.xy, (size_t)G * 8);
    if (dbg_conic_o) memcpy(dbg_conic_o, g.conic_o, (size_t)G * 16);
    if (dbg_depth) memcpy(dbg_depth, g.depth, (size_t)G * 4);
    if (dbg_tiles) memcpy(dbg_tiles, g.tiles_touched, (size_t)G * 4);
    free(list); free(ranges);
    free(g.xy); free(g.conic_o); free(g.depth); free(g.cov3D); free(g.rect); free(g.tiles_touched);
    return R;
}

/*
 * Backward. dL_dcolor [CH,H,W] in; outputs (all zero-initialised here):
 *   dmeans3D [G,3], dmeans2D [G,3] (NDC-scaled xy, z=0), dcolors [G,CH], dopacity [G], dscales [G,3], drot [G,4]
 * Per-Gaussian sums over pixels are accumulated in float64 (the reference uses fp32 atomics in
 * arbitrary order; float64 removes that noise from the oracle).
 */
int64_t gsd_oracle_raster_backward(const OracleIn *in, const float *dL_dcolor, float *dmeans3D,
                                   float *dmeans2D, float *dcolors, float *dopacity, float *dscales,
                                   float *drot) {
    const int G = in->G, W = in->W, H = in->H, CH = in->CH;
    if (CH > MAXCH) return -1;
    OracleGeom g;
    g.xy = (float *)calloc((size_t)G * 2 + 1, 4);
    g.conic_o = (float *)calloc((size_t)G * 4 + 1, 4);
    g.depth = (float *)calloc((size_t)G + 1, 4);
    g.cov3D = (float *)calloc((size_t)G * 6 + 1, 4);
    g.radii = (int *)calloc((size_t)G + 1, 4);
    g.rect = (int *)calloc((size_t)G * 4 + 1, 4);
    g.tiles_touched = (uint32_t *)calloc((size_t)G + 1, 4);
    preprocess(in, &g);
    KV *list; int64_t *ranges;
    int64_t R = bin_and_sort(in, &g, &list, &ranges);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    /* accumulators: mean2D(2) conic(3) opacity(1) colors(CH) */
    const int NA = 6 + CH;
    double *acc = (double *)calloc((size_t)G * NA + 1, sizeof(double));

#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        int64_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
        int n = (int)(e - s);
        if (n == 0) continue;
        double *tacc = (double *)calloc((size_t)n * NA, sizeof(double));
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                size_t pid = (size_t)py * W + px;
                /* replay forward to get final_T / last contributor (same fp32 ops) */
                float T = 1.0f;
                int contributor = 0, last = 0;
                for (int64_t j = s; j < e; ++j) {
                    ++contributor;
                    uint32_t id = list[j].id;
                    float dx = g.xy[2 * id] - (float)px, dy = g.xy[2 * id + 1] - (float)py;
                    const float *co = g.conic_o + 4 * id;
                    float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break;
                    T = test_T;
                    last = contributor;
                }
                const float T_final = T;
                float dLp[MAXCH], accum_rec[MAXCH] = {0}, last_color[MAXCH] = {0};
                float bg_dot = 0.f;
                for (int c = 0; c < CH; ++c) {
                    dLp[c] = dL_dcolor[(size_t)c * H * W + pid];
                    bg_dot += in->bg[c] * dLp[c];
                }
                float last_alpha = 0.f;
                const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
                /* back-to-front over the contributors [0,last) */
                for (int k = last - 1; k >= 0; --k) {
                    uint32_t id = list[s + k].id;
                    float dx = g.xy[2 * id] - (float)px, dy = g.xy[2 * id + 1] - (float)py;
                    const float *co = g.conic_o + 4 * id;
                    float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    float Gv = expf(power);
                    float alpha = fminf(0.99f, co[3] * Gv);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.f;
                    double *a = tacc + (size_t)k * NA;
                    for (int c = 0; c < CH; ++c) {
                        float col = in->colors[(size_t)id * CH + c];
                        accum_rec[c] = last_alpha * last_color[c] + (1.f - last_alpha) * accum_rec[c];
                        last_color[c] = col;
                        dL_dalpha += (col - accum_rec[c]) * dLp[c];
                        a[6 + c] += (double)(dchannel_dcolor * dLp[c]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    float dL_dG = co[3] * dL_dalpha;
                    float gdx = Gv * dx, gdy = Gv * dy;
                    float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    float dG_ddely = -gdy * co[2] - gdx * co[1];
                    a[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                    a[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                    a[2] += (double)(-0.5f * gdx * dx * dL_dG);
                    a[3] += (double)(-0.5f * gdx * dy * dL_dG); /* half of d/d(conic.b): see cov2D backward */
                    a[4] += (double)(-0.5f * gdy * dy * dL_dG);
                    a[5] += (double)(Gv * dL_dalpha);
                }
            }
        for (int k = 0; k < n; ++k) {
            uint32_t id = list[s + k].id;
            for (int v = 0; v < NA; ++v) {
                double val = tacc[(size_t)k * NA + v];
                if (val != 0.0) {
#pragma omp atomic
                    acc[(size_t)id * NA + v] += val;
                }
            }
        }
        free(tacc);
    }

    const float fx = W / (2.0f * in->tanfovx), fy = H / (2.0f * in->tanfovy);
    const float *V = in->viewmatrix, *P = in->projmatrix;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < G; ++i) {
        for (int c = 0; c < 3; ++c) { dmeans3D[3 * i + c] = 0.f; dmeans2D[3 * i + c] = 0.f; dscales[3 * i + c] = 0.f; }
        for (int c = 0; c < 4; ++c) drot[4 * i + c] = 0.f;
        for (int c = 0; c < CH; ++c) dcolors[(size_t)i * CH + c] = 0.f;
        dopacity[i] = 0.f;
        if (!(g.radii[i] > 0)) continue;
        const double *a = acc + (size_t)i * NA;
        float dm2x = (float)a[0], dm2y = (float)a[1];
        float dcon[3] = {(float)a[2], (float)a[3], (float)a[4]};
        dopacity[i] = (float)a[5];
        for (int c = 0; c < CH; ++c) dcolors[(size_t)i * CH + c] = (float)a[6 + c];
        dmeans2D[3 * i] = dm2x; dmeans2D[3 * i + 1] = dm2y;

        /* ---- cov2D backward: dL/dconic -> dL/dcov2D(a,b,c) -> dL/dSigma, dL/dmean (through J) ---- */
        const float *p = in->means3D + 3 * i;
        const float *c6 = g.cov3D + 6 * i;
        float abc[3], M[2][3], t[3];
        int xin, yin;
        cov2d_project(p, fx, fy, in->tanfovx, in->tanfovy, c6, V, abc, M, t, &xin, &yin);
        float ca = abc[0], cb = abc[1], cc = abc[2];
        float denom = ca * cc - cb * cb;
        float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        float dS[6] = {0, 0, 0, 0, 0, 0};
        if (denom2inv != 0) {
            dL_da = denom2inv * (-cc * cc * dcon[0] + 2 * cb * cc * dcon[1] + (denom - ca * cc) * dcon[2]);
            dL_dc = denom2inv * (-ca * ca * dcon[2] + 2 * ca * cb * dcon[1] + (denom - ca * cc) * dcon[0]);
            dL_db = denom2inv * 2 * (cb * cc * dcon[0] - (denom + 2 * cb * cb) * dcon[1] + ca * cb * dcon[2]);
            /* dL/dSigma_ij (unique entries; off-diagonals carry both symmetric halves) */
            dS[0] = M[0][0] * M[0][0] * dL_da + M[0][0] * M[1][0] * dL_db + M[1][0] * M[1][0] * dL_dc;
            dS[3] = M[0][1] * M[0][1] * dL_da + M[0][1] * M[1][1] * dL_db + M[1][1] * M[1][1] * dL_dc;
            dS[5] = M[0][2] * M[0][2] * dL_da + M[0][2] * M[1][2] * dL_db + M[1][2] * M[1][2] * dL_dc;
            dS[1] = 2 * M[0][0] * M[0][1] * dL_da + (M[0][0] * M[1][1] + M[0][1] * M[1][0]) * dL_db + 2 * M[1][0] * M[1][1] * dL_dc;
            dS[2] = 2 * M[0][0] * M[0][2] * dL_da + (M[0][0] * M[1][2] + M[0][2] * M[1][0]) * dL_db + 2 * M[1][0] * M[1][2] * dL_dc;
            dS[4] = 2 * M[0][2] * M[0][1] * dL_da + (M[0][1] * M[1][2] + M[0][2] * M[1][1]) * dL_db + 2 * M[1][1] * M[1][2] * dL_dc;
        }
        float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        /* dL/dM = 2*Gm*M*Sigma with Gm = [[da, db/2],[db/2, dc]] */
        float MS[2][3];
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c) MS[r][c] = M[r][0] * S[0][c] + M[r][1] * S[1][c] + M[r][2] * S[2][c];
        float dM[2][3];
        for (int c = 0; c < 3; ++c) {
            dM[0][c] = 2 * MS[0][c] * dL_da + MS[1][c] * dL_db;
            dM[1][c] = 2 * MS[1][c] * dL_dc + MS[0][c] * dL_db;
        }
        /* M = J * Rw  => dL/dJ = dL/dM * Rw^T ; Rw[i][j] = V[j*4+i] */
        float dJ00 = dM[0][0] * V[0] + dM[0][1] * V[4] + dM[0][2] * V[8];
        float dJ02 = dM[0][0] * V[2] + dM[0][1] * V[6] + dM[0][2] * V[10];
        float dJ11 = dM[1][0] * V[1] + dM[1][1] * V[5] + dM[1][2] * V[9];
        float dJ12 = dM[1][0] * V[2] + dM[1][1] * V[6] + dM[1][2] * V[10];
        float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        float dtx = (xin ? 1.f : 0.f) * -fx * tz2 * dJ02;
        float dty = (yin ? 1.f : 0.f) * -fy * tz2 * dJ12;
        float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
        /* t = Rw p + tw => dL/dp = Rw^T dL/dt */
        float dmean[3];
        dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
        dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
        dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;

        /* ---- projection backward: dL/dmean2D (NDC) -> dL/dmean3D ---- */
        float mh[4];
        xform4x4(P, p, mh);
        float mw = 1.0f / (mh[3] + 0.0000001f);
        float mul1 = (P[0] * p[0] + P[4] * p[1] + P[8] * p[2] + P[12]) * mw * mw;
        float mul2 = (P[1] * p[0] + P[5] * p[1] + P[9] * p[2] + P[13]) * mw * mw;
        dmean[0] += (P[0] * mw - P[3] * mul1) * dm2x + (P[1] * mw - P[3] * mul2) * dm2y;
        dmean[1] += (P[4] * mw - P[7] * mul1) * dm2x + (P[5] * mw - P[7] * mul2) * dm2y;
        dmean[2] += (P[8] * mw - P[11] * mul1) * dm2x + (P[9] * mw - P[11] * mul2) * dm2y;
        dmeans3D[3 * i] = dmean[0]; dmeans3D[3 * i + 1] = dmean[1]; dmeans3D[3 * i + 2] = dmean[2];

        /* ---- cov3D backward: Sigma = R D R^T, D = diag((mod*s)^2) ---- */
        const float *q = in->rotations + 4 * i;
        const float *sc = in->scales + 3 * i;
        float r = q[0], x = q[1], y = q[2], z = q[3];
        float Rm[3][3] = {
            {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
            {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
            {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        float mod = in->scale_modifier;
        float sv[3] = {mod * sc[0], mod * sc[1], mod * sc[2]};
        /* symmetric gradient matrix: off-diagonals halved */
        float Gs[3][3] = {{dS[0], 0.5f * dS[1], 0.5f * dS[2]},
                          {0.5f * dS[1], dS[3], 0.5f * dS[4]},
                          {0.5f * dS[2], 0.5f * dS[4], dS[5]}};
        float GR[3][3]; /* Gs * R */
        for (int a_ = 0; a_ < 3; ++a_)
            for (int b_ = 0; b_ < 3; ++b_)
                GR[a_][b_] = Gs[a_][0] * Rm[0][b_] + Gs[a_][1] * Rm[1][b_] + Gs[a_][2] * Rm[2][b_];
        float g_[3][3]; /* dL/dR = 2 * Gs * R * D */
        for (int k = 0; k < 3; ++k) {
            float rgr = Rm[0][k] * GR[0][k] + Rm[1][k] * GR[1][k] + Rm[2][k] * GR[2][k]; /* (R^T Gs R)_kk */
            dscales[3 * i + k] = 2.f * sv[k] * rgr * mod;
            for (int a_ = 0; a_ < 3; ++a_) g_[a_][k] = 2.f * GR[a_][k] * sv[k] * sv[k];
        }
        drot[4 * i + 0] = 2.f * (-z * g_[0][1] + y * g_[0][2] + z * g_[1][0] - x * g_[1][2] - y * g_[2][0] + x * g_[2][1]);
        drot[4 * i + 1] = 2.f * (y * g_[0][1] + z * g_[0][2] + y * g_[1][0] - 2.f * x * g_[1][1] - r * g_[1][2] + z * g_[2][0] + r * g_[2][1] - 2.f * x * g_[2][2]);
        drot[4 * i + 2] = 2.f * (-2.f * y * g_[0][0] + x * g_[0][1] + r * g_[0][2] + x * g_[1][0] + z * g_[1][2] - r * g_[2][0] + z * g_[2][1] - 2.f * y * g_[2][2]);
        drot[4 * i + 3] = 2.f * (-2.f * z * g_[0][0] - r * g_[0][1] + x * g_[0][2] + r * g_[1][0] - 2.f * z * g_[1][1] + y * g_[1][2] + x * g_[2][0] + y * g_[2][1]);
    }
    free(acc); free(list); free(ranges);
    free(g.xy); free(g.conic_o); free(g.depth); free(g.cov3D); free(g.radii); free(g.rect); free(g.tiles_touched);
    return R;
}

int gsd_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the timed CPU arm asks for the host's cores explicitly */
void gsd_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
