// skinning.cu — linear-blend skinning of the Gaussians from the GNN particles ("bones"), sm_100a.
//
//   gsd_skin_bone_transforms   per bone: Procrustes rotation of its graph neighbourhood (3x3 SVD), rest-pose affine, quaternion
//                              replaces the Python loop over bones of interpolate_motions, /root/reference/src/render/utils.py:150-204
//   gsd_skin_apply             per Gaussian: inverse-distance weights over ALL bones, blended position and rotation
//                              replaces utils.py:206-239 (cdist + [n_particles, n_bones, 3] temporaries + per-bone Python loop)
//
// The reference materialises weights [n_particles, n_bones] and xyz_transformed [n_particles, n_bones, 3]
// (100k x 2k: 0.8 GB + 2.4 GB); here a particle streams the bone table through shared memory once and keeps 8 accumulators,
// so the algorithmic bytes are 28·n_particles in + 28·n_particles out + 80·n_bones: the kernel is FP32-issue bound
// (≈ 27 instructions per particle-bone pair), not HBM bound.
#include "common.cuh"
#include <cfloat>

#define SKIN_TF 20      // floats per bone record: R (9, row-major) | c = t + b - R b (3) | q (4) | b (3) | pad
#define SKIN_TILE 256   // bones per shared-memory tile
#define SKIN_PPT 2      // particles per thread

// ------------------------------------------------------------------------------------------------------
// 3x3 SVD by one-sided Jacobi in double (the matrix itself is the reference's fp32 F).  A = U diag(S) V^T, S descending.
// Columns of U belonging to zero singular values are left zero; the caller completes what it needs.
// ------------------------------------------------------------------------------------------------------
__device__ void svd3(const double F[9], double U[9], double S[3], double V[9]) {
    double a[3][3]; // a[j] = column j of the working matrix
    double v[3][3]; // v[j] = column j of V
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) { a[j][i] = F[3 * i + j]; v[j][i] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double alpha = a[p][0] * a[p][0] + a[p][1] * a[p][1] + a[p][2] * a[p][2];
            const double beta = a[q][0] * a[q][0] + a[q][1] * a[q][1] + a[q][2] * a[q][2];
            const double gamma = a[p][0] * a[q][0] + a[p][1] * a[q][1] + a[p][2] * a[q][2];
            if (fabs(gamma) > 1e-300 && fabs(gamma) > 1e-15 * sqrt(alpha * beta)) {
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double ap = a[p][i], aq = a[q][i];
                    a[p][i] = c * ap - s * aq;
                    a[q][i] = s * ap + c * aq;
                    const double vp = v[p][i], vq = v[q][i];
                    v[p][i] = c * vp - s * vq;
                    v[q][i] = s * vp + c * vq;
                }
            }
        }
        if (!rotated) break;
    }
    double n[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) n[j] = sqrt(a[j][0] * a[j][0] + a[j][1] * a[j][1] + a[j][2] * a[j][2]);
    int o0 = 0, o1 = 1, o2 = 2; // sort descending
    if (n[o0] < n[o1]) { int t = o0; o0 = o1; o1 = t; }
    if (n[o1] < n[o2]) { int t = o1; o1 = o2; o2 = t; }
    if (n[o0] < n[o1]) { int t = o0; o0 = o1; o1 = t; }
    const int ord[3] = {o0, o1, o2};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = ord[j];
        S[j] = n[k];
        const double inv = n[k] > 0.0 ? 1.0 / n[k] : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { U[3 * i + j] = a[k][i] * inv; V[3 * i + j] = v[k][i]; }
    }
}

__device__ __forceinline__ double det3(const double M[9]) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// mat2quat of /root/reference/src/render/utils.py:68-109 (w, x, y, z), all four branches
__device__ void mat2quat_ref(const float R[9], float q[4]) {
    const float r00 = R[0], r01 = R[1], r02 = R[2], r10 = R[3], r11 = R[4], r12 = R[5], r20 = R[6], r21 = R[7], r22 = R[8];
    const float t = fmaxf(r00 + r11 + r22, -1.f);
    if (t > -1.f) {
        float s = sqrtf(t + 1.f);
        q[0] = 0.5f * s;
        s = 0.5f / s;
        q[1] = (r21 - r12) * s; q[2] = (r02 - r20) * s; q[3] = (r10 - r01) * s;
    } else if (r00 >= r11 && r00 >= r22) {
        const float s = 0.5f / sqrtf(1.f + r00 - r11 - r22);
        q[0] = (r21 - r12) * s; q[1] = 0.5f * s; q[2] = (r10 + r01) * s; q[3] = (r20 + r02) * s;
    } else if (r11 >= r22 && r11 > r00) {
        const float s = 0.5f / sqrtf(1.f + r11 - r00 - r22);
        q[0] = (r02 - r20) * s; q[1] = (r21 + r12) * s; q[2] = 0.5f * s; q[3] = (r01 + r10) * s;
    } else {
        const float s = 0.5f / sqrtf(1.f + r22 - r00 - r11);
        q[0] = (r10 - r01) * s; q[1] = (r02 + r20) * s; q[2] = (r12 + r21) * s; q[3] = 0.5f * s;
    }
}

// ------------------------------------------------------------------------------------------------------
// one thread per bone.  Neighbours = the columns of the bone's CSR row that are bones themselves (col < n_bones; the tool
// node of the rollout graph is dropped exactly like relations[:nobj, :nobj] in dynamics_module.py:154).
//   F = sum_j (b_j + m_j - b_i - m_i)(b_j - b_i)^T            (utils.py:162-169, W = identity)
//   rank 0 / no neighbour: R = I;  rank 1: rotation taking the x axis onto U[:,0] about their common normal (utils.py:171-186),
//   with LAPACK's sign of U[:,0] (x component <= 0);  rank >= 2: U diag(1,1,±1) V^T with det +1 (utils.py:188-202), except the
//   reference's rank-3 / det(F) < 0 case, where `S[3,3] = -1` raises inside its try block and R falls back to the identity.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gsd_skin_bone_kernel(int n_bones, const float *__restrict__ bones, const float *__restrict__ motions, const int32_t *__restrict__ row_ptr,
                     const int32_t *__restrict__ cols, float *__restrict__ tf, float *__restrict__ rot_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bones) return;
    const float bx = bones[3 * i], by = bones[3 * i + 1], bz = bones[3 * i + 2];
    const float mx = motions[3 * i], my = motions[3 * i + 1], mz = motions[3 * i + 2];
    float Ff[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int n_adj = 0;
    for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
        const int j = cols[e];
        if (j < 0 || j >= n_bones) continue;
        ++n_adj;
        const float ox = bones[3 * j] - bx, oy = bones[3 * j + 1] - by, oz = bones[3 * j + 2] - bz;
        const float nx = (bones[3 * j] + motions[3 * j]) - (bx + mx), ny = (bones[3 * j + 1] + motions[3 * j + 1]) - (by + my),
                    nz = (bones[3 * j + 2] + motions[3 * j + 2]) - (bz + mz);
        Ff[0] += nx * ox; Ff[1] += nx * oy; Ff[2] += nx * oz;
        Ff[3] += ny * ox; Ff[4] += ny * oy; Ff[5] += ny * oz;
        Ff[6] += nz * ox; Ff[7] += nz * oy; Ff[8] += nz * oz;
    }
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (n_adj > 0) {
        double F[9], U[9], S[3], V[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) F[k] = (double)Ff[k];
        svd3(F, U, S, V);
        const double tol = 3.0 * (double)FLT_EPSILON * S[0]; // torch.linalg.matrix_rank default for a 3x3 fp32 matrix
        const int rank = (S[0] > tol ? 1 : 0) + (S[1] > tol ? 1 : 0) + (S[2] > tol ? 1 : 0);
        if (rank == 1) {
            double ax = U[0], ay = U[3], az = U[6];
            if (ax > 0.0) { ax = -ax; ay = -ay; az = -az; } // LAPACK's sign of the first left singular vector
            // perp = axis x e_x = (0, az, -ay)
            const double pn = sqrt(az * az + ay * ay);
            if (pn >= 1e-6) {
                const double py = az / pn, pz = -ay / pn;
                // third = e_x x perp = (0, -pz, py);  third_after = axis x perp
                const double tx = 0.0, ty = -pz, tz = py;
                const double wx = ay * pz - az * py, wy = az * 0.0 - ax * pz, wz = ax * py - ay * 0.0;
                // R = Y X^T,  X = [e_x | perp | third],  Y = [axis | perp | third_after]  (columns)
                const double X[9] = {1.0, 0.0, tx, 0.0, py, ty, 0.0, pz, tz};
                const double Y[9] = {ax, 0.0, wx, ay, py, wy, az, pz, wz};
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        R[3 * r + c] = (float)(Y[3 * r] * X[3 * c] + Y[3 * r + 1] * X[3 * c + 1] + Y[3 * r + 2] * X[3 * c + 2]);
            }
        } else if (rank >= 2) {
            if (rank == 3 && det3(F) < 0.0) {
                // reference: S[cov_rank, cov_rank] = -1 with cov_rank = 3 raises, caught -> identity (utils.py:190-196)
            } else {
                if (rank == 2) { // complete U and V with the normals of their first two columns
                    U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
                    const double un = sqrt(U[2] * U[2] + U[5] * U[5] + U[8] * U[8]);
                    if (un > 0.0) { U[2] /= un; U[5] /= un; U[8] /= un; }
                }
                const double d = det3(U) * det3(V) < 0.0 ? -1.0 : 1.0;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        R[3 * r + c] = (float)(U[3 * r] * V[3 * c] + U[3 * r + 1] * V[3 * c + 1] + d * U[3 * r + 2] * V[3 * c + 2]);
            }
        }
    }
    float q[4];
    mat2quat_ref(R, q);
    const float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f); // F.normalize
    float *o = tf + (size_t)i * SKIN_TF;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = R[k];
    o[9] = mx + bx - (R[0] * bx + R[1] * by + R[2] * bz);
    o[10] = my + by - (R[3] * bx + R[4] * by + R[5] * bz);
    o[11] = mz + bz - (R[6] * bx + R[7] * by + R[8] * bz);
    o[12] = q[0] / qn; o[13] = q[1] / qn; o[14] = q[2] / qn; o[15] = q[3] / qn;
    o[16] = bx; o[17] = by; o[18] = bz; o[19] = 0.f;
    if (rot_out) {
#pragma unroll
        for (int k = 0; k < 9; ++k) rot_out[(size_t)i * 9 + k] = R[k];
    }
}

// ------------------------------------------------------------------------------------------------------
// SKIN_PPT particles per thread; the bone table streams through shared memory in tiles (every lane reads the same record:
// broadcast LDS.128).  w = 1 / max(|x - b|, 1e-4)  (utils.py:210-213), or the caller's weight row.
//   xyz' = sum_b w_b (R_b x + c_b) / sum_b w_b,   q' = normalize(sum_b w_b q_b) (x) q      (utils.py:216-237)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gsd_skin_apply_kernel(int n_particles, int n_bones, const float *__restrict__ xyz, const float *__restrict__ quat,
                      const float *__restrict__ tf, const float *__restrict__ weights_in, float *__restrict__ xyz_out,
                      float *__restrict__ quat_out, float *__restrict__ weights_out) {
    __shared__ float4 stf[SKIN_TILE * (SKIN_TF / 4)];
    const int base = (blockIdx.x * blockDim.x + threadIdx.x) * SKIN_PPT;
    float px[SKIN_PPT], py[SKIN_PPT], pz[SKIN_PPT];
    float sx[SKIN_PPT], sy[SKIN_PPT], sz[SKIN_PPT], sw[SKIN_PPT], q0[SKIN_PPT], q1[SKIN_PPT], q2[SKIN_PPT], q3[SKIN_PPT];
#pragma unroll
    for (int u = 0; u < SKIN_PPT; ++u) {
        const int p = min(base + u, n_particles - 1);
        px[u] = xyz[3 * p]; py[u] = xyz[3 * p + 1]; pz[u] = xyz[3 * p + 2];
        sx[u] = sy[u] = sz[u] = sw[u] = q0[u] = q1[u] = q2[u] = q3[u] = 0.f;
    }
    for (int b0 = 0; b0 < n_bones; b0 += SKIN_TILE) {
        const int nb = min(SKIN_TILE, n_bones - b0);
        __syncthreads();
        for (int k = threadIdx.x; k < nb * (SKIN_TF / 4); k += blockDim.x) stf[k] = reinterpret_cast<const float4 *>(tf + (size_t)b0 * SKIN_TF)[k];
        __syncthreads();
#pragma unroll 2
        for (int b = 0; b < nb; ++b) {
            const float4 r0 = stf[b * 5], r1 = stf[b * 5 + 1], r2 = stf[b * 5 + 2], r3 = stf[b * 5 + 3], r4 = stf[b * 5 + 4];
            // r0 = (R00 R01 R02 R10)  r1 = (R11 R12 R20 R21)  r2 = (R22 cx cy cz)  r3 = q  r4 = (bx by bz -)
#pragma unroll
            for (int u = 0; u < SKIN_PPT; ++u) {
                float w;
                if (weights_in) {
                    w = weights_in[(size_t)min(base + u, n_particles - 1) * n_bones + b0 + b];
                } else {
                    const float dx = px[u] - r4.x, dy = py[u] - r4.y, dz = pz[u] - r4.z;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    w = rsqrtf(fmaxf(d2, 1e-8f)); // 1 / clamp(dist, min = 1e-4)
                }
                const float yx = fmaf(r0.x, px[u], fmaf(r0.y, py[u], fmaf(r0.z, pz[u], r2.y)));
                const float yy = fmaf(r0.w, px[u], fmaf(r1.x, py[u], fmaf(r1.y, pz[u], r2.z)));
                const float yz = fmaf(r1.z, px[u], fmaf(r1.w, py[u], fmaf(r2.x, pz[u], r2.w)));
                sx[u] = fmaf(w, yx, sx[u]); sy[u] = fmaf(w, yy, sy[u]); sz[u] = fmaf(w, yz, sz[u]);
                q0[u] = fmaf(w, r3.x, q0[u]); q1[u] = fmaf(w, r3.y, q1[u]); q2[u] = fmaf(w, r3.z, q2[u]); q3[u] = fmaf(w, r3.w, q3[u]);
                sw[u] += w;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < SKIN_PPT; ++u) {
        const int p = base + u;
        if (p >= n_particles) continue;
        const float inv = weights_in ? 1.f : 1.f / sw[u]; // caller-supplied weights are used as they are
        xyz_out[3 * p] = sx[u] * inv; xyz_out[3 * p + 1] = sy[u] * inv; xyz_out[3 * p + 2] = sz[u] * inv;
        if (quat && quat_out) {
            float a0 = q0[u] * inv, a1 = q1[u] * inv, a2 = q2[u] * inv, a3 = q3[u] * inv;
            const float n = fmaxf(sqrtf(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3), 1e-12f);
            a0 /= n; a1 /= n; a2 /= n; a3 /= n;
            const float b0q = quat[4 * p], b1q = quat[4 * p + 1], b2q = quat[4 * p + 2], b3q = quat[4 * p + 3];
            quat_out[4 * p] = a0 * b0q - a1 * b1q - a2 * b2q - a3 * b3q;
            quat_out[4 * p + 1] = a0 * b1q + a1 * b0q + a2 * b3q - a3 * b2q;
            quat_out[4 * p + 2] = a0 * b2q - a1 * b3q + a2 * b0q + a3 * b1q;
            quat_out[4 * p + 3] = a0 * b3q + a1 * b2q - a2 * b1q + a3 * b0q;
        }
        if (weights_out && !weights_in) { // dense [n_particles, n_bones] like the reference returns; only on request
            for (int b = 0; b < n_bones; ++b) {
                const float dx = px[u] - tf[(size_t)b * SKIN_TF + 16], dy = py[u] - tf[(size_t)b * SKIN_TF + 17], dz = pz[u] - tf[(size_t)b * SKIN_TF + 18];
                const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                weights_out[(size_t)p * n_bones + b] = rsqrtf(fmaxf(d2, 1e-8f)) * inv;
            }
        }
    }
}

extern "C" int gsd_skin_bone_transforms(int32_t n_bones, const float *bones, const float *motions, const int32_t *row_ptr,
                                        const int32_t *cols, float *bone_tf, float *rot_out, void *stream) {
    if (n_bones < 0 || (n_bones > 0 && (!bones || !motions || !row_ptr || !cols || !bone_tf))) { gsd_set_error("invalid arguments"); return GSD_ERR_INVALID; }
    if (n_bones == 0) return GSD_OK;
    gsd_skin_bone_kernel<<<(n_bones + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_bones, bones, motions, row_ptr, cols, bone_tf, rot_out);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}

extern "C" int gsd_skin_apply(int32_t n_particles, int32_t n_bones, const float *xyz, const float *quat, const float *bone_tf,
                              const float *weights_in, float *xyz_out, float *quat_out, float *weights_out, void *stream) {
    if (n_particles < 0 || n_bones <= 0 || (n_particles > 0 && (!xyz || !bone_tf || !xyz_out)) || ((quat == nullptr) != (quat_out == nullptr))) {
        gsd_set_error("invalid arguments");
        return GSD_ERR_INVALID;
    }
    if (n_particles == 0) return GSD_OK;
    const int per_cta = 128 * SKIN_PPT;
    gsd_skin_apply_kernel<<<(n_particles + per_cta - 1) / per_cta, 128, 0, (cudaStream_t)stream>>>(n_particles, n_bones, xyz, quat, bone_tf, weights_in,
                                                                                                  xyz_out, quat_out, weights_out);
    GSD_LAUNCH_CHECK();
    return GSD_OK;
}
