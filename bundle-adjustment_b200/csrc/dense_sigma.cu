// dense_sigma.cu -- image points of ONE image with a fully populated dispersion matrix Sigma_ll (2m x 2m): north_star (2),
// BASELINE.json configs[3] "per-image dense Sigma_ll blocks".
//
// EXTENSION, PARITY UNPINNED beyond the block-diagonal case: the reference cannot express this -- an ImageCoordinate observation
// group is always the two rows of one point (camera/ImageCoordinate.java:102-104) with a 2 x 2 weight (PDF:296-319).  What is
// built here is the same least-squares contribution the reference's stacking loop (PDF:475-505) would produce for a group of
// 2m rows with the weight P = sigma0^2 Sigma^-1:  N += A'PA,  n += A'Pw,  Omega += v'Pv.  With a block-diagonal Sigma it is
// exactly the sum of the reference's per-point groups (tests pin that case against the faithful oracle).
//
// Per image, on the stream, one after the other (different images may see the same pair of points):
//   k_image_rows     compact Jacobian rows of the image, Ac[2m][128]: columns 0..2 the point, 3..8 the exterior orientation,
//                    9.. the camera's raw parameters (x0, y0, c, coefficients), then w; zeros for fixed parameters
//   k_gemm           T = P Ac            (FP64 tensor-core tiles, dense_kernels.cu)
//   k_gemm           G = Ac' T           (the exterior-orientation / camera block and its right-hand side)
//   k_image_scatter  G -> N, n;  point rows: N[p, EO | camera] += a_p' T_p,  n[p] += a_p' T_p[w]
//   k_image_pairs    N[p, q] += a_p' P[p, q] a_q  for every pair of points of the image
// P is computed once per adjustment (Sigma does not change): blocked Cholesky + inverse of Sigma / sigma0^2 on the device, the
// same schedule as the main system.  The standard sweeps skip these observations (their per-point weights are set to zero).
#include "common.h"
#include "dense_driver.hpp"
#include "model.cuh"
#include "sweep_common.cuh"

namespace jaicov {

constexpr int kRowsLd = 128;   // leading dimension of Ac / T (one 128-column tile)

// ---- compact rows of one image -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_image_rows(DevProblem P, int img, double *__restrict__ Ac, int64_t rp) {
    __shared__ CamSmem cs;
    __shared__ ImgPose s_pose;
    const int tid = threadIdx.x;
    const int cam = P.cam_of_img[img];
    load_camera(P, cam, cs, tid, blockDim.x);
    if (tid < 14) reinterpret_cast<double *>(&s_pose)[tid] = P.pose[(int64_t)img * kPoseStride + tid];
    __syncthreads();
    const CamView cv = view_of(P, cs);
    const int64_t o0 = P.pt_ptr[img], m = P.pt_ptr[img + 1] - o0;
    const int32_t *ec = P.eo_col + 6 * (int64_t)img, *cc = P.campos_col + P.cam_kbase[cam];
    const int wcol = 12 + cs.ncoef;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + tid; q < rp / 2; q += (int64_t)gridDim.x * blockDim.x) {
        double *r0 = Ac + (2 * q) * kRowsLd, *r1 = r0 + kRowsLd;
        for (int c = 0; c < kRowsLd; c++) { r0[c] = 0.0; r1[c] = 0.0; }
        if (q >= m) continue;                      // identity padding of P: zero rows
        const int64_t j = o0 + q;
        const int pt = P.obj_idx[j];
        const int32_t *pc = P.pt_col + 3 * (int64_t)pt;
        BaseRows r;
        eval_observation(s_pose, cv, P.xyz[3 * (int64_t)pt], P.xyz[3 * (int64_t)pt + 1], P.xyz[3 * (int64_t)pt + 2], P.xy[2 * j],
                         P.xy[2 * j + 1], r, [&](int k, double v0, double v1) {
                             if (col_active(cc[3 + k])) { r0[12 + k] = v0; r1[12 + k] = v1; }
                         });
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (col_active(pc[i])) { r0[i] = r.ax[i]; r1[i] = r.ay[i]; }
            if (col_active(ec[i])) { r0[3 + i] = -r.ax[i]; r1[3 + i] = -r.ay[i]; }
            if (col_active(ec[3 + i])) { r0[6 + i] = r.ax[4 + i]; r1[6 + i] = r.ay[4 + i]; }
        }
        if (col_active(cc[0])) r0[9] = 1.0;
        if (col_active(cc[1])) r1[10] = 1.0;
        if (col_active(cc[2])) { r0[11] = r.ax[3]; r1[11] = r.ay[3]; }
        r0[wcol] = r.w0;
        r1[wcol] = r.w1;
    }
}

// column of the system matrix of compact column c (3.. : exterior orientation, camera) -- -1 if fixed / not a parameter column
__device__ __forceinline__ int32_t shared_col(const DevProblem &P, int img, int cam, int c, int ncoef) {
    if (c >= 3 && c < 9) return P.eo_col[6 * (int64_t)img + c - 3];
    if (c >= 9 && c < 12 + ncoef) return P.campos_col[P.cam_kbase[cam] + c - 9];
    return -1;
}

// G = Ac'T (128 x 128): exterior orientation / camera block and its right-hand side; then the rows of every point
__global__ void __launch_bounds__(256) k_image_scatter(DevProblem P, int img, const double *__restrict__ Ac, const double *__restrict__ T,
                                                       const double *__restrict__ G, double *__restrict__ M, double *__restrict__ rhs) {
    const int cam = P.cam_of_img[img];
    const int ncoef = P.coef_ptr[cam + 1] - P.coef_ptr[cam];
    const int wcol = 12 + ncoef, d = P.d;
    const int64_t ld = P.np;
    const int64_t o0 = P.pt_ptr[img], m = P.pt_ptr[img + 1] - o0;
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < (wcol - 3) * (wcol - 2); i += blockDim.x) {
            const int a = 3 + i / (wcol - 2), b = 3 + i % (wcol - 2);
            const int32_t ca = shared_col(P, img, cam, a, ncoef);
            if (!col_active(ca)) continue;
            if (b == wcol) { rhs[ca - d] += G[a * kRowsLd + wcol]; continue; }
            if (b > a) continue;
            const int32_t cb = shared_col(P, img, cam, b, ncoef);
            if (!col_active(cb)) continue;
            M[lower_idx(ca - d, cb - d, ld)] += G[a * kRowsLd + b];
        }
        return;
    }
    // one thread per (point of the image, shared column or w)
    const int nsh = wcol - 2;      // columns 3 .. wcol
    for (int64_t i = (int64_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x; i < m * nsh; i += (int64_t)(gridDim.x - 1) * blockDim.x) {
        const int64_t q = i / nsh;
        const int b = 3 + (int)(i % nsh);
        const int pt = P.obj_idx[o0 + q];
        const int32_t *pc = P.pt_col + 3 * (int64_t)pt;
        const double *a0 = Ac + (2 * q) * kRowsLd, *a1 = a0 + kRowsLd, *t0 = T + (2 * q) * kRowsLd, *t1 = t0 + kRowsLd;
        const int32_t cb = b == wcol ? 0 : shared_col(P, img, cam, b, ncoef);
        if (b != wcol && !col_active(cb)) continue;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (!col_active(pc[c])) continue;
            const double v = a0[c] * t0[b] + a1[c] * t1[b];
            if (b == wcol) rhs[pc[c] - d] += v;       // one thread per (point, column): no other writer of this entry in this launch
            else M[lower_idx(pc[c] - d, cb - d, ld)] += v;
        }
    }
}

// N[p, q] += a_p' P[rows p, rows q] a_q for all pairs of points of the image (q' <= q in observation order; every entry of the
// lower triangle has exactly one writer in this launch)
__global__ void __launch_bounds__(256) k_image_pairs(DevProblem P, int img, const double *__restrict__ Ac, const double *__restrict__ Pw,
                                                     int64_t ldp, double *__restrict__ M) {
    const int64_t o0 = P.pt_ptr[img], m = P.pt_ptr[img + 1] - o0;
    const int d = P.d;
    const int64_t ld = P.np;
    const int64_t q = (int64_t)blockIdx.y * blockDim.y + threadIdx.y, q2 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m || q2 > q) return;
    const int32_t *pa = P.pt_col + 3 * (int64_t)P.obj_idx[o0 + q], *pb = P.pt_col + 3 * (int64_t)P.obj_idx[o0 + q2];
    const double *a0 = Ac + (2 * q) * kRowsLd, *a1 = a0 + kRowsLd, *b0 = Ac + (2 * q2) * kRowsLd, *b1 = b0 + kRowsLd;
    const double p00 = Pw[(2 * q) * ldp + 2 * q2], p01 = Pw[(2 * q) * ldp + 2 * q2 + 1], p10 = Pw[(2 * q + 1) * ldp + 2 * q2],
                 p11 = Pw[(2 * q + 1) * ldp + 2 * q2 + 1];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (!col_active(pa[i])) continue;
        const double u0 = a0[i] * p00 + a1[i] * p10, u1 = a0[i] * p01 + a1[i] * p11;      // (a_p' P)[i][.]
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (q2 == q && j > i) continue;
            if (!col_active(pb[j])) continue;
            M[lower_idx(pa[i] - d, pb[j] - d, ld)] += u0 * b0[j] + u1 * b1[j];
        }
    }
}

// ---- Omega and the matrix-free product: v = w - Ac x (or Ac x), t = P v ----------------------------------------------------------
// mode 0: v = w - A x (Omega);  mode 1: v = A x (product);  mode 2: v = w (right-hand side of the product)
__global__ void __launch_bounds__(128) k_image_residual(DevProblem P, int img, const double *__restrict__ Ac, const double *__restrict__ x,
                                                        int mode, double *__restrict__ v, int64_t rp) {
    const int cam = P.cam_of_img[img];
    const int ncoef = P.coef_ptr[cam + 1] - P.coef_ptr[cam];
    const int wcol = 12 + ncoef;
    const int64_t o0 = P.pt_ptr[img], m = P.pt_ptr[img + 1] - o0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rp; r += (int64_t)gridDim.x * blockDim.x) {
        if (r >= 2 * m) { v[r] = 0.0; continue; }
        const double *a = Ac + r * kRowsLd;
        double s = 0.0;
        if (mode != 2) {
            const int32_t *pc = P.pt_col + 3 * (int64_t)P.obj_idx[o0 + r / 2];
            for (int c = 0; c < 3; c++)
                if (col_active(pc[c])) s += a[c] * x[pc[c]];
            for (int c = 3; c < wcol; c++) {
                const int32_t col = shared_col(P, img, cam, c, ncoef);
                if (col_active(col)) s += a[c] * x[col];
            }
        }
        v[r] = mode == 0 ? a[wcol] - s : (mode == 1 ? s : a[wcol]);
    }
}

// t = Pw v (one warp per row), out[0] += v't (if out)
__global__ void __launch_bounds__(256) k_image_weight_times(const double *__restrict__ Pw, int64_t ldp, int64_t rp, const double *__restrict__ v,
                                                            double *__restrict__ t, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rp) return;
    double s = 0.0;
    for (int64_t k = lane; k < rp; k += 32) s += Pw[row * ldp + k] * v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        t[row] = s;
        if (out) atomicAdd(out, v[row] * s);
    }
}

// y += Ac' t scattered to the system columns (matrix-free product / its right-hand side)
__global__ void __launch_bounds__(128) k_image_scatter_vector(DevProblem P, int img, const double *__restrict__ Ac, const double *__restrict__ t,
                                                              double *__restrict__ y) {
    const int cam = P.cam_of_img[img];
    const int ncoef = P.coef_ptr[cam + 1] - P.coef_ptr[cam];
    const int wcol = 12 + ncoef;
    const int64_t o0 = P.pt_ptr[img], m = P.pt_ptr[img + 1] - o0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < 2 * m; r += (int64_t)gridDim.x * blockDim.x) {
        const double *a = Ac + r * kRowsLd;
        const double tr = t[r];
        const int32_t *pc = P.pt_col + 3 * (int64_t)P.obj_idx[o0 + r / 2];
        for (int c = 0; c < 3; c++)
            if (col_active(pc[c])) atomicAdd(y + pc[c], a[c] * tr);
        for (int c = 3; c < wcol; c++) {
            const int32_t col = shared_col(P, img, cam, c, ncoef);
            if (col_active(col)) atomicAdd(y + col, a[c] * tr);
        }
    }
}

// zero the per-point weights of the image's observations: the standard sweeps then contribute nothing for them
__global__ void k_zero_weights(double *__restrict__ rw, int64_t begin, int64_t end) {
    const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) rw[i] = 0.0;
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
void launch_gemm(const GemmDesc &g, cudaStream_t s);

void launch_zero_weights(double *rw, int64_t obs_begin, int64_t obs_end, cudaStream_t s) {
    const int64_t n = 3 * (obs_end - obs_begin);
    if (n <= 0) return;
    g_launch_count++;
    k_zero_weights<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rw, 3 * obs_begin, 3 * obs_end);
}

void launch_image_rows(const DevProblem &P, int img, double *Ac, int64_t rp, cudaStream_t s) {
    g_launch_count++;
    k_image_rows<<<(unsigned)std::min<int64_t>(148 * 4, (rp / 2 + 127) / 128), 128, 0, s>>>(P, img, Ac, rp);
}

// N += A'PA, n += A'Pw of one image with the dense weight Pw (rp x rp, ldp); Ac, T: rp x 128 scratch, G: 128 x 128 scratch
void launch_dense_image_assemble(const DevProblem &P, int img, int64_t m, const double *Pw, int64_t ldp, int64_t rp, double *Ac, double *T,
                                 double *G, double *M, double *rhs, cudaStream_t s) {
    launch_image_rows(P, img, Ac, rp, s);
    GemmDesc g;      // T = Pw Ac
    g.al = 0; g.bl = 1; g.mt = (int)(rp / kBlk); g.nt = 1; g.K = rp; g.alpha = 1.0; g.beta = 0.0;
    g.A = Pw; g.lda = ldp; g.B = Ac; g.ldb = kRowsLd; g.C = T; g.ldc = kRowsLd;
    launch_gemm(g, s);
    GemmDesc q;      // G = Ac' T
    q.al = 1; q.bl = 1; q.mt = 1; q.nt = 1; q.K = rp; q.alpha = 1.0; q.beta = 0.0;
    q.A = Ac; q.lda = kRowsLd; q.B = T; q.ldb = kRowsLd; q.C = G; q.ldc = kRowsLd;
    launch_gemm(q, s);
    g_launch_count++;
    k_image_scatter<<<1 + (unsigned)std::min<int64_t>(148 * 4, (m * 80 + 255) / 256), 256, 0, s>>>(P, img, Ac, T, G, M, rhs);
    g_launch_count++;
    const dim3 blk(16, 16), grd((unsigned)((m + 15) / 16), (unsigned)((m + 15) / 16));
    k_image_pairs<<<grd, blk, 0, s>>>(P, img, Ac, Pw, ldp, M);
}

// out[0] += v'Pv with v = w - A x  (Omega part of the image); v, t: rp scratch
void launch_dense_image_omega(const DevProblem &P, int img, const double *Pw, int64_t ldp, int64_t rp, double *Ac, const double *x, double *v,
                              double *t, double *out, cudaStream_t s) {
    launch_image_rows(P, img, Ac, rp, s);
    g_launch_count += 2;
    k_image_residual<<<(unsigned)std::min<int64_t>(148 * 4, (rp + 127) / 128), 128, 0, s>>>(P, img, Ac, x, 0, v, rp);
    k_image_weight_times<<<(unsigned)((rp + 7) / 8), 256, 0, s>>>(Pw, ldp, rp, v, t, out);
}

// matrix-free product of the image's block: y += A'P(A x) (mode 1), or rhs += A'Pw and wpw += w'Pw (mode 2)
void launch_dense_image_product(const DevProblem &P, int img, const double *Pw, int64_t ldp, int64_t rp, const double *Ac, const double *x,
                                int mode, double *v, double *t, double *y, double *wpw, cudaStream_t s) {
    g_launch_count += 3;
    k_image_residual<<<(unsigned)std::min<int64_t>(148 * 4, (rp + 127) / 128), 128, 0, s>>>(P, img, Ac, x, mode, v, rp);
    k_image_weight_times<<<(unsigned)((rp + 7) / 8), 256, 0, s>>>(Pw, ldp, rp, v, t, mode == 2 ? wpw : nullptr);
    k_image_scatter_vector<<<(unsigned)std::min<int64_t>(148 * 4, (rp + 127) / 128), 128, 0, s>>>(P, img, Ac, t, y);
}

}  // namespace jaicov
