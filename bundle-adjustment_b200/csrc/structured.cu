// structured.cu -- kernels of the point-block ("structured") solver: SURVEY.md section 8(f-2), second half.
//
// When no observation couples two different object points (no scale bars, no directly observed point groups), the
// leading block of N that belongs to the object coordinates is block diagonal (one block of <= 3 columns per point):
//
//        [ P   C ]   p: object coordinates (u_p columns, block diagonal P)
//    N = [ C'  R ]   r: interior orientation, distortion, exterior orientations (nc columns)
//
// The reference does not exploit this (dspsv + dsptri on the packed n x n matrix, MathExtension.java:338-366); the
// RESULT it produces -- dx and the complete inverse of the bordered system K = [[0, B], [B', N]] (BA:246-273) -- is
// reproduced here from the block structure, on the same Jacobi-scaled system the dense route uses:
//
//    Z  = [ C | B_p' ]            (u_p x m, m = nc + d; stored transposed: Zt, m rows)
//    Y  = P^-1 Z                  (block-wise, Yt)
//    K' = [[R, B_r'],[B_r, 0]] - Z' Y          reduced (camera) system with the datum border, m x m
//    K'^-1 =: Q'                  S~ = S + E D^-1 E' is SPD (S, E, -D the blocks of K'): Cholesky + inverse of S~ with the
//                                 tensor-core schedule of dense_driver.hpp, then the d-row border algebra
//    K^-1 [p, p] = P^-1 + Y Q' Y',   K^-1 [p, r|lambda] = -Y Q',   K^-1 [r|lambda, r|lambda] = Q'
//
// The O(u_p^2 m) product Y Q' Y' and the two O(u_p m^2) products are launches of k_gemm (dense_kernels.cu); the kernels
// below are the HBM-bound glue.  Everything is single-owner and fixed-order: bitwise reproducible.
#include "common.h"

namespace jaicov {

// ---- point blocks: P_b^-1 by Cholesky of the <= 3 x 3 block --------------------------------------------------------------
__global__ void __launch_bounds__(128) k_point_block_inv(const double *__restrict__ M, int64_t ld, const int32_t *__restrict__ blk_start,
                                                         const int32_t *__restrict__ blk_size, int nBlk,
                                                         const double *__restrict__ V, double *__restrict__ Pinv,
                                                         int *__restrict__ info) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlk) return;
    const int64_t c0 = blk_start[b];
    const int sz = blk_size[b];
    double a[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < sz; i++)    // the blocks of M hold N unscaled: apply the Jacobi scaling V N V here
        for (int j = 0; j <= i; j++) a[i][j] = (V[c0 + i] * M[(c0 + i) * ld + c0 + j]) * V[c0 + j];
    // L L' = A
    double l[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    bool bad = false;
    for (int j = 0; j < 3; j++) {
        double s = a[j][j];
        for (int k = 0; k < j; k++) s -= l[j][k] * l[j][k];
        if (!(s > 0.0)) { bad = true; s = 1.0; }
        const double ljj = sqrt(s);
        l[j][j] = ljj;
        for (int i = j + 1; i < 3; i++) {
            double t = a[i][j];
            for (int k = 0; k < j; k++) t -= l[i][k] * l[j][k];
            l[i][j] = t / ljj;
        }
    }
    if (bad) atomicCAS(info, 0, (int)c0 + 1);
    // W = L^-1 (lower), A^-1 = W' W
    double w[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < 3; j++) {
        w[j][j] = 1.0 / l[j][j];
        for (int i = j + 1; i < 3; i++) {
            double t = 0.0;
            for (int k = j; k < i; k++) t -= l[i][k] * w[k][j];
            w[i][j] = t / l[i][i];
        }
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double t = 0.0;
            for (int k = (i > j ? i : j); k < 3; k++) t += w[k][i] * w[k][j];
            Pinv[(size_t)b * 9 + i * 3 + j] = (i < sz && j < sz) ? t : 0.0;
        }
}

void launch_point_block_inv(const double *M, int64_t ld, const int32_t *blk_start, const int32_t *blk_size, int nBlk, const double *V,
                            double *Pinv, int *info, cudaStream_t s) {
    if (nBlk == 0) return;
    g_launch_count++;
    k_point_block_inv<<<(unsigned)((nBlk + 127) / 128), 128, 0, s>>>(M, ld, blk_start, blk_size, nBlk, V, Pinv, info);
}

// the structured route clears only what the assembly accumulates into: the point blocks (here) and the rows of the
// camera / image unknowns (one memset); the rest of the object-point region of M is never read
__global__ void __launch_bounds__(128) k_zero_point_blocks(double *__restrict__ M, int64_t ld, const int32_t *__restrict__ blk_start,
                                                           const int32_t *__restrict__ blk_size, int nBlk) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlk) return;
    const int64_t c0 = blk_start[b];
    const int sz = blk_size[b];
    for (int i = 0; i < sz; i++)
        for (int j = 0; j <= i; j++) M[(c0 + i) * ld + c0 + j] = 0.0;
}

void launch_zero_point_blocks(double *M, int64_t ld, const int32_t *blk_start, const int32_t *blk_size, int nBlk, cudaStream_t s) {
    if (nBlk == 0) return;
    g_launch_count++;
    k_zero_point_blocks<<<(unsigned)((nBlk + 127) / 128), 128, 0, s>>>(M, ld, blk_start, blk_size, nBlk);
}

// Yt[:, p] *= V[p]: with the scaled operand both products of the inverse come out as V (.) V directly
__global__ void __launch_bounds__(256) k_scale_yt(double *__restrict__ Yt, StructDims D, const double *__restrict__ V) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D.up) return;
    Yt[(int64_t)blockIdx.y * D.Tp + p] *= V[p];
}

void launch_scale_yt(double *Yt, const StructDims &D, const double *V, cudaStream_t s) {
    g_launch_count++;
    k_scale_yt<<<dim3((unsigned)((D.up + 255) / 256), (unsigned)(D.nc + D.d)), 256, 0, s>>>(Yt, D, V);
}

// ---- Zt (m x Tp) = [C' ; B_p] and Yt = Zt P^-1 (block-wise); rows >= m and columns >= up are zero -----------------------------

__device__ __forceinline__ double zt_entry(const double *__restrict__ M, const double *__restrict__ Btv, const StructDims &D,
                                           int64_t j, int64_t p) {
    if (j < D.nc) return M[(D.up + j) * D.np + p];
    return Btv[(j - D.nc) * D.np + p];
}

__global__ void __launch_bounds__(256) k_build_zy(const double *__restrict__ M, const double *__restrict__ Btv, StructDims D,
                                                  const int32_t *__restrict__ col_blk, const int32_t *__restrict__ blk_start,
                                                  const int32_t *__restrict__ blk_size, const double *__restrict__ Pinv,
                                                  double *__restrict__ Zt, double *__restrict__ Yt) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t j = blockIdx.y;
    if (p >= D.Tp) return;
    double z = 0.0, y = 0.0;
    if (j < D.nc + D.d && p < D.up) {
        const int b = col_blk[p];
        const int64_t c0 = blk_start[b];
        const int sz = blk_size[b], q = (int)(p - c0);
        const double *pi = Pinv + (size_t)b * 9;
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < sz) {
                const double zk = zt_entry(M, Btv, D, j, c0 + k);
                if (k == q) z = zk;
                y += zk * pi[k * 3 + q];
            }
    }
    Zt[j * D.Tp + p] = z;
    Yt[j * D.Tp + p] = y;
}

void launch_build_zy(const double *M, const double *Btv, const StructDims &D, const int32_t *col_blk, const int32_t *blk_start,
                     const int32_t *blk_size, const double *Pinv, double *Zt, double *Yt, cudaStream_t s) {
    g_launch_count++;
    k_build_zy<<<dim3((unsigned)((D.Tp + 255) / 256), (unsigned)D.mp), 256, 0, s>>>(M, Btv, D, col_blk, blk_start, blk_size, Pinv, Zt, Yt);
}

// ---- Kp (lower) = [[R, .],[B_r, 0]]; the product Z'Y is subtracted by a GEMM launch ---------------------------------------
__global__ void __launch_bounds__(256) k_init_kp(const double *__restrict__ M, const double *__restrict__ Btv, StructDims D,
                                                 double *__restrict__ Kp) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= D.mp) return;
    double v = 0.0;
    if (j <= i) {
        if (i < D.nc) v = M[(D.up + i) * D.np + D.up + j];
        else if (i < D.nc + D.d && j < D.nc) v = Btv[(i - D.nc) * D.np + D.up + j];
    }
    Kp[i * D.mp + j] = v;
}

void launch_init_kp(const double *M, const double *Btv, const StructDims &D, double *Kp, cudaStream_t s) {
    g_launch_count++;
    k_init_kp<<<dim3((unsigned)((D.mp + 255) / 256), (unsigned)D.mp), 256, 0, s>>>(M, Btv, D, Kp);
}

// ---- border of the reduced system: D = -K'[l,l], D^-1, E = K'[l, r] (rows), ED = D^-1 E -------------------------------------
// sb[]: [0,49) D^-1, [49,98) Q'[l,l] (written later), [98] singular flag
__device__ void gauss_jordan_inverse(int d, double A[kMaxDatum][2 * kMaxDatum], bool &sing) {
    sing = false;
    for (int c = 0; c < d; c++) {
        int pr = c;
        for (int i = c + 1; i < d; i++)
            if (fabs(A[i][c]) > fabs(A[pr][c])) pr = i;
        if (!(fabs(A[pr][c]) > 0.0)) { sing = true; return; }
        if (pr != c)
            for (int j = 0; j < 2 * d; j++) { const double t = A[c][j]; A[c][j] = A[pr][j]; A[pr][j] = t; }
        const double pv = 1.0 / A[c][c];
        for (int j = 0; j < 2 * d; j++) A[c][j] *= pv;
        for (int i = 0; i < d; i++) {
            if (i == c) continue;
            const double f = A[i][c];
            for (int j = 0; j < 2 * d; j++) A[i][j] -= f * A[c][j];
        }
    }
}

__global__ void __launch_bounds__(256) k_border_prep(const double *__restrict__ Kp, StructDims D, double *__restrict__ Eb,
                                                     double *__restrict__ ED, double *__restrict__ sb) {
    __shared__ double s_dinv[kMaxDatum][kMaxDatum];
    const int d = D.d, nc = D.nc;
    if (threadIdx.x == 0) {
        double A[kMaxDatum][2 * kMaxDatum];
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++) {
                const int hi = i > j ? i : j, lo = i > j ? j : i;
                A[i][j] = -Kp[(int64_t)(nc + hi) * D.mp + nc + lo];
                A[i][d + j] = (i == j) ? 1.0 : 0.0;
            }
        bool sing;
        gauss_jordan_inverse(d, A, sing);
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++) {
                const double v = sing ? 0.0 : 0.5 * (A[i][d + j] + A[j][d + i]);
                s_dinv[i][j] = v;
                sb[i * kMaxDatum + j] = v;
            }
        sb[98] = sing ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int64_t j = threadIdx.x; j < nc; j += blockDim.x) {
        double e[kMaxDatum];
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++) e[a] = a < d ? Kp[(int64_t)(nc + a) * D.mp + j] : 0.0;
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < d) {
                double t = 0.0;
#pragma unroll
                for (int b = 0; b < kMaxDatum; b++)
                    if (b < d) t += s_dinv[a][b] * e[b];
                Eb[a * D.ncp + j] = e[a];
                ED[a * D.ncp + j] = t;
            }
    }
}

void launch_border_prep(const double *Kp, const StructDims &D, double *Eb, double *ED, double *sb, cudaStream_t s) {
    g_launch_count++;
    k_border_prep<<<1, 256, 0, s>>>(Kp, D, Eb, ED, sb);
}

// ---- S~ = S + E D^-1 E' (lower, identity padding) -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_form_stilde(const double *__restrict__ Kp, StructDims D, const double *__restrict__ Eb,
                                                     const double *__restrict__ ED, double *__restrict__ Sm) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= D.ncp) return;
    double v = 0.0;
    if (j <= i) {
        if (i < D.nc) {
            v = Kp[i * D.mp + j];
            for (int a = 0; a < D.d; a++) v += ED[a * D.ncp + i] * Eb[a * D.ncp + j];
        } else if (i == j) {
            v = 1.0;
        }
    }
    Sm[i * D.ncp + j] = v;
}

void launch_form_stilde(const double *Kp, const StructDims &D, const double *Eb, const double *ED, double *Sm, cudaStream_t s) {
    g_launch_count++;
    k_form_stilde<<<dim3((unsigned)((D.ncp + 255) / 256), (unsigned)D.ncp), 256, 0, s>>>(Kp, D, Eb, ED, Sm);
}

// ---- F[a][i] = sum_j S~^-1[i][j] ED[a][j]   (S~^-1 symmetric, full storage); one warp per row i ---------------------------------
__global__ void __launch_bounds__(256) k_border_f(const double *__restrict__ Sm, StructDims D, const double *__restrict__ ED,
                                                  double *__restrict__ Fb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + warp;
    if (i >= D.nc) return;
    double acc[kMaxDatum];
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) acc[a] = 0.0;
    for (int64_t j = lane; j < D.nc; j += 32) {
        const double sv = Sm[i * D.ncp + j];
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < D.d) acc[a] += sv * ED[a * D.ncp + j];
    }
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) {
        double v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && a < D.d) Fb[a * D.ncp + i] = v;
    }
}

// Q'[l,l] = -D^-1 + ED F'   (d x d), one CTA
__global__ void __launch_bounds__(256) k_border_ll(StructDims D, const double *__restrict__ ED, const double *__restrict__ Fb,
                                                   double *__restrict__ sb) {
    __shared__ double red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int a = 0; a < D.d; a++)
        for (int b = 0; b <= a; b++) {
            double v = 0.0;
            for (int64_t j = threadIdx.x; j < D.nc; j += blockDim.x) v += ED[a * D.ncp + j] * Fb[b * D.ncp + j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int w = 0; w < 8; w++) t += red[w];
                t -= sb[a * kMaxDatum + b];
                sb[49 + a * kMaxDatum + b] = t;
                sb[49 + b * kMaxDatum + a] = t;
            }
            __syncthreads();
        }
}

void launch_border_f(const double *Sm, const StructDims &D, const double *ED, double *Fb, double *sb, cudaStream_t s) {
    if (D.d == 0) return;
    g_launch_count += 2;
    k_border_f<<<(unsigned)((D.nc + 7) / 8), 256, 0, s>>>(Sm, D, ED, Fb);
    k_border_ll<<<1, 256, 0, s>>>(D, ED, Fb, sb);
}

// ---- Q' (mp x mp, full symmetric storage, zero padding) ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_qprime(const double *__restrict__ Sm, StructDims D, const double *__restrict__ Fb,
                                                     const double *__restrict__ sb, double *__restrict__ Kp) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= D.mp) return;
    const int64_t nc = D.nc, m = D.nc + D.d;
    double v = 0.0;
    if (i < nc && j < nc) v = Sm[i * D.ncp + j];
    else if (i < nc && j < m) v = Fb[(j - nc) * D.ncp + i];
    else if (i < m && j < nc) v = Fb[(i - nc) * D.ncp + j];
    else if (i < m && j < m) v = sb[49 + (i - nc) * kMaxDatum + (j - nc)];
    Kp[i * D.mp + j] = v;
}

void launch_fill_qprime(const double *Sm, const StructDims &D, const double *Fb, const double *sb, double *Kp, cudaStream_t s) {
    g_launch_count++;
    k_fill_qprime<<<dim3((unsigned)((D.mp + 255) / 256), (unsigned)D.mp), 256, 0, s>>>(Sm, D, Fb, sb, Kp);
}

// ---- solution ------------------------------------------------------------------------------------------------------------------------
// zp = P^-1 n_p  (n = row 0 of the right-hand-side block, already Jacobi-scaled)
__global__ void __launch_bounds__(256) k_point_rhs(const double *__restrict__ nrm, StructDims D, const int32_t *__restrict__ col_blk,
                                                   const int32_t *__restrict__ blk_start, const int32_t *__restrict__ blk_size,
                                                   const double *__restrict__ Pinv, double *__restrict__ zp) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D.Tp) return;
    double z = 0.0;
    if (p < D.up) {
        const int b = col_blk[p];
        const int64_t c0 = blk_start[b];
        const int sz = blk_size[b], q = (int)(p - c0);
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < sz) z += Pinv[(size_t)b * 9 + k * 3 + q] * nrm[c0 + k];
    }
    zp[p] = z;
}

// rp[j] = (j < nc ? n_r[j] : 0) - Zt[j,:] zp ; one CTA per row, fixed-order reduction
__global__ void __launch_bounds__(256) k_reduced_rhs(const double *__restrict__ Zt, StructDims D, const double *__restrict__ zp,
                                                     const double *__restrict__ nrm, double *__restrict__ rp) {
    __shared__ double red[8];
    const int64_t j = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = 0.0;
    if (j < D.nc + D.d) {
        const double *row = Zt + j * D.Tp;
        for (int64_t p = threadIdx.x; p < D.up; p += blockDim.x) v += row[p] * zp[p];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += red[w];
        double r = 0.0;
        if (j < D.nc) r = nrm[D.up + j] - t;
        else if (j < D.nc + D.d) r = -t;
        rp[j] = r;
    }
}

// yr = Q' rp ; one warp per row
__global__ void __launch_bounds__(256) k_reduced_solve(const double *__restrict__ Kp, StructDims D, const double *__restrict__ rp,
                                                       double *__restrict__ yr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + warp;
    if (i >= D.mp) return;
    const int64_t m = D.nc + D.d;
    double v = 0.0;
    if (i < m)
        for (int64_t j = lane; j < m; j += 32) v += Kp[i * D.mp + j] * rp[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) yr[i] = v;
}

// y_p = zp - Yt' yr (Jacobi-scaled solution, not yet multiplied by V), lambda, and the d rows
// Nt = K^-1[lambda, x] (= -Q'[l,:] Yt on the point columns, Q'[l, r] on the others) in the same sweep over Yt.
// Nt spans the null space of N with B Nt' = I: k_datum_project uses it to enforce the datum conditions B y = 0 to
// rounding level.  Without that step the conditions hold only to the accuracy of Q' (~1e-9 relative), and what is
// left is a drift along the unobservable datum directions that no later iteration removes.
__global__ void __launch_bounds__(256) k_back_substitute(const double *__restrict__ Yt, StructDims D, const double *__restrict__ Kp,
                                                         const double *__restrict__ yr, const double *__restrict__ zp,
                                                         double *__restrict__ ys, double *__restrict__ Nt,
                                                         double *__restrict__ dxref) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < D.d) dxref[e] = yr[D.nc + e];
    if (e >= D.u) return;
    if (e < D.up) {
        const int64_t m = D.nc + D.d;
        double acc = 0.0, g[kMaxDatum];
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++) g[a] = 0.0;
        const double *ql = Kp + (int64_t)D.nc * D.mp;
#pragma unroll 2
        for (int64_t j = 0; j < m; j++) {
            const double yv = Yt[j * D.Tp + e];
            acc += yv * yr[j];
#pragma unroll
            for (int a = 0; a < kMaxDatum; a++)
                if (a < D.d) g[a] += ql[a * D.mp + j] * yv;
        }
        ys[e] = zp[e] - acc;
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < D.d) Nt[a * D.np + e] = -g[a];
    } else {
        ys[e] = yr[e - D.up];
        for (int a = 0; a < D.d; a++) Nt[a * D.np + e] = Kp[(int64_t)(D.nc + a) * D.mp + (e - D.up)];
    }
}

// t = B~ ys (one CTA, fixed order)
__global__ void __launch_bounds__(256) k_datum_residual(StructDims D, const double *__restrict__ Btv, const double *__restrict__ ys,
                                                        double *__restrict__ t) {
    __shared__ double red[8][kMaxDatum];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[kMaxDatum];
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) acc[a] = 0.0;
    for (int64_t e = threadIdx.x; e < D.u; e += blockDim.x) {
        const double y = ys[e];
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < D.d) acc[a] += Btv[a * D.np + e] * y;
    }
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) {
        double v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    if (threadIdx.x < kMaxDatum) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += red[w][threadIdx.x];
        t[threadIdx.x] = threadIdx.x < D.d ? v : 0.0;
    }
}

// dx = V (ys - Nt' t); border rows of the inverse: Tq[a][e] = V[e] K^-1[lambda_a, x_e]
__global__ void __launch_bounds__(256) k_datum_project(StructDims D, const double *__restrict__ ys, const double *__restrict__ Nt,
                                                       const double *__restrict__ t, const double *__restrict__ V,
                                                       double *__restrict__ dxref, double *__restrict__ Tq) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.u) return;
    double y = ys[e];
    const double v = V[e];
    for (int a = 0; a < D.d; a++) {
        const double na = Nt[a * D.np + e];
        y -= na * t[a];
        Tq[a * D.np + e] = v * na;
    }
    dxref[D.d + e] = v * y;
}

void launch_structured_solution(const double *nrm, const StructDims &D, const int32_t *col_blk, const int32_t *blk_start,
                                const int32_t *blk_size, const double *Pinv, const double *Zt, const double *Yt, const double *Kp,
                                const double *Btv, const double *V, double *zp, double *rp, double *yr, double *ys, double *Nt,
                                double *t, double *dxref, double *Tq, cudaStream_t s) {
    g_launch_count += 6;
    k_point_rhs<<<(unsigned)((D.Tp + 255) / 256), 256, 0, s>>>(nrm, D, col_blk, blk_start, blk_size, Pinv, zp);
    k_reduced_rhs<<<(unsigned)D.mp, 256, 0, s>>>(Zt, D, zp, nrm, rp);
    k_reduced_solve<<<(unsigned)((D.mp + 7) / 8), 256, 0, s>>>(Kp, D, rp, yr);
    const int64_t n = D.u > D.d ? D.u : D.d;
    k_back_substitute<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(Yt, D, Kp, yr, zp, ys, Nt, dxref);
    k_datum_residual<<<1, 256, 0, s>>>(D, Btv, ys, t);
    k_datum_project<<<(unsigned)((D.u + 255) / 256), 256, 0, s>>>(D, ys, Nt, t, V, dxref, Tq);
}

// ---- placement of the inverse into M (lower, row-major): rows of the r group, identity padding, P^-1 on the point blocks -------
// (T1t was formed from the V-scaled Yt: its columns carry V already; the row factor and the r-r block are scaled here)
__global__ void __launch_bounds__(256) k_place_rows(double *__restrict__ M, StructDims D, const double *__restrict__ T1t,
                                                    const double *__restrict__ Kp, const double *__restrict__ V, int64_t row_end) {
    const int64_t r = D.up + blockIdx.x;        // row of M
    if (r >= row_end) return;
    double *row = M + r * D.np;
    const int64_t i = blockIdx.x;
    if (i < D.nc) {
        const double *t = T1t + i * D.Tp;
        const double vr = V[r];
        for (int64_t c = threadIdx.x; c <= r; c += blockDim.x)
            row[c] = c < D.up ? -t[c] * vr : (V[c] * Kp[i * D.mp + (c - D.up)]) * vr;
    } else {
        for (int64_t c = threadIdx.x; c <= r; c += blockDim.x) row[c] = (c == r) ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(128) k_add_point_blocks(double *__restrict__ M, int64_t ld, const int32_t *__restrict__ blk_start,
                                                          const int32_t *__restrict__ blk_size, int nBlk,
                                                          const double *__restrict__ Pinv, const double *__restrict__ V) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlk) return;
    const int64_t c0 = blk_start[b];
    const int sz = blk_size[b];
    for (int i = 0; i < sz; i++)
        for (int j = 0; j <= i; j++) M[(c0 + i) * ld + c0 + j] += (V[c0 + j] * Pinv[(size_t)b * 9 + i * 3 + j]) * V[c0 + i];
}

void launch_structured_place(double *M, const StructDims &D, const double *T1t, const double *Kp, const int32_t *blk_start,
                             const int32_t *blk_size, int nBlk, const double *Pinv, const double *V, cudaStream_t s) {
    int64_t row_end = D.Tp < D.np ? D.Tp : D.np;
    if (row_end < D.u) row_end = D.u;
    g_launch_count += 2;
    k_place_rows<<<(unsigned)(row_end - D.up), 256, 0, s>>>(M, D, T1t, Kp, V, row_end);
    k_add_point_blocks<<<(unsigned)((nBlk + 127) / 128), 128, 0, s>>>(M, D.np, blk_start, blk_size, nBlk, Pinv, V);
}

// ---- the same placement on a rank's column tiles (multi-GPU): X is np x (128 ntc), local tile t = global columns ktab[t].. ------
// T1l holds Q'Y' for this rank's object-coordinate tiles only (the first ntp local tiles), leading dimension ldt
__global__ void __launch_bounds__(256) k_place_cols(double *__restrict__ X, int64_t ldx, int ntc, const int32_t *__restrict__ ktab,
                                                    StructDims D, const double *__restrict__ T1l, int64_t ldt,
                                                    const double *__restrict__ Kp, const double *__restrict__ V, int64_t row_end) {
    const int64_t r = D.up + blockIdx.x;
    if (r >= row_end) return;
    const int64_t i = blockIdx.x;
    double *row = X + r * ldx;
    const double vr = i < D.nc ? V[r] : 0.0;
    for (int64_t cl = threadIdx.x; cl < (int64_t)ntc * 128; cl += blockDim.x) {
        const int64_t c = ktab[cl >> 7] + (cl & 127);
        if (c > r) continue;
        double v;
        if (i < D.nc) v = c < D.up ? -T1l[i * ldt + cl] * vr : (V[c] * Kp[i * D.mp + (c - D.up)]) * vr;
        else v = (c == r) ? 1.0 : 0.0;
        row[cl] = v;
    }
}

__global__ void __launch_bounds__(128) k_add_point_blocks_cols(double *__restrict__ X, int64_t ldx, const int32_t *__restrict__ col_local,
                                                               const int32_t *__restrict__ blk_start, const int32_t *__restrict__ blk_size,
                                                               int nBlk, const double *__restrict__ Pinv,
                                                               const double *__restrict__ V) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlk) return;
    const int64_t c0 = blk_start[b];
    const int sz = blk_size[b];
    for (int j = 0; j < sz; j++) {
        const int64_t c = c0 + j;
        const int t = col_local[c >> 7];
        if (t < 0) continue;
        for (int i = j; i < sz; i++)
            X[(c0 + i) * ldx + (int64_t)t * 128 + (c & 127)] += (V[c] * Pinv[(size_t)b * 9 + i * 3 + j]) * V[c0 + i];
    }
}

void launch_structured_place_cols(double *X, int64_t ldx, int ntc, const int32_t *ktab, const int32_t *col_local, const StructDims &D,
                                  const double *T1l, int64_t ldt, const double *Kp, const int32_t *blk_start, const int32_t *blk_size,
                                  int nBlk, const double *Pinv, const double *V, cudaStream_t s) {
    if (ntc == 0) return;
    int64_t row_end = D.Tp < D.np ? D.Tp : D.np;
    if (row_end < D.u) row_end = D.u;
    g_launch_count += 2;
    k_place_cols<<<(unsigned)(row_end - D.up), 256, 0, s>>>(X, ldx, ntc, ktab, D, T1l, ldt, Kp, V, row_end);
    k_add_point_blocks_cols<<<(unsigned)((nBlk + 127) / 128), 128, 0, s>>>(X, ldx, col_local, blk_start, blk_size, nBlk, Pinv, V);
}

}  // namespace jaicov
