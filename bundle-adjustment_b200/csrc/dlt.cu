// dlt.cu -- batched direct linear transformation: initial interior / exterior orientation of many images at once
// (SURVEY.md 8 f-4).  Replaces DirectLinearTransformation.adjust (dlt/DirectLinearTransformation.java:67-184) called
// image by image: one warp per image accumulates the 11 x 11 normal equations of the linear model
//     x = X b11 + Y b12 + Z b13 + b14 - x (X b31 + Y b32 + Z b33)          (dlt/DLTPartialDerivativeFactory.java:62-65)
// over the image's homologous points (lanes stride the points, fixed-order shuffle reduction), then lane 0 runs the
// reference's iteration: first pass without, later passes with the restrictions as border rows (DLT:116-165),
// Jacobi preconditioner (DLT:341-347), a dense pivoted solve of the <= 17 x 17 bordered system, back-scaling and the
// orientation parameters (DLT:186-266).  The design matrix does not depend on the coefficients, so the point sweep is
// done once: n = A'l - N b in the later passes.
#include "common.h"

namespace jaicov {

constexpr int kDltB = 11;        // coefficients
constexpr int kDltMaxR = 6;      // restrictions
constexpr int kDltN = kDltB + kDltMaxR;
constexpr int kDltWarps = 4;     // images per CTA

struct DltWork {
    double N[kDltB][kDltB];      // A'A (full symmetric)
    double Al[kDltB];            // A'l
    double K[kDltN][kDltN + 1];  // bordered, preconditioned system | right-hand side
};

__device__ __forceinline__ void dlt_rows(double x, double y, double X, double Y, double Z, double a0[kDltB], double a1[kDltB]) {
    a0[0] = X; a0[1] = Y; a0[2] = Z; a0[3] = 1.0; a0[4] = a0[5] = a0[6] = a0[7] = 0.0; a0[8] = -x * X; a0[9] = -x * Y; a0[10] = -x * Z;
    a1[0] = a1[1] = a1[2] = a1[3] = 0.0; a1[4] = X; a1[5] = Y; a1[6] = Z; a1[7] = 1.0; a1[8] = -y * X; a1[9] = -y * Y; a1[10] = -y * Z;
}

__device__ inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// gradient (11) and misclosure of one restriction (DPF:68-236), r1, r2, r3 = rows of the 3 x 3 part of B
__device__ void dlt_restriction(int kind, const double *b, double c, double x0, double y0, double g[kDltB], double &w) {
    const double *r1 = b, *r2 = b + 4, *r3 = b + 8;
    const double b1 = dot3(r1, r1), b2 = dot3(r2, r2), b3 = dot3(r3, r3), bx = dot3(r1, r3), by = dot3(r2, r3);
    for (int k = 0; k < kDltB; k++) g[k] = 0.0;
    switch (kind) {
    case 4:   // FIXED_PRINCIPAL_POINT_X: x0 = bx / b3
        for (int k = 0; k < 3; k++) { g[k] = r3[k] / b3; g[8 + k] = r1[k] / b3 - 2.0 * bx * r3[k] / (b3 * b3); }
        w = x0 - bx / b3;
        break;
    case 5:   // FIXED_PRINCIPAL_POINT_Y
        for (int k = 0; k < 3; k++) { g[4 + k] = r3[k] / b3; g[8 + k] = r2[k] / b3 - 2.0 * by * r3[k] / (b3 * b3); }
        w = y0 - by / b3;
        break;
    case 2:   // FIXED_PRINCIPLE_DISTANCE_X: c^2 = b1 / b3 - bx^2 / b3^2
        for (int k = 0; k < 3; k++) {
            g[k] = 2.0 * (r1[k] * b3 - bx * r3[k]) / (b3 * b3);
            g[8 + k] = 4.0 * (r3[k] * bx * bx - 0.5 * b3 * (r3[k] * b1 + bx * r1[k])) / (b3 * b3 * b3);
        }
        w = c * c - b1 / b3 + bx * bx / (b3 * b3);
        break;
    case 3:   // FIXED_PRINCIPLE_DISTANCE_Y
        for (int k = 0; k < 3; k++) {
            g[4 + k] = 2.0 * (r2[k] * b3 - by * r3[k]) / (b3 * b3);
            g[8 + k] = 4.0 * (r3[k] * by * by - 0.5 * b3 * (r3[k] * b2 + by * r2[k])) / (b3 * b3 * b3);
        }
        w = c * c - b2 / b3 + by * by / (b3 * b3);
        break;
    case 0:   // IDENTICAL_PRINCIPLE_DISTANCE: b3 (b1 - b2) - bx^2 + by^2 = 0
        for (int k = 0; k < 3; k++) {
            g[k] = 2.0 * (b3 * r1[k] - bx * r3[k]);
            g[4 + k] = -2.0 * (b3 * r2[k] - by * r3[k]);
            g[8 + k] = 2.0 * (r3[k] * (b1 - b2) - bx * r1[k] + by * r2[k]);
        }
        w = -b3 * (b1 - b2) + bx * bx - by * by;
        break;
    default:  // 1, ROTATION_WITHOUT_SHEAR: -b3 (r1 . r2) + bx by = 0
        for (int k = 0; k < 3; k++) {
            g[k] = -b3 * r2[k] + by * r3[k];
            g[4 + k] = -b3 * r1[k] + bx * r3[k];
            g[8 + k] = -2.0 * r3[k] * dot3(r1, r2) + by * r1[k] + bx * r2[k];
        }
        w = b3 * dot3(r1, r2) - bx * by;
        break;
    }
}

// Gaussian elimination with partial pivoting on the n x (n + 1) augmented system; false if a pivot vanishes
__device__ bool dlt_solve(double K[kDltN][kDltN + 1], int n) {
    for (int c = 0; c < n; c++) {
        int pr = c;
        for (int i = c + 1; i < n; i++)
            if (fabs(K[i][c]) > fabs(K[pr][c])) pr = i;
        if (!(fabs(K[pr][c]) > 0.0)) return false;
        if (pr != c)
            for (int j = c; j <= n; j++) { const double t = K[c][j]; K[c][j] = K[pr][j]; K[pr][j] = t; }
        for (int i = c + 1; i < n; i++) {
            const double f = K[i][c] / K[c][c];
            for (int j = c; j <= n; j++) K[i][j] -= f * K[c][j];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = K[i][n];
        for (int j = i + 1; j < n; j++) s -= K[i][j] * K[j][n];
        K[i][n] = s / K[i][i];
    }
    return true;
}

// out (per image, 20 doubles): b11..b33 (11, back-scaled), c = (cx + cy) / 2, x0, y0, X0, Y0, Z0, omega, phi, kappa
// status: 1 converged, 0 no convergence within max_iterations, -1 failed (< 6 points, singular system, NaN)
__global__ void __launch_bounds__(32 * kDltWarps) k_dlt_batch(int n_img, const int64_t *__restrict__ pt_ptr, const double *__restrict__ xy,
                                                              const double *__restrict__ XYZ, const double *__restrict__ io, int nR,
                                                              const int32_t *__restrict__ restr, int max_iterations,
                                                              double *__restrict__ out, int32_t *__restrict__ status,
                                                              int32_t *__restrict__ passes_out) {
    __shared__ DltWork work[kDltWarps];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int img = blockIdx.x * kDltWarps + wl;
    if (img >= n_img) return;
    DltWork &W = work[wl];
    const int64_t p0 = pt_ptr[img], p1 = pt_ptr[img + 1];
    // ---- coordinate scale (DLT:75-106) -------------------------------------------------------------------------------
    double sw = 0.0, si = 0.0;
    for (int64_t p = p0 + lane; p < p1; p += 32) {
        sw += XYZ[3 * p] * XYZ[3 * p] + XYZ[3 * p + 1] * XYZ[3 * p + 1] + XYZ[3 * p + 2] * XYZ[3 * p + 2];
        si += xy[2 * p] * xy[2 * p] + xy[2 * p + 1] * xy[2 * p + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sw += __shfl_xor_sync(0xffffffffu, sw, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
    const double scale = si > 0.0 ? sqrt(sw / si) : 1.0;
    // ---- A'A and A'l, one sweep (DPF:239-337 with b = 0) -----------------------------------------------------------------
    double acc[kDltB * (kDltB + 1) / 2 + kDltB];
#pragma unroll
    for (int k = 0; k < kDltB * (kDltB + 1) / 2 + kDltB; k++) acc[k] = 0.0;
    for (int64_t p = p0 + lane; p < p1; p += 32) {
        double a0[kDltB], a1[kDltB];
        const double x = xy[2 * p], y = xy[2 * p + 1];
        dlt_rows(x, y, XYZ[3 * p] / scale, XYZ[3 * p + 1] / scale, XYZ[3 * p + 2] / scale, a0, a1);
        int k = 0;
#pragma unroll
        for (int i = 0; i < kDltB; i++)
#pragma unroll
            for (int j = i; j < kDltB; j++) acc[k++] += a0[i] * a0[j] + a1[i] * a1[j];
#pragma unroll
        for (int i = 0; i < kDltB; i++) acc[k++] += a0[i] * x + a1[i] * y;
    }
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < kDltB; i++)
#pragma unroll
            for (int j = i; j < kDltB; j++) {
                double v = acc[k++];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) { W.N[i][j] = v; W.N[j][i] = v; }
            }
#pragma unroll
        for (int i = 0; i < kDltB; i++) {
            double v = acc[k++];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) W.Al[i] = v;
        }
    }
    if (lane != 0) return;
    double *res = out + (size_t)img * 20;
    for (int k = 0; k < 20; k++) res[k] = 0.0;
    passes_out[img] = 0;
    if (p1 - p0 < 6) { status[img] = -1; return; }          // DLT:96-104
    const double c = io[3 * img], x0 = io[3 * img + 1], y0 = io[3 * img + 2];
    // ---- iteration (DLT:108-180) -------------------------------------------------------------------------------------------
    double b[kDltB];
    for (int k = 0; k < kDltB; k++) b[k] = 0.0;
    int runs = max_iterations - 1, passes = 0;
    bool is_estimated = max_iterations == 0, complete = is_estimated, is_converge = true, include = false;
    while (true) {
        passes++;
        const int R = include ? nR : 0, n = kDltB + R;
        double V[kDltN];
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= n; j++) W.K[i][j] = 0.0;
        for (int i = 0; i < kDltB; i++) {
            double s = W.Al[i];
            for (int j = 0; j < kDltB; j++) { W.K[i][j] = W.N[i][j]; s -= W.N[i][j] * b[j]; }
            W.K[i][n] = s;                                   // n = A'(l - A b)
        }
        for (int r = 0; r < R; r++) {
            double g[kDltB], w;
            dlt_restriction(restr[r], b, c, x0, y0, g, w);
            for (int k = 0; k < kDltB; k++) { W.K[k][kDltB + r] = g[k]; W.K[kDltB + r][k] = g[k]; }
            W.K[kDltB + r][n] = w;
        }
        for (int i = 0; i < n; i++) V[i] = W.K[i][i] > kEps ? 1.0 / sqrt(W.K[i][i]) : 1.0;     // DLT:341-347
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < n; j++) W.K[i][j] = (V[i] * W.K[i][j]) * V[j];
            W.K[i][n] *= V[i];
        }
        complete = is_estimated || nR == 0;
        if (!dlt_solve(W.K, n)) { status[img] = -1; passes_out[img] = passes; return; }
        double mx = 0.0;
        bool bad = false;
        for (int k = 0; k < kDltB; k++) {
            const double dv = V[k] * W.K[k][n];
            if (isnan(dv) || isinf(dv)) bad = true;
            mx = fmax(mx, fabs(dv));
            b[k] += dv;
        }
        include = true;
        if (bad) { status[img] = -1; passes_out[img] = passes; return; }
        if (mx <= sqrt(kEps) && runs > 0) is_estimated = true;
        else if (runs-- <= 1) {
            if (complete) is_converge = false;
            is_estimated = true;
        }
        if (complete) break;
    }
    // ---- back-scaling and orientation (DLT:186-266) ---------------------------------------------------------------------------
    for (int k = 0; k < kDltB; k++)
        if (k != 3 && k != 7) b[k] /= scale;
    const double *r1 = b, *r2 = b + 4, *r3 = b + 8;
    const double bb = dot3(r3, r3), sb = sqrt(bb);
    const double px = dot3(r1, r3) / bb, py = dot3(r2, r3) / bb;
    const double cx = sqrt(dot3(r1, r1) / bb - px * px), cy = sqrt(dot3(r2, r2) / bb - py * py);
    double Rm[3][3];
    for (int k = 0; k < 3; k++) {
        Rm[k][0] = -(px * r3[k] - r1[k]) / sb / cx;
        Rm[k][1] = -(py * r3[k] - r2[k]) / sb / cy;
        Rm[k][2] = -r3[k] / sb;
    }
    const double det = Rm[0][0] * Rm[1][1] * Rm[2][2] + Rm[0][1] * Rm[1][2] * Rm[2][0] + Rm[0][2] * Rm[1][0] * Rm[2][1] -
                       Rm[0][2] * Rm[1][1] * Rm[2][0] - Rm[0][0] * Rm[1][2] * Rm[2][1] - Rm[0][1] * Rm[1][0] * Rm[2][2];
    if (det < 0)
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) Rm[i][j] = -Rm[i][j];
    // projection centre: F t = f with F = rows r1, r2, r3 and f = (-b14, -b24, -1)
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) W.K[i][j] = b[4 * i + j];
        W.K[i][3] = i == 0 ? -b[3] : (i == 1 ? -b[7] : -1.0);
    }
    // (3 x 3 with the same pivoted elimination; the augmented column sits at index n = 3)
    {
        double (*K3)[kDltN + 1] = W.K;
        bool ok = true;
        for (int cidx = 0; cidx < 3 && ok; cidx++) {
            int pr = cidx;
            for (int i = cidx + 1; i < 3; i++)
                if (fabs(K3[i][cidx]) > fabs(K3[pr][cidx])) pr = i;
            if (!(fabs(K3[pr][cidx]) > 0.0)) { ok = false; break; }
            if (pr != cidx)
                for (int j = 0; j <= 3; j++) { const double t = K3[cidx][j]; K3[cidx][j] = K3[pr][j]; K3[pr][j] = t; }
            for (int i = cidx + 1; i < 3; i++) {
                const double f = K3[i][cidx] / K3[cidx][cidx];
                for (int j = cidx; j <= 3; j++) K3[i][j] -= f * K3[cidx][j];
            }
        }
        if (!ok) { status[img] = -1; passes_out[img] = passes; return; }
        for (int i = 2; i >= 0; i--) {
            double s = K3[i][3];
            for (int j = i + 1; j < 3; j++) s -= K3[i][j] * K3[j][3];
            K3[i][3] = s / K3[i][i];
        }
    }
    for (int k = 0; k < kDltB; k++) res[k] = b[k];
    res[11] = 0.5 * (cx + cy); res[12] = px; res[13] = py;
    res[14] = W.K[0][3]; res[15] = W.K[1][3]; res[16] = W.K[2][3];
    res[17] = atan2(-Rm[1][2], Rm[2][2]); res[18] = asin(Rm[0][2]); res[19] = atan2(-Rm[0][1], Rm[0][0]);
    status[img] = is_converge ? 1 : 0;
    passes_out[img] = passes;
}

void launch_dlt_batch(int n_img, const int64_t *pt_ptr, const double *xy, const double *XYZ, const double *io, int nR,
                      const int32_t *restr, int max_iterations, double *out, int32_t *status, int32_t *passes, cudaStream_t s) {
    if (n_img <= 0) return;
    g_launch_count++;
    k_dlt_batch<<<(unsigned)((n_img + kDltWarps - 1) / kDltWarps), 32 * kDltWarps, 0, s>>>(n_img, pt_ptr, xy, XYZ, io, nR, restr,
                                                                                          max_iterations, out, status, passes);
}

}  // namespace jaicov
