// propagate.cu -- covariance propagation on the device-resident cofactor matrix: Sigma = sigma^2 J Qxx J'  (SURVEY.md 8 f-3).
//
// Replaces the two sparse-times-packed products of CoordinateTransformationExteriorOrientation.transform
// (tranformation/CoordinateTransformationExteriorOrientation.java:110-114: CoVar.transBmult(J, CJT); J.mult(sigma2, CJT, covariance)).
// J has 3 rows per transformed point and at most 15 non-zeros per row (6 + 6 exterior-orientation columns, 3 object
// coordinates), so every 3 x 3 block of Sigma needs a 15 x 15 gather from Qxx: one thread per block pair, no
// intermediate n x 3R matrix.  Qxx is read where it lives (lower triangle of M, or a rank's column tiles).
#include "common.h"

namespace jaicov {

struct QxxView {
    const double *lower;        // single GPU: np x np lower triangle (internal index = reference column - d), else nullptr
    int64_t ld;
    const double *X;            // multi-GPU: this rank's column tiles
    int64_t ldx;
    const int32_t *col_local;   // multi-GPU: global tile -> local tile or -1
    const double *border;       // d x np rows K^-1[lambda, x]
    int64_t np;
    const double *q11;          // d x d (pitch kMaxDatum)
    int d;
    int rank;
};

// entry (r, c) of the full symmetric matrix in reference column numbering; multi-GPU: 0 for entries owned elsewhere
__device__ __forceinline__ double qxx_at(const QxxView &q, int r, int c) {
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    if (hi < q.d) return q.rank == 0 ? q.q11[lo * kMaxDatum + hi] : 0.0;
    if (lo < q.d) return q.rank == 0 ? q.border[(int64_t)lo * q.np + (hi - q.d)] : 0.0;
    if (q.lower) return q.lower[(int64_t)(hi - q.d) * q.ld + (lo - q.d)];
    const int e = lo - q.d;
    const int jl = q.col_local[e >> 7];
    return jl >= 0 ? q.X[(int64_t)(hi - q.d) * q.ldx + jl * 128 + (e & 127)] : 0.0;
}

constexpr int kJc = 15;   // non-zero columns per transformed point

// One thread per pair (i1 >= i2) of transformed points: B = sigma2 * J_i2 G J_i1' with G = Qxx[cols(i2), cols(i1)];
// written to the packed upper triangle (MTJ UpperSymmPackMatrix: (r, c), r <= c, at r + c (c + 1) / 2).
__global__ void __launch_bounds__(128) k_propagate(QxxView q, int nT, const double *__restrict__ Jv, const int32_t *__restrict__ Jc,
                                                   double sigma2, double *__restrict__ out) {
    const int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)nT * (nT + 1) / 2;
    if (pair >= total) return;
    int i1 = (int)((sqrt(8.0 * (double)pair + 1.0) - 1.0) * 0.5);
    while ((int64_t)(i1 + 1) * (i1 + 2) / 2 <= pair) i1++;
    while ((int64_t)i1 * (i1 + 1) / 2 > pair) i1--;
    const int i2 = (int)(pair - (int64_t)i1 * (i1 + 1) / 2);      // i2 <= i1
    const double *J1 = Jv + (size_t)i1 * 3 * kJc, *J2 = Jv + (size_t)i2 * 3 * kJc;
    const int32_t *c1 = Jc + (size_t)i1 * kJc, *c2 = Jc + (size_t)i2 * kJc;
    // t[k][b] = sum_a J2[k][a] G[a][b]
    double t[3][kJc];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int b = 0; b < kJc; b++) t[k][b] = 0.0;
    for (int a = 0; a < kJc; a++) {
        const int ca = c2[a];
        if (ca < 0) continue;
        const double j0 = J2[a], j1 = J2[kJc + a], j2 = J2[2 * kJc + a];
#pragma unroll
        for (int b = 0; b < kJc; b++) {
            const int cb = c1[b];
            if (cb < 0) continue;
            const double g = qxx_at(q, ca, cb);
            t[0][b] += j0 * g; t[1][b] += j1 * g; t[2][b] += j2 * g;
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)          // row 3 i2 + k
#pragma unroll
        for (int l = 0; l < 3; l++) {    // column 3 i1 + l
            if (i1 == i2 && k > l) continue;
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < kJc; b++) s += t[k][b] * J1[l * kJc + b];
            const int64_t r = 3 * (int64_t)i2 + k, c = 3 * (int64_t)i1 + l;
            out[r + c * (c + 1) / 2] = sigma2 * s;
        }
}

void launch_propagate(const double *lower, int64_t ld, const double *X, int64_t ldx, const int32_t *col_local, const double *border,
                      int64_t np, const double *q11, int d, int rank, int nT, const double *Jv, const int32_t *Jc, double sigma2,
                      double *out, cudaStream_t s) {
    if (nT <= 0) return;
    QxxView q{lower, ld, X, ldx, col_local, border, np, q11, d, rank};
    const int64_t total = (int64_t)nT * (nT + 1) / 2;
    g_launch_count++;
    k_propagate<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(q, nT, Jv, Jc, sigma2, out);
}

}  // namespace jaicov
