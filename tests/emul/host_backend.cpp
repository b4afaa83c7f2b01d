// host_backend.cpp -- TEST-ONLY host emulation of the dense backend.
//
// Instantiates the product's blocked schedule (bundle-adjustment_b200/csrc/dense_driver.hpp) with plain-loop tile
// operations so that the recursion / k-range / aliasing logic can be unit-tested without a GPU.  This file is never
// linked into libjaicov_b200.so; the product has no CPU path.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../bundle-adjustment_b200/csrc/dense_driver.hpp"

using namespace jaicov;

struct HostBackend {
    int info = 0;
    long gemm_calls = 0, diag_calls = 0;
    double flops = 0;

    void gemm(const GemmDesc &g) {
        gemm_calls++;
        const int T = kTile;
        std::vector<double> acc((size_t)T * T);
        for (int it = 0; it < g.mt; it++)
            for (int jt = 0; jt < g.nt; jt++) {
                if (g.tri_out && it < jt) continue;
                int64_t kbeg = 0, kend = g.K;
                if (g.kmode == K_B_LOWER) kbeg = (int64_t)jt * T;
                else if (g.kmode == K_A_LOWER) kend = std::min<int64_t>(g.K, (int64_t)(it + 1) * T);
                else if (g.kmode == K_MAX_IJ) kbeg = (int64_t)std::max(it, jt) * T;
                // read all inputs of the tile first (the CUDA kernel finishes its loads before it stores)
                for (int i = 0; i < T; i++)
                    for (int j = 0; j < T; j++) {
                        double s = 0.0;
                        for (int64_t k = kbeg; k < kend; k++) {
                            const double a = g.al == 0 ? g.A[((int64_t)it * T + i) * g.lda + k] : g.A[k * g.lda + (int64_t)it * T + i];
                            const double b = g.bl == 0 ? g.B[((int64_t)jt * T + j) * g.ldb + k] : g.B[k * g.ldb + (int64_t)jt * T + j];
                            s += a * b;
                        }
                        acc[(size_t)i * T + j] = s;
                    }
                flops += 2.0 * T * T * (double)(kend - kbeg);
                for (int i = 0; i < T; i++)
                    for (int j = 0; j < T; j++) {
                        double *c = g.C + ((int64_t)it * T + i) * g.ldc + (int64_t)jt * T + j;
                        *c = g.beta == 0.0 ? g.alpha * acc[(size_t)i * T + j] : g.alpha * acc[(size_t)i * T + j] + g.beta * *c;
                    }
            }
    }

    // factor the 128x128 diagonal block in place (lower) and write its inverse (upper part zero) to dinv
    void potrf_diag(double *a, int64_t ld, double *dinv, int row0) {
        diag_calls++;
        const int T = kTile;
        for (int j = 0; j < T; j++) {
            double s = a[j * ld + j];
            for (int k = 0; k < j; k++) s -= a[j * ld + k] * a[j * ld + k];
            if (!(s > 0.0) && info == 0) info = row0 + j + 1;
            const double d = std::sqrt(s);
            a[j * ld + j] = d;
            for (int i = j + 1; i < T; i++) {
                double t = a[i * ld + j];
                for (int k = 0; k < j; k++) t -= a[i * ld + k] * a[j * ld + k];
                a[i * ld + j] = t / d;
            }
        }
        for (int j = 0; j < T; j++) {
            for (int i = 0; i < T; i++) dinv[i * T + j] = 0.0;
            dinv[j * T + j] = 1.0 / a[j * ld + j];
            for (int i = j + 1; i < T; i++) {
                double s = 0.0;
                for (int k = j; k < i; k++) s += a[i * ld + k] * dinv[k * T + j];
                dinv[i * T + j] = -s / a[i * ld + i];
            }
        }
    }

    void copy2d(double *dst, int64_t ldd, const double *src, int64_t lds, int64_t rows, int64_t cols) {
        for (int64_t r = 0; r < rows; r++) std::memcpy(dst + r * ldd, src + r * lds, sizeof(double) * cols);
    }
};

extern "C" {

// M: np x np row-major; on entry the LOWER triangle holds an SPD matrix (upper is poisoned here with NaN);
// R: mt*128 x np right-hand-side rows (solved in place); on exit M lower = inverse.
int emul_spd_solve_invert(int64_t np, double *M, int mt, double *R, int invert, double *stats) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t r = 0; r < np; r++)
        for (int64_t c = r + 1; c < np; c++)
            if ((r / kTile) != (c / kTile)) M[r * np + c] = nan;   // strictly-upper off-diagonal tiles are never read
    std::vector<double> Dinv((size_t)np * kTile, nan), W;
    HostBackend be;
    DenseSchedule<HostBackend> ds{be, M, np, np, Dinv.data()};
    ds.potrf();
    if (mt > 0) ds.solve_rows(R, np, mt);
    if (invert) {
        W.assign((size_t)np * np, nan);
        ds.invert_from_factor(W.data());
    }
    if (stats) { stats[0] = (double)be.gemm_calls; stats[1] = (double)be.diag_calls; stats[2] = be.flops; }
    return be.info;
}
}
