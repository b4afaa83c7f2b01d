// dense_driver.hpp -- blocked FP64 factor / solve / invert schedule (K5, K6 of SURVEY.md section 2.1).
//
// Replaces what the reference gets from LAPACK dspsv + dsptri (MathExtension.java:338-366) and dpptrf + dpptri
// (MathExtension.java:304-324) with a Cholesky route on an SPD matrix (the bordered, indefinite datum system is
// reduced to an SPD one in api.cu, see DESIGN.md "datum identity").
//
// Everything O(n^3) is expressed as 128x128-tile GEMMs so that one tensor-core kernel (dense_kernels.cu) does all
// the work; the only non-GEMM kernel is the 128x128 diagonal-block factor+invert.  Storage is row-major, lower
// triangle; all dimensions are multiples of 128 (the system is padded with an identity block).
//
//   potrf   M = L L'            recursive right-looking:  L21 = M21 L11^-T (TRSM),  M22 -= L21 L21' (SYRK)
//   trsm    X = B L^-T / B L^-1 recursive; the 128-wide base case multiplies by the explicitly inverted
//                               diagonal block Dinv (computed by the diagonal kernel), in place
//   trtri   W = L^-1            recursive, out of place into a second buffer W:
//                               W21 = -W22 (L21 W11); triangular operands skip their zero k-tiles
//   lauum   M = W' W = M^-1     ONE launch: every lower output tile (i,j) contracts k >= max(i,j)
//
// The schedule is a template over a Backend so that the index logic can be exercised on the host by the unit
// tests (tests/emul/host_backend.cpp, test-only); the product instantiates it with the CUDA backend only.
#pragma once
#include <cstdint>

namespace jaicov {

constexpr int kTile = 128;

enum KMode : int {
    K_FULL = 0,       // contract over all of [0, K)
    K_B_LOWER = 1,    // B[k][n] lower triangular (k >= n): start at the tile column's diagonal
    K_A_LOWER = 2,    // A[m][k] lower triangular (k <= m): stop after the tile row's diagonal
    K_MAX_IJ = 3      // A[k][m], B[k][n] both lower triangular: k >= max(m, n)
};

struct GemmDesc {
    int al = 0;       // 0: A is [m][k] (k contiguous); 1: A is [k][m] (m contiguous)
    int bl = 0;       // 0: B is [n][k] (k contiguous); 1: B is [k][n] (n contiguous)
    int mt = 0, nt = 0;   // output tiles of 128 x 128
    int64_t K = 0;        // contraction length, multiple of 128
    double alpha = 1.0, beta = 0.0;
    const double *A = nullptr;
    int64_t lda = 0;
    const double *B = nullptr;
    int64_t ldb = 0;
    double *C = nullptr;
    int64_t ldc = 0;
    int tri_out = 0;  // only tiles with tile-row >= tile-col (mt == nt, C on the diagonal)
    int kmode = K_FULL;
};

template <class BE>
struct DenseSchedule {
    BE &be;
    double *M;        // np x np, row-major, lower triangle
    int64_t ld;
    int64_t np;
    double *Dinv;     // [np/128] blocks of 128 x 128: inverses of the diagonal blocks of L (upper part zero)

    int nblocks() const { return (int)(np / kTile); }

    // ---- Cholesky ---------------------------------------------------------------------------------------------
    void potrf(int64_t r0, int nb) {
        if (nb == 1) {
            be.potrf_diag(M + r0 * ld + r0, ld, Dinv + r0 * kTile, (int)(r0));
            return;
        }
        const int hb = nb / 2, rest = nb - hb;
        const int64_t h = (int64_t)hb * kTile;
        potrf(r0, hb);
        trsm_rlt(M + (r0 + h) * ld, ld, rest, r0, hb);
        GemmDesc g;
        g.al = 0; g.bl = 0; g.mt = rest; g.nt = rest; g.K = h; g.alpha = -1.0; g.beta = 1.0;
        g.A = M + (r0 + h) * ld + r0; g.lda = ld;
        g.B = g.A; g.ldb = ld;
        g.C = M + (r0 + h) * ld + (r0 + h); g.ldc = ld;
        g.tri_out = 1;
        be.gemm(g);
        potrf(r0 + h, rest);
    }
    void potrf() { potrf(0, nblocks()); }

    // B[:, c0 : c0+w) <- B[:, c0 : c0+w) * L[c0.., c0..]^-T   (B has mt*128 rows, leading dimension ldb)
    void trsm_rlt(double *B, int64_t ldb, int mt, int64_t c0, int wb) {
        if (wb == 1) {
            GemmDesc g;
            g.al = 0; g.bl = 0; g.mt = mt; g.nt = 1; g.K = kTile; g.alpha = 1.0; g.beta = 0.0;
            g.A = B + c0; g.lda = ldb;
            g.B = Dinv + c0 * kTile; g.ldb = kTile;
            g.C = B + c0; g.ldc = ldb;   // in place: one CTA owns the full 128-wide strip rows it reads
            be.gemm(g);
            return;
        }
        const int hb = wb / 2;
        const int64_t h = (int64_t)hb * kTile;
        trsm_rlt(B, ldb, mt, c0, hb);
        GemmDesc g;
        g.al = 0; g.bl = 0; g.mt = mt; g.nt = wb - hb; g.K = h; g.alpha = -1.0; g.beta = 1.0;
        g.A = B + c0; g.lda = ldb;
        g.B = M + (c0 + h) * ld + c0; g.ldb = ld;
        g.C = B + c0 + h; g.ldc = ldb;
        be.gemm(g);
        trsm_rlt(B, ldb, mt, c0 + h, wb - hb);
    }

    // B[:, c0 : c0+w) <- B[:, c0 : c0+w) * L[c0.., c0..]^-1
    void trsm_rln(double *B, int64_t ldb, int mt, int64_t c0, int wb) {
        if (wb == 1) {
            GemmDesc g;
            g.al = 0; g.bl = 1; g.mt = mt; g.nt = 1; g.K = kTile; g.alpha = 1.0; g.beta = 0.0;
            g.A = B + c0; g.lda = ldb;
            g.B = Dinv + c0 * kTile; g.ldb = kTile;
            g.C = B + c0; g.ldc = ldb;
            be.gemm(g);
            return;
        }
        const int hb = wb / 2;
        const int64_t h = (int64_t)hb * kTile;
        trsm_rln(B, ldb, mt, c0 + h, wb - hb);
        GemmDesc g;
        g.al = 0; g.bl = 1; g.mt = mt; g.nt = hb; g.K = (int64_t)(wb - hb) * kTile; g.alpha = -1.0; g.beta = 1.0;
        g.A = B + c0 + h; g.lda = ldb;
        g.B = M + (c0 + h) * ld + c0; g.ldb = ld;
        g.C = B + c0; g.ldc = ldb;
        be.gemm(g);
        trsm_rln(B, ldb, mt, c0, hb);
    }

    // right-hand sides are the ROWS of R (mt*128 rows x np): R <- R * L^-T * L^-1 = (M^-1 R')'
    void solve_rows(double *R, int64_t ldr, int mt) {
        trsm_rlt(R, ldr, mt, 0, nblocks());
        trsm_rln(R, ldr, mt, 0, nblocks());
    }

    // ---- inverse of the factor, out of place: W (lower) = L^-1; destroys the strictly lower part of M ----------
    void trtri(double *W, int64_t r0, int nb) {
        if (nb == 1) {
            be.copy2d(W + r0 * ld + r0, ld, Dinv + r0 * kTile, kTile, kTile, kTile);
            return;
        }
        const int hb = nb / 2, rest = nb - hb;
        const int64_t h = (int64_t)hb * kTile;
        trtri(W, r0, hb);
        trtri(W, r0 + h, rest);
        GemmDesc g;   // X = L21 * W11  -> W21
        g.al = 0; g.bl = 1; g.mt = rest; g.nt = hb; g.K = h; g.alpha = 1.0; g.beta = 0.0;
        g.A = M + (r0 + h) * ld + r0; g.lda = ld;
        g.B = W + r0 * ld + r0; g.ldb = ld;
        g.C = W + (r0 + h) * ld + r0; g.ldc = ld;
        g.kmode = K_B_LOWER;
        be.gemm(g);
        GemmDesc q;   // M21 = -W22 * X
        q.al = 0; q.bl = 1; q.mt = rest; q.nt = hb; q.K = (int64_t)rest * kTile; q.alpha = -1.0; q.beta = 0.0;
        q.A = W + (r0 + h) * ld + (r0 + h); q.lda = ld;
        q.B = W + (r0 + h) * ld + r0; q.ldb = ld;
        q.C = M + (r0 + h) * ld + r0; q.ldc = ld;
        q.kmode = K_A_LOWER;
        be.gemm(q);
        be.copy2d(W + (r0 + h) * ld + r0, ld, M + (r0 + h) * ld + r0, ld, (int64_t)rest * kTile, h);
    }

    // M (lower) <- W' W
    void lauum(const double *W) {
        GemmDesc g;
        g.al = 1; g.bl = 1; g.mt = nblocks(); g.nt = nblocks(); g.K = np; g.alpha = 1.0; g.beta = 0.0;
        g.A = W; g.lda = ld; g.B = W; g.ldb = ld; g.C = M; g.ldc = ld;
        g.tri_out = 1; g.kmode = K_MAX_IJ;
        be.gemm(g);
    }

    void invert_from_factor(double *W) {
        trtri(W, 0, nblocks());
        lauum(W);
    }
};

}  // namespace jaicov
