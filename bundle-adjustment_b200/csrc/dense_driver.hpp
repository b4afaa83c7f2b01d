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
#include <algorithm>
#include <cmath>
#include <cstdint>

namespace jaicov {

constexpr int kTile = 128;

enum KMode : int {
    K_FULL = 0,       // contract over all of [0, K)
    K_B_LOWER = 1,    // B[k][n] lower triangular (k >= n): start at the tile column's diagonal
    K_A_LOWER = 2,    // A[m][k] lower triangular (k <= m): stop after the tile row's diagonal
    K_MAX_IJ = 3,     // A[k][m], B[k][n] both lower triangular: k >= max(m, n)
    K_COL_BEG = 4,    // column tile jt of B is zero above global row ktab[jt]: start at ktab[jt] - koff, skip the tile if >= K
    K_ROW_MASK = 5    // output tile (it, jt) is only wanted if its global row roff + 128 it >= ktab[jt]; full contraction
};

#if defined(__CUDACC__)
#define JAICOV_HD __host__ __device__
#else
#define JAICOV_HD
#endif

// Tile (it >= jt) of a lower-triangular launch of mt x mt tiles for the linear CTA index l.
// band == 0: row by row (row it, columns 0 .. it) -- the order every measured number of round 1 was taken with.
// band  > 0: bands of `band` tile rows, column by column inside a band, so that the CTAs resident together share the
//            column operand strip and keep the band's row strips in L2 (experiment for the big symmetric products,
//            JAICOV_TILE_BAND).  Both orders visit every lower tile exactly once and keep
//            "smaller max(it, jt) first" up to the band height (K_MAX_IJ launches start their longest contractions first).
JAICOV_HD inline void tri_tile_decode(int64_t l, int mt, int band, int &it, int &jt) {
    int row = (int)((sqrt(8.0 * (double)l + 1.0) - 1.0) * 0.5);
    while ((int64_t)(row + 1) * (row + 2) / 2 <= l) row++;
    while ((int64_t)row * (row + 1) / 2 > l) row--;
    if (band <= 0) {
        it = row;
        jt = (int)(l - (int64_t)row * (row + 1) / 2);
        return;
    }
    const int r0 = row / band * band;
    const int h = (mt - r0 < band) ? mt - r0 : band;
    int64_t idx = l - (int64_t)r0 * (r0 + 1) / 2;
    const int64_t rect = (int64_t)(r0 + 1) * h;        // columns 0 .. r0 hold all h rows of the band
    if (idx < rect) {
        jt = (int)(idx / h);
        it = r0 + (int)(idx % h);
        return;
    }
    idx -= rect;
    for (int j = r0 + 1; j < r0 + h; j++) {            // the band's own triangle
        const int cnt = r0 + h - j;
        if (idx < cnt) { jt = j; it = j + (int)idx; return; }
        idx -= cnt;
    }
    it = jt = mt - 1;                                  // not reached for l < mt (mt + 1) / 2
}

// Tile (it, jt) of a full (rectangular) launch of mt x nt tiles for the linear CTA index l: bands of `band` tile rows, column by
// column inside a band (band <= 0: row by row).  The CTAs resident together then cover a roughly square patch of C and share few
// operand strips (used by the int8-digit tile kernel, which is L2 / DRAM bound without it; ozaki.cu).
JAICOV_HD inline void rect_tile_decode(int64_t l, int mt, int nt, int band, int &it, int &jt) {
    if (band <= 0) {
        it = (int)(l / nt);
        jt = (int)(l - (int64_t)it * nt);
        return;
    }
    const int64_t per = (int64_t)band * nt;
    const int grp = (int)(l / per), first = grp * band;
    const int gsz = (mt - first < band) ? mt - first : band;
    const int64_t rem = l - grp * per;
    it = first + (int)(rem % gsz);
    jt = (int)(rem / gsz);
}

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
    int tile_band = 0;  // tri_out launches: tile order, see tri_tile_decode (0 = row by row)
    int kmode = K_FULL;
    const int32_t *coltab = nullptr; // if set: 2-D launch, blockIdx.y picks the global column tile coltab[y] / 128, blockIdx.x the
    int ncoltab = 0;                 //         global row tile; tiles above the diagonal exit (trapezoid update of many panels at once)
    int coltab_full = 0;             // coltab launches: every row tile is wanted (no trapezoid skip)
    int c_local = 0;                 // coltab launches: C is compact -- column tile blockIdx.y of C holds global tile coltab[y] / 128
    const int32_t *ktab = nullptr;   // K_COL_BEG / K_ROW_MASK: per column tile, first global row of interest
    int64_t koff = 0;                // K_COL_BEG: global row of k = 0
    int64_t roff = 0;                // K_ROW_MASK: global row of output tile row 0
    int64_t row_min = 0;             // column-table launches: no wanted tile lies above this global row (host-known): operand rows above it
                                     // are never read -- with owner-only storage they do not exist (virtual panel base)
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

    // ---- column-panel inverse: X = (L L')^-1 E for a SUBSET of 128-wide column tiles ------------------------------
    // X is np x (128 ntc), row-major with leading dimension ldx; column tile jl holds global columns
    // [ktab[jl], ktab[jl]+128) and starts as the corresponding columns of the identity.  Two wide triangular
    // sweeps; the zero structure (forward: rows above the column's diagonal are zero; backward: only rows on or
    // below it are wanted, Qxx is symmetric) is skipped tile by tile, so the work is 2/3 n^3 * (ntc / nblocks).
    // No communication: with L replicated every GPU inverts its own column tiles (multi-GPU).
    // X <- L^-1 X
    void trsm_lln(double *X, int64_t ldx, int ntc, const int32_t *ktab, int64_t r0, int nb) {
        if (nb == 1) {
            GemmDesc g;
            g.al = 0; g.bl = 1; g.mt = 1; g.nt = ntc; g.K = kTile; g.alpha = 1.0; g.beta = 0.0;
            g.A = Dinv + r0 * kTile; g.lda = kTile;
            g.B = X + r0 * ldx; g.ldb = ldx;
            g.C = X + r0 * ldx; g.ldc = ldx;     // in place: tile (0, jt) reads and writes the same 128 x 128 block
            g.kmode = K_COL_BEG; g.ktab = ktab; g.koff = r0;
            be.gemm(g);
            return;
        }
        const int hb = nb / 2, rest = nb - hb;
        const int64_t h = (int64_t)hb * kTile;
        trsm_lln(X, ldx, ntc, ktab, r0, hb);
        GemmDesc g;
        g.al = 0; g.bl = 1; g.mt = rest; g.nt = ntc; g.K = h; g.alpha = -1.0; g.beta = 1.0;
        g.A = M + (r0 + h) * ld + r0; g.lda = ld;
        g.B = X + r0 * ldx; g.ldb = ldx;
        g.C = X + (r0 + h) * ldx; g.ldc = ldx;
        g.kmode = K_COL_BEG; g.ktab = ktab; g.koff = r0;
        be.gemm(g);
        trsm_lln(X, ldx, ntc, ktab, r0 + h, rest);
    }
    // X <- L^-T X, rows >= ktab[column tile] only
    void trsm_llt(double *X, int64_t ldx, int ntc, const int32_t *ktab, int64_t r0, int nb) {
        if (nb == 1) {
            GemmDesc g;
            g.al = 1; g.bl = 1; g.mt = 1; g.nt = ntc; g.K = kTile; g.alpha = 1.0; g.beta = 0.0;
            g.A = Dinv + r0 * kTile; g.lda = kTile;
            g.B = X + r0 * ldx; g.ldb = ldx;
            g.C = X + r0 * ldx; g.ldc = ldx;
            g.kmode = K_ROW_MASK; g.ktab = ktab; g.roff = r0;
            be.gemm(g);
            return;
        }
        const int hb = nb / 2, rest = nb - hb;
        const int64_t h = (int64_t)hb * kTile;
        trsm_llt(X, ldx, ntc, ktab, r0 + h, rest);
        GemmDesc g;
        g.al = 1; g.bl = 1; g.mt = hb; g.nt = ntc; g.K = (int64_t)rest * kTile; g.alpha = -1.0; g.beta = 1.0;
        g.A = M + (r0 + h) * ld + r0; g.lda = ld;
        g.B = X + (r0 + h) * ldx; g.ldb = ldx;
        g.C = X + r0 * ldx; g.ldc = ldx;
        g.kmode = K_ROW_MASK; g.ktab = ktab; g.roff = r0;
        be.gemm(g);
        trsm_llt(X, ldx, ntc, ktab, r0, hb);
    }
    void inverse_columns(double *X, int64_t ldx, int ntc, const int32_t *ktab) {
        trsm_lln(X, ldx, ntc, ktab, 0, nblocks());
        trsm_llt(X, ldx, ntc, ktab, 0, nblocks());
    }

    // ---- distributed Cholesky, one rank's share (block-column panels of `pw` tiles, owner = panel mod nranks) ------
    // Every rank holds the whole matrix; a panel is factored by its owner and broadcast, every rank applies it to
    // its OWN later panels only.  `comm` supplies bcast_panel(panel index, row0, rows, col0, cols, root): on the
    // root it ships M[row0.., col0..] (+ the panel's Dinv blocks), elsewhere it receives into the same place.
    // Look-ahead: the owner of panel k+1 updates and factors it before touching its other panels, so the
    // broadcast of k+1 overlaps everybody's remaining updates with panel k.
    // own_cols: first rows (= columns) of the 128-wide column tiles this rank owns, ascending, n_own of them
    // (device-visible for the CUDA backend); used to apply a received panel to ALL own later tiles in one launch.
    template <class Comm>
    void potrf_distributed(Comm &comm, int rank, int nranks, int pw, const int32_t *own_cols = nullptr, int n_own = 0,
                           const int32_t *own_cols_host = nullptr) {
        const int nb = nblocks();
        const int npan = (nb + pw - 1) / pw;
        auto p0 = [&](int p) { return (int64_t)p * pw * kTile; };
        auto pbl = [&](int p) { return std::min(pw, nb - p * pw); };
        auto factor_panel = [&](int p) {
            const int64_t c0 = p0(p);
            const int wb = pbl(p);
            potrf(c0, wb);
            const int below = nb - (p * pw + wb);
            if (below > 0) trsm_rlt(M + (c0 + (int64_t)wb * kTile) * ld, ld, below, c0, wb);
        };
        auto update_panel = [&](int j, int k) {   // panel j -= L[rows >= j, panel k] * L[panel j rows, panel k]'
            const int64_t cj = p0(j), ck = p0(k);
            GemmDesc g;
            g.al = 0; g.bl = 0; g.mt = nb - j * pw; g.nt = pbl(j); g.K = (int64_t)pbl(k) * kTile; g.alpha = -1.0; g.beta = 1.0;
            g.A = M + cj * ld + ck; g.lda = ld;
            g.B = M + cj * ld + ck; g.ldb = ld;
            g.C = M + cj * ld + cj; g.ldc = ld;
            be.gemm(g);
        };
        if (rank == 0) {
            factor_panel(0);
            comm.panel_ready(0);
        }
        for (int k = 0; k < npan; k++) {
            const int64_t c0 = p0(k);
            comm.bcast_panel(k, c0, np - c0, c0, (int64_t)pbl(k) * kTile, k % nranks);
            // look-ahead: next panel first
            if (k + 1 < npan && (k + 1) % nranks == rank) {
                update_panel(k + 1, k);
                // panel k+1 has now seen every panel <= k (earlier ones were applied in earlier iterations)
                factor_panel(k + 1);
                comm.panel_ready(k + 1);
            }
            if (own_cols && own_cols_host) {
                // one trapezoid launch: every own column tile right of panel k+1, all rows on or below its diagonal
                const int64_t first = p0(k + 2);
                int skip = 0;
                while (skip < n_own && own_cols_host[skip] < first) skip++;
                if (skip < n_own) {
                    GemmDesc g;
                    g.al = 0; g.bl = 0; g.mt = nb; g.nt = n_own - skip; g.K = (int64_t)pbl(k) * kTile; g.alpha = -1.0; g.beta = 1.0;
                    g.A = M + c0; g.lda = ld;
                    g.B = M + c0; g.ldb = ld;
                    g.C = M; g.ldc = ld;
                    g.coltab = own_cols + skip; g.ncoltab = n_own - skip; g.row_min = own_cols_host[skip];
                    be.gemm(g);
                }
            } else {
                for (int j = k + 2; j < npan; j++)
                    if (j % nranks == rank) update_panel(j, k);
            }
        }
    }

    // ---- owner-only storage: factor, solve and column-tile inverse with the factor STREAMED panel by panel -----------------------
    // Every rank keeps only its OWN block-column panels of the system (`this->M` = Mo: full height, the own 128-tiles side by side
    // in ascending global order, `this->ld` = 128 * number of own tiles) -- 1/P of the matrix instead of all of it.  A factored
    // panel reaches the other ranks through `comm` and lives there only in a bounded receive window; everything that needs the
    // factor consumes it panel by panel, in the order the panels become available:
    //   forward  (k = 0 .. npan-1, fused with the factorisation):  right-hand-side rows  y_k = L_kk^-1 b_k,
    //            b_>k -= L_>k,k y_k (skinny kernels);   inverse columns  X[k, :] <- L_kk^-1 X[k, :],  X[>k, :] -= L_>k,k X[k, :]
    //   backward (k = npan-1 .. 0, the owners publish their panels a second time):  x_k = L_kk^-T (y_k - L_>k,k' x_>k)
    //            (column-oriented skinny kernels);   X[k, :] -= L_>k,k' X[>k, :],  X[k, :] <- L_kk^-T X[k, :]
    // -- the same operations, in the same order per entry, as potrf_distributed + solve_rows + inverse_columns on a replicated
    // factor (the panel is the first split of their recursions).  A panel is addressed through a VIRTUAL base: L(r, c) =
    // ref.base[r * ref.ld + c] for global r >= c0 and c inside the panel, wherever it lives (own storage or the receive window).
    struct PanelRef { double *base; int64_t ld; };
    // Comm concept:  publish_panel(p)            owner: panel p is final in Mo -> make it available to the other ranks
    //                PanelRef get_panel(k, c0, cols, root, own_ref)   panel k is readable through the returned reference
    //                done_panel(k)               everything that reads panel k has been enqueued
    //                phase_boundary()            between the forward and the backward phase (stage timing)
    template <class Comm>
    void factor_solve_invert_streamed(Comm &comm, int rank, int nranks, int pw, const int32_t *own_cols, int n_own,
                                      const int32_t *own_cols_host, double *R, double *Y, double *X, int64_t ldx, int ntc,
                                      const int32_t *ktab, bool backward) {
        const int nb = nblocks();
        const int npan = (nb + pw - 1) / pw;
        const int64_t pwc = (int64_t)pw * kTile;
        auto p0 = [&](int p) { return (int64_t)p * pwc; };
        auto pbl = [&](int p) { return std::min(pw, nb - p * pw); };
        auto own_ref = [&](int p) {      // own panel p: local panel p / nranks starts at local column (p / nranks) * pwc
            return PanelRef{M + (int64_t)(p / nranks) * pwc - p0(p), ld};
        };
        auto view = [&](const PanelRef &r) { return DenseSchedule<BE>{be, r.base, r.ld, np, Dinv}; };
        auto factor_panel = [&](int p) {
            DenseSchedule<BE> v = view(own_ref(p));
            const int64_t c0 = p0(p);
            const int wb = pbl(p);
            v.potrf(c0, wb);
            const int below = nb - (p * pw + wb);
            if (below > 0) v.trsm_rlt(v.M + (c0 + (int64_t)wb * kTile) * v.ld, v.ld, below, c0, wb);
        };
        auto substitute_forward = [&](int k, const PanelRef &L) {
            DenseSchedule<BE> v = view(L);
            const int64_t c0 = p0(k);
            const int wb = pbl(k), below = nb - (k * pw + wb);
            const int64_t c1 = c0 + (int64_t)wb * kTile;
            if (R)      // the <= 8 right-hand sides (rows of R, np apart; Y = scratch of the same size): one skinny launch per 128-block
                for (int j = k * pw; j < k * pw + wb; j++) be.solve_fwd_block(L.base, L.ld, Dinv, R, Y, np, j);
            if (X && ntc > 0) {
                v.trsm_lln(X, ldx, ntc, ktab, c0, wb);
                if (below > 0) {
                    GemmDesc g;
                    g.al = 0; g.bl = 1; g.mt = below; g.nt = ntc; g.K = (int64_t)wb * kTile; g.alpha = -1.0; g.beta = 1.0;
                    g.A = L.base + c1 * L.ld + c0; g.lda = L.ld; g.B = X + c0 * ldx; g.ldb = ldx; g.C = X + c1 * ldx; g.ldc = ldx;
                    g.kmode = K_COL_BEG; g.ktab = ktab; g.koff = c0;
                    be.gemm(g);
                }
            }
        };
        auto substitute_backward = [&](int k, const PanelRef &L) {
            DenseSchedule<BE> v = view(L);
            const int64_t c0 = p0(k);
            const int wb = pbl(k), below = nb - (k * pw + wb);
            const int64_t c1 = c0 + (int64_t)wb * kTile;
            if (R)
                for (int j = k * pw + wb - 1; j >= k * pw; j--) be.solve_bwd_block_col(L.base, L.ld, Dinv, R, Y, np, j);
            if (X && ntc > 0) {
                if (below > 0) {
                    GemmDesc g;
                    g.al = 1; g.bl = 1; g.mt = wb; g.nt = ntc; g.K = (int64_t)below * kTile; g.alpha = -1.0; g.beta = 1.0;
                    g.A = L.base + c1 * L.ld + c0; g.lda = L.ld; g.B = X + c1 * ldx; g.ldb = ldx; g.C = X + c0 * ldx; g.ldc = ldx;
                    g.kmode = K_ROW_MASK; g.ktab = ktab; g.roff = c0;
                    be.gemm(g);
                }
                v.trsm_llt(X, ldx, ntc, ktab, c0, wb);
            }
        };
        // ---- factorisation with look-ahead, forward substitution on the fly -------------------------------------------------------
        if (rank == 0) {
            factor_panel(0);
            comm.publish_panel(0);
        }
        for (int k = 0; k < npan; k++) {
            const int64_t c0 = p0(k);
            const int root = k % nranks;
            const PanelRef L = comm.get_panel(k, c0, (int64_t)pbl(k) * kTile, root, own_ref(k));
            if (k + 1 < npan && (k + 1) % nranks == rank) {      // look-ahead: the next panel first
                const PanelRef own = own_ref(k + 1);
                const int64_t cj = p0(k + 1);
                GemmDesc g;
                g.al = 0; g.bl = 0; g.mt = nb - (k + 1) * pw; g.nt = pbl(k + 1); g.K = (int64_t)pbl(k) * kTile; g.alpha = -1.0; g.beta = 1.0;
                g.A = L.base + cj * L.ld + c0; g.lda = L.ld;
                g.B = g.A; g.ldb = L.ld;
                g.C = own.base + cj * own.ld + cj; g.ldc = own.ld;
                be.gemm(g);
                factor_panel(k + 1);
                comm.publish_panel(k + 1);
            }
            {   // one trapezoid launch: every own tile right of panel k+1, all rows on or below its diagonal (C compact)
                const int64_t first = p0(k + 2);
                int skip = 0;
                while (skip < n_own && own_cols_host[skip] < first) skip++;
                if (skip < n_own) {
                    GemmDesc g;
                    g.al = 0; g.bl = 0; g.mt = nb; g.nt = n_own - skip; g.K = (int64_t)pbl(k) * kTile; g.alpha = -1.0; g.beta = 1.0;
                    g.A = L.base + c0; g.lda = L.ld;
                    g.B = g.A; g.ldb = L.ld;
                    g.C = M + (int64_t)skip * kTile; g.ldc = ld;
                    g.coltab = own_cols + skip; g.ncoltab = n_own - skip; g.c_local = 1; g.row_min = own_cols_host[skip];
                    be.gemm(g);
                }
            }
            substitute_forward(k, L);
            comm.done_panel(k);
        }
        comm.phase_boundary();
        if (!backward) return;
        // ---- backward substitution: the owners publish their panels once more, last panel first -----------------------------------
        // (a panel is published one step ahead, before the work of the current step is enqueued: its broadcast overlaps that work)
        if ((npan - 1) % nranks == rank) comm.publish_panel(npan - 1);
        for (int k = npan - 1; k >= 0; k--) {
            const int root = k % nranks;
            const PanelRef L = comm.get_panel(k, p0(k), (int64_t)pbl(k) * kTile, root, own_ref(k));
            if (k > 0 && (k - 1) % nranks == rank) comm.publish_panel(k - 1);
            substitute_backward(k, L);
            comm.done_panel(k);
        }
    }
};

}  // namespace jaicov
