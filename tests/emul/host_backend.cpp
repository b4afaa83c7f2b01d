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

// ---- int8-slice ("Ozaki scheme I") emulation of an FP64 GEMM -----------------------------------------------------------
// Numerics study for DESIGN.md section 9 (FP64 products on the INT8 tcgen05 tensor cores): every operand row is scaled by a
// power of two to (-1, 1) and cut into `s` signed digits |q| <= 64 (first digit 6 bits, the others 7 bits); digit products
// are summed EXACTLY in integers per group g = i + j (what an s8 x s8 -> s32 tensor-core accumulator does), groups with
// i + j > s + 1 are dropped, and the group sums are combined in FP64.  This is the arithmetic the planned kernel would do,
// bit for bit; it lives in the test-only host backend and nowhere in the product.
struct OzakiOperand {
    int64_t rows = 0, K = 0;
    int s = 0;
    std::vector<int8_t> q;      // [s][rows][K]
    std::vector<int> e;         // [rows]: |x| < 2^e on the row's valid range
    const int8_t *slice(int j, int64_t r) const { return q.data() + ((size_t)j * rows + r) * K; }
};

// x(r, k) for k in [klo(r), khi(r)); everything outside the valid range becomes zero digits
template <class Get, class Lo, class Hi>
static void ozaki_split(OzakiOperand &o, int64_t rows, int64_t K, int s, Get get, Lo klo, Hi khi) {
    o.rows = rows; o.K = K; o.s = s;
    o.q.assign((size_t)s * rows * K, 0);
    o.e.assign(rows, 0);
    for (int64_t r = 0; r < rows; r++) {
        const int64_t k0 = klo(r), k1 = khi(r);
        double amax = 0.0;
        for (int64_t k = k0; k < k1; k++) amax = std::max(amax, std::fabs(get(r, k)));
        if (!(amax > 0.0) || !std::isfinite(amax)) continue;
        const int e = std::ilogb(amax) + 1;
        o.e[r] = e;
        for (int64_t k = k0; k < k1; k++) {
            double t = std::ldexp(get(r, k), 6 - e);           // |t| < 64
            if (!std::isfinite(t)) t = 0.0;                    // rows nobody reads may hold anything (the device converts NaN to 0 too)
            for (int j = 0; j < s; j++) {
                const double qd = std::nearbyint(t);
                o.q[((size_t)j * rows + r) * K + k] = (int8_t)qd;
                t = (t - qd) * 128.0;                          // exact: |t - qd| <= 0.5
            }
        }
    }
}

struct HostBackend {
    int info = 0;
    long gemm_calls = 0, diag_calls = 0;
    double flops = 0;
    int ozaki = 0;              // > 0: number of int8 digits per operand; every eligible launch goes through gemm_ozaki
    int ozaki_min_tiles = 1;    // launches with fewer output tiles stay on the FP64 path (as the product would do)
    int ozaki_kscale = 1;       // balance the two operands along the contraction index by exact powers of two first (see gemm_ozaki)
    long ozaki_calls = 0;
    long long ozaki_max_abs_sum = 0;   // largest |integer group sum| seen (must stay below 2^31)

    void gemm_ozaki(const GemmDesc &g) {
        const int T = kTile, s = ozaki;
        // column-table launches index op(B) by GLOBAL tile: the operand is the same panel as op(A), all of its g.mt row tiles
        const int64_t Mr = (int64_t)g.mt * T, Nr = g.coltab ? (int64_t)g.mt * T : (int64_t)g.nt * T, K = g.K;
        OzakiOperand A, B;
        auto getA = [&](int64_t m, int64_t k) { return g.al == 0 ? g.A[m * g.lda + k] : g.A[k * g.lda + m]; };
        auto getB = [&](int64_t n, int64_t k) { return g.bl == 0 ? g.B[n * g.ldb + k] : g.B[k * g.ldb + n]; };
        auto loA = [&](int64_t m) { return g.kmode == K_MAX_IJ ? (m / T) * T : (int64_t)0; };
        const int64_t rmin = g.coltab ? (g.row_min / T) * T : 0;      // rows above it belong to no wanted tile: never read (csrc/ozaki.cu: row_min)
        auto hiA = [&](int64_t m) { return m < rmin ? (int64_t)0 : (g.kmode == K_A_LOWER ? std::min<int64_t>(K, (m / T + 1) * T) : K); };
        auto loB = [&](int64_t n) { return (g.kmode == K_MAX_IJ || g.kmode == K_B_LOWER) ? (n / T) * T : (int64_t)0; };
        auto hiB = [&](int64_t n) { return n < rmin ? (int64_t)0 : K; };
        // Contraction-index balancing: op(A)[., k] * 2^f_k and op(B)[., k] * 2^-f_k leave every product unchanged (exact powers of
        // two) but even out the magnitudes inside the operand rows, which is what the per-row digit grid resolves.  f_k = floor of
        // half the exponent gap of the two column maxima.  Needed where the operands span many orders of magnitude inside a row
        // (the structured route's Q'Y' and Y (Q'Y'), tests/ozaki_study.py --structured); nothing to do when A and B are the same array.
        std::vector<int> fk(K, 0);
        const bool same = (const void *)g.A == (const void *)g.B && g.lda == g.ldb && g.al == g.bl && g.kmode != K_A_LOWER && g.kmode != K_B_LOWER && Mr == Nr;
        if (ozaki_kscale && !same) {
            std::vector<double> ca(K, 0.0), cb(K, 0.0);
            for (int64_t m = 0; m < Mr; m++)
                for (int64_t k = loA(m); k < hiA(m); k++) ca[k] = std::max(ca[k], std::fabs(getA(m, k)));
            for (int64_t n = 0; n < Nr; n++)
                for (int64_t k = loB(n); k < hiB(n); k++) cb[k] = std::max(cb[k], std::fabs(getB(n, k)));
            for (int64_t k = 0; k < K; k++)
                if (ca[k] >= 2.3e-308 && cb[k] >= 2.3e-308 && std::isfinite(ca[k]) && std::isfinite(cb[k]))
                    fk[k] = std::max(-1000, std::min(1000, (std::ilogb(cb[k]) - std::ilogb(ca[k])) >> 1));   // arithmetic shift: floor((eb - ea) / 2)
        }
        ozaki_split(A, Mr, K, s, [&](int64_t m, int64_t k) { return std::ldexp(getA(m, k), fk[k]); }, loA, hiA);
        ozaki_split(B, Nr, K, s, [&](int64_t n, int64_t k) { return std::ldexp(getB(n, k), -fk[k]); }, loB, hiB);
        std::vector<long long> S(2 * s + 1);
        std::vector<double> cbuf((size_t)T * T);
        for (int it = 0; it < g.mt; it++)
            for (int jl = 0; jl < g.nt; jl++) {
                int jt = jl;
                if (g.coltab) {
                    jt = g.coltab[jl] / T;
                    if (!g.coltab_full && it < jt) continue;
                }
                if (g.tri_out && it < jt) continue;
                int64_t kbeg = 0, kend = K;
                if (g.kmode == K_B_LOWER) kbeg = (int64_t)jt * T;
                else if (g.kmode == K_A_LOWER) kend = std::min<int64_t>(K, (int64_t)(it + 1) * T);
                else if (g.kmode == K_MAX_IJ) kbeg = (int64_t)std::max(it, jt) * T;
                else if (g.kmode == K_COL_BEG) {
                    kbeg = std::max<int64_t>(0, (int64_t)g.ktab[jt] - g.koff);
                    if (kbeg >= K) continue;
                } else if (g.kmode == K_ROW_MASK) {
                    if (g.roff + (int64_t)it * T < (int64_t)g.ktab[jt]) continue;
                }
                flops += 2.0 * T * T * (double)(kend - kbeg);
                for (int i = 0; i < T; i++)
                    for (int j = 0; j < T; j++) {
                        const int64_t m = (int64_t)it * T + i, n = (int64_t)jt * T + j;
                        for (auto &v : S) v = 0;
                        for (int a = 0; a < s; a++) {
                            const int8_t *qa = A.slice(a, m);
                            for (int b = 0; a + b + 2 <= s + 1; b++) {
                                const int8_t *qb = B.slice(b, n);
                                int32_t acc = 0;                               // one digit pair never overflows: K * 4096 < 2^31
                                for (int64_t k = kbeg; k < kend; k++) acc += (int32_t)qa[k] * (int32_t)qb[k];
                                S[a + b + 2] += acc;
                            }
                        }
                        double r = 0.0;
                        for (int gq = s + 1; gq >= 2; gq--) {                  // Horner, smallest contributions first:
                            ozaki_max_abs_sum = std::max(ozaki_max_abs_sum, std::llabs(S[gq]));
                            r = r * 0.0078125 + (double)S[gq];                 // r / 128 + S_g (the kernel's epilogue)
                        }
                        r = std::ldexp(r, A.e[m] + B.e[n] - 12);               // digit weights 2^-(7 g - 2), g = 2 last
                        const int64_t ccol = (g.coltab && g.c_local) ? (int64_t)jl * T + j : n;
                        cbuf[(size_t)i * T + j] = g.alpha * r + (g.beta == 0.0 ? 0.0 : g.beta * g.C[m * g.ldc + ccol]);
                    }
                // the tile is stored after all of it has been computed (the digit planes are copies, so in-place launches are safe anyway)
                for (int i = 0; i < T; i++)
                    for (int j = 0; j < T; j++) {
                        const int64_t m = (int64_t)it * T + i, n = (int64_t)jt * T + j;
                        g.C[m * g.ldc + ((g.coltab && g.c_local) ? (int64_t)jl * T + j : n)] = cbuf[(size_t)i * T + j];
                    }
            }
    }

    void gemm(const GemmDesc &g) {
        gemm_calls++;
        // what csrc/ozaki.cu: launch_gemm_ozaki takes: everything without a column table, and the trapezoid update of the
        // distributed Cholesky (column table, both operands the same panel of the matrix, C addressed by global tiles)
        const bool trapezoid = g.coltab && !g.coltab_full && (const void *)g.A == (const void *)g.B && g.lda == g.ldb && g.al == g.bl && g.kmode == K_FULL;
        if (ozaki > 0 && (!g.coltab || trapezoid)) {
            const long tiles = g.tri_out ? (long)g.mt * (g.mt + 1) / 2 : (long)g.mt * g.nt;
            if (tiles >= ozaki_min_tiles) { ozaki_calls++; gemm_ozaki(g); return; }
        }
        const int T = kTile;
        std::vector<double> acc((size_t)T * T);
        for (int it = 0; it < g.mt; it++)
            for (int jl = 0; jl < g.nt; jl++) {
                int jt = jl;
                if (g.coltab) {
                    jt = g.coltab[jl] / T;
                    if (it < jt) continue;
                }
                if (g.tri_out && it < jt) continue;
                int64_t kbeg = 0, kend = g.K;
                if (g.kmode == K_B_LOWER) kbeg = (int64_t)jt * T;
                else if (g.kmode == K_A_LOWER) kend = std::min<int64_t>(g.K, (int64_t)(it + 1) * T);
                else if (g.kmode == K_MAX_IJ) kbeg = (int64_t)std::max(it, jt) * T;
                else if (g.kmode == K_COL_BEG) {
                    kbeg = std::max<int64_t>(0, (int64_t)g.ktab[jt] - g.koff);
                    if (kbeg >= g.K) continue;
                } else if (g.kmode == K_ROW_MASK) {
                    if (g.roff + (int64_t)it * T < (int64_t)g.ktab[jt]) continue;
                }
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
                        double *c = g.C + ((int64_t)it * T + i) * g.ldc + ((g.coltab && g.c_local) ? (int64_t)jl : (int64_t)jt) * T + j;
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

    // skinny solves of 8 right-hand sides stored as rows R[8][np] (csrc/dense_kernels.cu: k_solve_fwd_step, k_solve_bwd_col_*)
    void solve_fwd_block(const double *L, int64_t ld, const double *dinv, double *R, double *Y, int64_t np, int j) {
        const int T = kTile;
        const int64_t J = (int64_t)j * T;
        for (int q = 0; q < 8; q++) {
            double y[kTile];
            for (int i = 0; i < T; i++) {
                double s = 0.0;
                for (int k = 0; k < T; k++) s += dinv[(J + i) * T + k] * R[(int64_t)q * np + J + k];
                y[i] = s;
            }
            for (int i = 0; i < T; i++) Y[(int64_t)q * np + J + i] = y[i];
            for (int64_t r = J + T; r < np; r++) {
                double s = 0.0;
                for (int k = 0; k < T; k++) s += L[r * ld + J + k] * y[k];
                R[(int64_t)q * np + r] -= s;
            }
        }
    }
    void solve_bwd_block_col(const double *L, int64_t ld, const double *dinv, double *R, const double *Y, int64_t np, int j) {
        const int T = kTile;
        const int64_t J = (int64_t)j * T;
        for (int q = 0; q < 8; q++) {
            double t[kTile];
            for (int c = 0; c < T; c++) {
                double s = 0.0;
                for (int64_t r = J + T; r < np; r++) s += L[r * ld + J + c] * R[(int64_t)q * np + r];
                t[c] = Y[(int64_t)q * np + J + c] - s;
            }
            for (int i = 0; i < T; i++) {
                double s = 0.0;
                for (int k = 0; k < T; k++) s += dinv[(J + k) * T + i] * t[k];      // Dinv' t
                R[(int64_t)q * np + J + i] = s;
            }
        }
    }

    void copy2d(double *dst, int64_t ldd, const double *src, int64_t lds, int64_t rows, int64_t cols) {
        for (int64_t r = 0; r < rows; r++) std::memcpy(dst + r * ldd, src + r * lds, sizeof(double) * cols);
    }
};

#include <condition_variable>
#include <mutex>
#include <thread>

struct Barrier {
    std::mutex m; std::condition_variable cv; int n, count = 0, gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++count == n) { count = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

// virtual ranks in threads: a broadcast copies the root replica's panel (and its Dinv blocks) into every replica
struct HostComm {
    int rank, nranks;
    std::vector<double *> *Ms, *Dinvs;
    int64_t np;
    Barrier *bar;
    void bcast_panel(int, int64_t row0, int64_t rows, int64_t col0, int64_t cols, int root) {
        bar->wait();
        if (rank != root) {
            for (int64_t r = 0; r < rows; r++)
                std::memcpy((*Ms)[rank] + (row0 + r) * np + col0, (*Ms)[root] + (row0 + r) * np + col0, sizeof(double) * cols);
            std::memcpy((*Dinvs)[rank] + col0 * kTile, (*Dinvs)[root] + col0 * kTile, sizeof(double) * cols * kTile);
        }
        bar->wait();
    }
    void panel_ready(int) {}
};

// Comm of DenseSchedule::factor_solve_invert_streamed for virtual ranks in threads: the root's panel (own storage, compact columns)
// is copied into the receiver's two-slot window; nobody but the owner ever holds a panel for longer than it is in the window
struct HostStreamComm {
    int rank, nranks, pw;
    std::vector<double *> *Mos;          // own-panel storage of every rank
    std::vector<int64_t> *ldos;
    std::vector<double *> *Dinvs;
    std::vector<double> window[2];
    int64_t np;
    Barrier *bar;
    long received = 0;
    void publish_panel(int) {}
    DenseSchedule<HostBackend>::PanelRef get_panel(int k, int64_t c0, int64_t cols, int root, DenseSchedule<HostBackend>::PanelRef own) {
        bar->wait();                     // the root has finished writing the panel
        DenseSchedule<HostBackend>::PanelRef ref = own;
        if (rank != root) {
            std::vector<double> &w = window[k & 1];
            w.assign((size_t)(np - c0) * cols, std::numeric_limits<double>::quiet_NaN());
            const double *src = (*Mos)[root] + (int64_t)(k / nranks) * pw * kTile;      // root's local panel, row 0
            const int64_t lds = (*ldos)[root];
            for (int64_t r = c0; r < np; r++) std::memcpy(w.data() + (r - c0) * cols, src + r * lds, sizeof(double) * cols);
            std::memcpy((*Dinvs)[rank] + c0 * kTile, (*Dinvs)[root] + c0 * kTile, sizeof(double) * cols * kTile);
            ref.base = w.data() - c0 * cols - c0;
            ref.ld = cols;
            received++;
        }
        bar->wait();                     // everybody has its copy: the root may go on (it never overwrites a published panel)
        return ref;
    }
    void done_panel(int) {}
    void phase_boundary() {}
};

extern "C" {

// Owner-only storage (DenseSchedule::factor_solve_invert_streamed) with `nranks` virtual ranks: every rank gets ONLY its own
// block-column panels of M (everything else it holds is NaN), the right-hand sides R (nrhs rows, replicated) and the identity
// columns of its inverse tiles (tile c belongs to rank c % nranks).  Out: R <- solutions, Q (np x np, lower part) <- the inverse
// gathered from the ranks' tiles, maxdiff <- largest difference between the ranks' copies of the solutions.
int emul_streamed(int64_t np, const double *M, int nranks, int pw, int nrhs, double *R, double *Q, double *maxdiff, int ozaki, long *received) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const int nb = (int)(np / kTile);
    const int64_t pwc = (int64_t)pw * kTile;
    std::vector<std::vector<double>> Mo(nranks), Dr(nranks, std::vector<double>((size_t)np * kTile, nan)), Rr(nranks), Xr(nranks);
    std::vector<std::vector<int32_t>> own(nranks), ktab(nranks);
    std::vector<double *> Mos, Ds;
    std::vector<int64_t> ldos;
    for (int r = 0; r < nranks; r++) {
        for (int c = 0; c < nb; c++) {
            if ((c / pw) % nranks == r) own[r].push_back(c * kTile);
            if (c % nranks == r) ktab[r].push_back(c * kTile);
        }
        const int64_t ldo = std::max<int64_t>(1, (int64_t)own[r].size()) * kTile;
        Mo[r].assign((size_t)np * ldo, nan);
        for (size_t lt = 0; lt < own[r].size(); lt++)
            for (int64_t row = own[r][lt]; row < np; row++)        // lower part of the own tiles only
                for (int j = 0; j < kTile; j++)
                    if (own[r][lt] + j <= row) Mo[r][(size_t)row * ldo + lt * kTile + j] = M[row * np + own[r][lt] + j];
        Rr[r].assign((size_t)16 * np, 0.0);       // rows 0..7 right-hand sides, rows 8..15 the scratch Y
        for (int i = 0; i < nrhs; i++) std::memcpy(Rr[r].data() + (size_t)i * np, R + (size_t)i * np, sizeof(double) * np);
        const int64_t ldx = std::max<int64_t>(1, (int64_t)ktab[r].size()) * kTile;
        Xr[r].assign((size_t)np * ldx, 0.0);
        for (size_t jl = 0; jl < ktab[r].size(); jl++)
            for (int i = 0; i < kTile; i++) Xr[r][(size_t)(ktab[r][jl] + i) * ldx + jl * kTile + i] = 1.0;
        Mos.push_back(Mo[r].data()); Ds.push_back(Dr[r].data()); ldos.push_back(ldo);
    }
    Barrier bar(nranks);
    std::vector<int> infos(nranks, 0);
    std::vector<long> recv(nranks, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; r++)
        th.emplace_back([&, r] {
            HostBackend be;
            be.ozaki = ozaki;
            HostStreamComm comm{r, nranks, pw, &Mos, &ldos, &Ds, {}, np, &bar};
            DenseSchedule<HostBackend> ds{be, Mos[r], ldos[r], np, Ds[r]};
            ds.factor_solve_invert_streamed(comm, r, nranks, pw, own[r].data(), (int)own[r].size(), own[r].data(), Rr[r].data(), Rr[r].data() + 8 * np,
                                            Xr[r].data(), (int64_t)std::max<size_t>(1, ktab[r].size()) * kTile, (int)ktab[r].size(), ktab[r].data(), true);
            infos[r] = be.info;
            recv[r] = comm.received;
        });
    for (auto &t : th) t.join();
    (void)pwc;
    double md = 0.0;
    for (int r = 1; r < nranks; r++)
        for (int i = 0; i < nrhs; i++)
            for (int64_t e = 0; e < np; e++) md = std::max(md, std::fabs(Rr[r][(size_t)i * np + e] - Rr[0][(size_t)i * np + e]));
    *maxdiff = md;
    for (int i = 0; i < nrhs; i++) std::memcpy(R + (size_t)i * np, Rr[0].data() + (size_t)i * np, sizeof(double) * np);
    for (int r = 0; r < nranks; r++) {
        const int64_t ldx = std::max<int64_t>(1, (int64_t)ktab[r].size()) * kTile;
        for (size_t jl = 0; jl < ktab[r].size(); jl++)
            for (int64_t row = ktab[r][jl]; row < np; row++)
                for (int i = 0; i < kTile; i++) {
                    const int64_t col = ktab[r][jl] + i;
                    if (row >= col) Q[row * np + col] = Xr[r][(size_t)row * ldx + jl * kTile + i];
                }
    }
    if (received) *received = recv[nranks > 1 ? 1 : 0];
    int info = 0;
    for (int v : infos) if (v) info = v;
    return info;
}

// tile order of the lower-triangular GEMM launches (dense_driver.hpp: tri_tile_decode), for the bijection test
void emul_rect_tile_decode(int64_t l, int mt, int nt, int band, int *it, int *jt) { jaicov::rect_tile_decode(l, mt, nt, band, *it, *jt); }
void emul_tri_tile_decode(int64_t l, int mt, int band, int *it, int *jt) { jaicov::tri_tile_decode(l, mt, band, *it, *jt); }

// distributed Cholesky with `nranks` virtual ranks (threads), panel width pw tiles; on exit every replica must hold
// the same factor; replica 0 is returned in M (lower), followed by the column-panel inverse of the column tiles
// owned by each rank (tile c belongs to rank (c / pw) % nranks), gathered into Q (np x np, lower part valid).
int emul_distributed_ex(int64_t np, double *M, int nranks, int pw, double *Q, double *maxdiff, int merged, int ozaki);
int emul_distributed(int64_t np, double *M, int nranks, int pw, double *Q, double *maxdiff, int merged) {
    return emul_distributed_ex(np, M, nranks, pw, Q, maxdiff, merged, 0);
}
// ozaki > 0: every eligible launch of both stages (panel factor, trapezoid updates, column-tile inverse) through the digit emulation
int emul_distributed_ex(int64_t np, double *M, int nranks, int pw, double *Q, double *maxdiff, int merged, int ozaki) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const int nb = (int)(np / kTile);
    std::vector<std::vector<double>> Mr(nranks, std::vector<double>(M, M + np * np)), Dr(nranks, std::vector<double>((size_t)np * kTile, nan));
    std::vector<double *> Ms, Ds;
    for (int r = 0; r < nranks; r++) { Ms.push_back(Mr[r].data()); Ds.push_back(Dr[r].data()); }
    Barrier bar(nranks);
    std::vector<int> infos(nranks, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; r++)
        th.emplace_back([&, r] {
            HostBackend be;
            be.ozaki = ozaki;
            HostComm comm{r, nranks, &Ms, &Ds, np, &bar};
            DenseSchedule<HostBackend> ds{be, Ms[r], np, np, Ds[r]};
            std::vector<int32_t> own;
            for (int c = 0; c < (int)(np / kTile); c++) if ((c / pw) % nranks == r) own.push_back(c * kTile);
            if (merged) ds.potrf_distributed(comm, r, nranks, pw, own.data(), (int)own.size(), own.data());
            else ds.potrf_distributed(comm, r, nranks, pw);
            infos[r] = be.info;
        });
    for (auto &t : th) t.join();
    double md = 0.0;
    for (int r = 1; r < nranks; r++)
        for (int64_t i = 0; i < np; i++)
            for (int64_t j = 0; j <= i; j++) md = std::max(md, std::fabs(Mr[r][i * np + j] - Mr[0][i * np + j]));
    *maxdiff = md;
    for (int64_t i = 0; i < np; i++) for (int64_t j = 0; j <= i; j++) M[i * np + j] = Mr[0][i * np + j];
    // column-panel inverse per rank
    for (int r = 0; r < nranks; r++) {
        std::vector<int32_t> ktab;
        for (int c = 0; c < nb; c++) if ((c / pw) % nranks == r) ktab.push_back(c * kTile);
        const int ntc = (int)ktab.size();
        if (!ntc) continue;
        const int64_t ldx = (int64_t)ntc * kTile;
        std::vector<double> X((size_t)np * ldx, 0.0);
        for (int jl = 0; jl < ntc; jl++)
            for (int i = 0; i < kTile; i++) X[(size_t)(ktab[jl] + i) * ldx + jl * kTile + i] = 1.0;
        HostBackend be;
        be.ozaki = ozaki;
        DenseSchedule<HostBackend> ds{be, Ms[r], np, np, Ds[r]};
        ds.inverse_columns(X.data(), ldx, ntc, ktab.data());
        for (int jl = 0; jl < ntc; jl++)
            for (int64_t row = ktab[jl]; row < np; row++)
                for (int i = 0; i < kTile; i++) {
                    const int64_t col = ktab[jl] + i;
                    if (row >= col) Q[row * np + col] = X[(size_t)row * ldx + jl * kTile + i];
                }
    }
    int info = 0;
    for (int v : infos) if (v) info = v;
    return info;
}

// one tile-grid product through the host backend (ozaki = 0: FP64 loops; > 0: int8-digit emulation): the host twin of
// jaicov_gemm_tiles, used to validate the expectations of tools/ozaki_gpu_check.py without a GPU
int emul_gemm_tiles(int al, int bl, int mt, int nt, int64_t K, double alpha, double beta, const double *A, int64_t lda, const double *B,
                    int64_t ldb, double *C, int64_t ldc, int tri_out, int kmode, int ozaki) {
    HostBackend be;
    be.ozaki = ozaki < 0 ? -ozaki : ozaki;        // negative: without the contraction-index balancing
    be.ozaki_kscale = ozaki < 0 ? 0 : 1;
    GemmDesc g;
    g.al = al; g.bl = bl; g.mt = mt; g.nt = nt; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.tri_out = tri_out; g.kmode = kmode;
    be.gemm(g);
    return (int)be.ozaki_calls;
}

// M: np x np row-major; on entry the LOWER triangle holds an SPD matrix (upper is poisoned here with NaN);
// R: mt*128 x np right-hand-side rows (solved in place); on exit M lower = inverse.
int emul_spd_solve_invert_ex(int64_t np, double *M, int mt, double *R, int invert, double *stats, int ozaki, int ozaki_min_tiles);
int emul_spd_solve_invert(int64_t np, double *M, int mt, double *R, int invert, double *stats) {
    return emul_spd_solve_invert_ex(np, M, mt, R, invert, stats, 0, 1);
}

// the same with every GEMM launch of >= ozaki_min_tiles output tiles computed by the int8-slice emulation (ozaki digits);
// stats (5 doubles): gemm launches, diagonal blocks, flop, launches through the emulation, largest integer group sum
int emul_spd_solve_invert_ex(int64_t np, double *M, int mt, double *R, int invert, double *stats, int ozaki, int ozaki_min_tiles) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t r = 0; r < np; r++)
        for (int64_t c = r + 1; c < np; c++)
            if ((r / kTile) != (c / kTile)) M[r * np + c] = nan;   // strictly-upper off-diagonal tiles are never read
    std::vector<double> Dinv((size_t)np * kTile, nan), W;
    HostBackend be;
    be.ozaki = ozaki < 0 ? -ozaki : ozaki;        // a negative digit count switches the contraction-index balancing off (studies)
    be.ozaki_kscale = ozaki < 0 ? 0 : 1;
    be.ozaki_min_tiles = ozaki_min_tiles;
    DenseSchedule<HostBackend> ds{be, M, np, np, Dinv.data()};
    ds.potrf();
    if (mt > 0) ds.solve_rows(R, np, mt);
    if (invert) {
        W.assign((size_t)np * np, nan);
        ds.invert_from_factor(W.data());
    }
    if (stats) {
        stats[0] = (double)be.gemm_calls; stats[1] = (double)be.diag_calls; stats[2] = be.flops;
        if (ozaki != 0) { stats[3] = (double)be.ozaki_calls; stats[4] = (double)be.ozaki_max_abs_sum; }
    }
    return be.info;
}
}
