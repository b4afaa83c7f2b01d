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

extern "C" {

// tile order of the lower-triangular GEMM launches (dense_driver.hpp: tri_tile_decode), for the bijection test
void emul_tri_tile_decode(int64_t l, int mt, int band, int *it, int *jt) { jaicov::tri_tile_decode(l, mt, band, *it, *jt); }

// distributed Cholesky with `nranks` virtual ranks (threads), panel width pw tiles; on exit every replica must hold
// the same factor; replica 0 is returned in M (lower), followed by the column-panel inverse of the column tiles
// owned by each rank (tile c belongs to rank (c / pw) % nranks), gathered into Q (np x np, lower part valid).
int emul_distributed(int64_t np, double *M, int nranks, int pw, double *Q, double *maxdiff, int merged) {
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
