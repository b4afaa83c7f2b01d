// api.cu -- C ABI (include/jaicov_b200.h) and the device-resident adjustment loop.
//
// Restates the control flow of BundleAdjustment.estimateModel() (BundleAdjustment.java:203-387) around the CUDA
// stages; the Java host keeps the integer bookkeeping (prepareUnknownParameters :667-782, detectRankDefect
// :836-1042) and hands over flat arrays.  No CPU compute path exists here: without a usable sm_100 device every
// computing entry point fails with JAICOV_NOT_INITIALISED.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "common.h"
#include "dense_driver.hpp"
#include "dist.h"
#include "model.cuh"

namespace jaicov {

std::atomic<long long> g_launch_count{0};

// dense_kernels.cu
void launch_gemm(const GemmDesc &g, cudaStream_t s);
size_t ozaki_release_scratch();   // ozaki.cu
int ozaki_set_digits(int digits);
void launch_potrf_diag(double *A, int64_t ld, double *dinv, int row0, int *info, cudaStream_t s);
void launch_copy2d(double *dst, int64_t ldd, const double *src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t s);
void launch_solve_rows8(const double *L, int64_t ld, const double *dinv, double *R, double *Y, int64_t np, cudaStream_t s);
void launch_solve_fwd_block(const double *L, int64_t ld, const double *dinv, double *R, double *Y, int64_t np, int j, cudaStream_t s);
void launch_solve_bwd_block_col(const double *L, int64_t ld, const double *dinv, double *R, const double *Y, int64_t np, int j, double *partial,
                                cudaStream_t s);
// stage_kernels.cu
void launch_precond_diag(const SysView &M, int u, int64_t np, double *V, cudaStream_t s);
void launch_datum_rows(const double *xyz, const int32_t *pt_col, const int32_t *datum_pts, int nDatum, int free_mask, int d,
                       int64_t np, double *Bt, cudaStream_t s);
void launch_scale_system(double *M, int64_t ld, int u, const double *V, const double *Bt, int d, int64_t np, int64_t row0, cudaStream_t s);
void launch_scale_system_own(double *Mo, int64_t ldo, const int32_t *own_cols, int n_own, int u, const double *V, const double *Bt, int d,
                             int64_t np, cudaStream_t s);
void launch_build_rhs(double *Rt, double *Btv, int64_t np, int u, const double *V, const double *rhs, const double *Bt, int d,
                      int simulation, cudaStream_t s);
void launch_datum_solve(const double *Xt, const double *Btv, int d, int64_t np, int u, const double *V, double *dxref, double *H,
                        double *Tq, double *small, cudaStream_t s);
void launch_qxx_epilogue(double *M, int64_t ld, int u, const double *V, const double *H, const double *G, int d, int64_t np,
                         cudaStream_t s);
void launch_update(double *val, const int32_t *col, int64_t n, const double *dxref, int apply, unsigned long long *out,
                   cudaStream_t s);
void launch_pack_columns(const double *lower, int64_t ld, const double *border, int64_t np, const double *q11, int d, int64_t c0,
                         int64_t c1, double *out, cudaStream_t s);
void launch_get_block(const double *lower, int64_t ld, const double *border, int64_t np, const double *q11, int d, int r0, int r1,
                      int c0, int c1, double *out, cudaStream_t s);
void launch_group_w(int r, const double *const *tptr, const double *obs, double *w, cudaStream_t s);
void launch_group_stack(int r, const int32_t *col, const double *var, const double *Pw, int64_t ldp, double sigma2, const double *w,
                        int d, const SysView &M, double *rhs, cudaStream_t s);
void launch_group_omega(int r, const int32_t *col, const double *var, const double *Pw, int64_t ldp, double sigma2, const double *w,
                        const double *dxref, double *out, cudaStream_t s);
void launch_unpack_scaled(const double *ap, int r, double scale, double *M, int64_t ld, int64_t np, cudaStream_t s);
void launch_symmetrize(double *M, int64_t ld, int r, cudaStream_t s);
void launch_damp_diag(const SysView &M, int u, double lambda, cudaStream_t s);
void launch_scale_vector(double *x, int64_t n, double alpha, cudaStream_t s);
void launch_identity_columns(double *X, int64_t ldx, int64_t np, const int32_t *ktab, int ntc, cudaStream_t s);
void launch_qxx_epilogue_cols(double *X, int64_t ldx, int ntc, const int32_t *ktab, int u, const double *V, const double *H,
                              const double *G, int d, int64_t np, cudaStream_t s);
void launch_get_submatrix(const double *lower, int64_t ld, const double *X, int64_t ldx, const int32_t *col_local, const double *border,
                          int64_t np, const double *q11, int d, int rank, const int32_t *row_idx, int n_rows, const int32_t *col_idx,
                          int n_cols, double scale, double *out, cudaStream_t s);
void launch_get_block_dist(const double *X, int64_t ldx, const int32_t *col_local, const double *border, int64_t np,
                           const double *q11, int d, int rank, int r0, int r1, int c0, int c1, double *out, cudaStream_t s);

void launch_propagate(const double *lower, int64_t ld, const double *X, int64_t ldx, const int32_t *col_local, const double *border,
                      int64_t np, const double *q11, int d, int rank, int nT, const double *Jv, const int32_t *Jc, double sigma2,
                      double *out, cudaStream_t s);

void launch_dlt_batch(int n_img, const int64_t *pt_ptr, const double *xy, const double *XYZ, const double *io, int nR,
                      const int32_t *restr, int max_iterations, double *out, int32_t *status, int32_t *passes, cudaStream_t s);

void launch_group_product(int r, const int32_t *col, const double *var, const double *Pw, int64_t ldp, double sigma2, const double *w,
                          int nv, int64_t n, const double *X, double *Y, double *rhs, double *wpw, cudaStream_t s);
void launch_normal_product_points(const DevProblem &P, int nv, int64_t n, const double *X, double *Y, double *rhs, double *wpw,
                                  cudaStream_t s);
void launch_normal_product_rest(const DevProblem &P, int nv, int64_t n, const double *X, double *Y, double *rhs, double *wpw,
                                const double *Bt, int64_t ldb, cudaStream_t s);

// dense_sigma.cu: image points of one image with a fully populated dispersion matrix (extension, north_star (2))
void launch_zero_weights(double *rw, int64_t obs_begin, int64_t obs_end, cudaStream_t s);
void launch_image_rows(const DevProblem &P, int img, double *Ac, int64_t rp, cudaStream_t s);
void launch_dense_image_assemble(const DevProblem &P, int img, int64_t m, const double *Pw, int64_t ldp, int64_t rp, double *Ac, double *T,
                                 double *G, double *M, double *rhs, cudaStream_t s);
void launch_dense_image_omega(const DevProblem &P, int img, const double *Pw, int64_t ldp, int64_t rp, double *Ac, const double *x, double *v,
                              double *t, double *out, cudaStream_t s);
void launch_dense_image_product(const DevProblem &P, int img, const double *Pw, int64_t ldp, int64_t rp, const double *Ac, const double *x,
                                int mode, double *v, double *t, double *y, double *wpw, cudaStream_t s);

void launch_img_of_obs(const int64_t *pt_ptr, int nImg, int32_t *img_of_obs, cudaStream_t s);
size_t csc_temp_bytes(int64_t n);
void launch_build_csc(const int32_t *obj, int64_t obs0, int64_t n, int nPt, int32_t *keys_out, int64_t *iota, void *temp,
                      size_t temp_bytes, int64_t *pt_obs_ptr, int64_t *pt_obs, int32_t minmax_host[2], cudaStream_t s);

struct CudaBackend {
    cudaStream_t stream;
    int *info;
    double *solve_partial = nullptr;   // scratch of the column-oriented skinny backward solve (owner-only storage)
    void solve_fwd_block(const double *L, int64_t ld, const double *dinv, double *R, double *Y, int64_t np, int j) {
        launch_solve_fwd_block(L, ld, dinv, R, Y, np, j, stream);
    }
    void solve_bwd_block_col(const double *L, int64_t ld, const double *dinv, double *R, const double *Y, int64_t np, int j) {
        launch_solve_bwd_block_col(L, ld, dinv, R, Y, np, j, solve_partial, stream);
    }
    void gemm(const GemmDesc &g) { launch_gemm(g, stream); }
    void potrf_diag(double *a, int64_t ld, double *dinv, int row0) { launch_potrf_diag(a, ld, dinv, row0, info, stream); }
    void copy2d(double *dst, int64_t ldd, const double *src, int64_t lds, int64_t rows, int64_t cols) {
        launch_copy2d(dst, ldd, src, lds, rows, cols, stream);
    }
};

// Process-wide cache of large device allocations: cudaMalloc/cudaFree of tens of GB cost hundreds of milliseconds and
// every estimateModel() call needs the same buffers again, so big blocks are parked here instead of being freed.
struct DevCache {
    static constexpr size_t kMinBytes = 1;     // every buffer: cudaFree / cudaMalloc cost milliseconds each once tens of GB are mapped
    struct Entry { int device; size_t bytes; void *p; };
    std::vector<Entry> free_list;
    std::mutex mu;                             // handles may live on different host threads
    static int current_device() { int d = 0; cudaGetDevice(&d); return d; }
    void *take(size_t bytes) {
        const int dev = current_device();
        std::lock_guard<std::mutex> lock(mu);
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].bytes == bytes && free_list[i].device == dev) {
                void *p = free_list[i].p;
                free_list.erase(free_list.begin() + i);
                return p;
            }
        return nullptr;
    }
    void give(size_t bytes, void *p, int dev) {
        // bounded by count only (a handle holds ~100 buffers); memory pressure is handled where it shows: a failed
        // cudaMalloc purges the cache and retries (DevBuf::alloc).  `dev` is the device the buffer was allocated on (the
        // caller's current device may be another one by the time a handle is destroyed).
        std::lock_guard<std::mutex> lock(mu);
        if (free_list.size() >= 1024) {   // drop the oldest entry
            cudaFree(free_list.front().p);
            free_list.erase(free_list.begin());
        }
        free_list.push_back(Entry{dev, bytes, p});
    }
    size_t purge() {
        std::lock_guard<std::mutex> lock(mu);
        size_t bytes = 0;
        for (auto &e : free_list) { cudaFree(e.p); bytes += e.bytes; }
        free_list.clear();
        return bytes;
    }
};
static DevCache g_cache;

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int dev = -1;                              // device the buffer lives on
    void alloc(size_t count) {
        release();
        n = count;
        if (!count) return;
        dev = DevCache::current_device();
        const size_t bytes = count * sizeof(T);
        if (bytes >= DevCache::kMinBytes) {
            p = static_cast<T *>(g_cache.take(bytes));
            if (p) return;
        }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaErrorMemoryAllocation) {   // make room and retry once
            cudaGetLastError();
            g_cache.purge();
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) { p = nullptr; n = 0; throw CudaError{e, "cudaMalloc", __FILE__, __LINE__}; }
    }
    void upload(const std::vector<T> &h) {
        alloc(h.size());
        if (!h.empty()) JCHECK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    void upload(const T *src, size_t count) {      // straight from the caller's buffer (no host copy in between)
        alloc(count);
        if (count) JCHECK(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    }
    void release() {
        if (p) {
            if (n * sizeof(T) >= DevCache::kMinBytes) g_cache.give(n * sizeof(T), p, dev);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct Group {
    int r = 0;
    std::vector<int32_t> kind, index, comp;
    std::vector<double> obs, var, sigma;   // sigma: packed upper dispersion (may be empty)
    DevBuf<const double *> tptr;
    DevBuf<int32_t> col;
    DevBuf<double> d_obs, d_var, w, Pw;
    int64_t ldp = 0;
};

// image with a fully populated dispersion of its 2m image coordinates (jaicov_set_image_dispersion)
struct ImgSigma {
    int img = -1;
    std::vector<double> sigma;     // packed upper, (2m)(2m+1)/2
    DevBuf<double> Pw;             // sigma0^2 Sigma^-1, rp x rp (identity padding)
    int64_t rp = 0;
};

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// One NCCL communicator per process (one process per GPU): created by the first jaicov_dist_init and re-used by later
// handles, so that a sequence of adjustments does not pay communicator set-up again.
static DistContext g_dist;
static bool g_dist_ready = false;

}  // namespace jaicov

using namespace jaicov;

struct jaicov_handle {
    jaicov_options opt{};
    std::string err;
    // host copies of the flattened object graph
    std::vector<double> io_val, r0, coef_val, eo_val, xy, var, rho, xyz, bar_len, bar_var;
    std::vector<int32_t> io_col, coef_ptr, coef_type, coef_order, coef_col, cam_of_img, eo_col, obj_idx, pt_col, bar_a, bar_b;
    std::vector<int64_t> pt_ptr;
    std::vector<uint8_t> is_datum;
    std::vector<Group> groups;
    std::vector<ImgSigma> img_sigma;
    DevBuf<double> ds_Ac, ds_T, ds_G, ds_v, ds_t;   // scratch of the dense-dispersion images (sized for the largest one)
    DevBuf<uint8_t> d_img_dense;
    int free_flags[7] = {0, 0, 0, 0, 0, 0, 0};
    int n_unknowns = -1, n_observations = 0;
    bool has_datum_call = false;
    // device
    DevProblem P;
    AssemblyScratch S;
    DevBuf<double> d_io_val, d_r0, d_coef_val, d_zern_c, d_eo_val, d_pose, d_xy, d_var, d_rho, d_xyz, d_bar_len, d_bar_var;
    DevBuf<int32_t> d_io_col, d_coef_ptr, d_coef_type, d_coef_order, d_coef_col, d_zern_m, d_zern_ptr, d_zern_p, d_cam_kbase,
        d_campos_col, d_cam_of_img, d_eo_col, d_obj_idx, d_img_of_obs, d_pt_col, d_bar_a, d_bar_b, d_img_work_ptr, d_datum_pts;
    DevBuf<int64_t> d_pt_ptr, d_pt_obs_ptr, d_pt_obs;
    DevBuf<WorkItem> d_work;
    DevBuf<int32_t> d_cam_std;
    DevBuf<double> d_img_partial, d_cam_partial, d_pt_partial, d_omega_partial, d_coef_r0pow, d_rw, d_dxp;
    std::vector<int32_t> kbase_host;     // start of every camera's raw parameters in the camera-parameter list
    std::vector<PtGroup> pt_groups;      // camera groups of the by-point sweep (one group unless the cameras have > 68 raw parameters in total)
    DevBuf<double> M, W, Dinv, rhs, V, Bt, Btv, Rt, H, Tq, small, dxref, omega_parts;
    DevBuf<int> info;
    DevBuf<unsigned long long> upd;
    int64_t m_obs = 0;                   // image points
    bool obs_on_device = false;          // jaicov_set_image_points copied the observations straight to the device
    int nDatumPts = 0, free_mask = 0;
    bool prepared = false, have_qxx = false, have_neq = false;
    bool gathered_qxx = false;           // single-process multi-GPU, device 0: M holds the gathered Qxx (lower) instead of the factor
    bool resident = false;               // the device still holds the problem of the last estimate (values centred), although `prepared` was cleared
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t evk[6] = {};             // around the three observation sweeps (by-image, by-point, Omega) of the last pass
    bool evk_omega = false;
    jaicov_stats stats{};
    double centroid[3] = {0, 0, 0};
    // multi-GPU (one process per GPU); the communicator is process-wide and outlives the handle
    DistContext &dist;
    bool dist_on = false;
    // single-process multi-GPU handle (jaicov_options.n_devices > 1): this handle only fans the calls out to one ordinary
    // distributed handle per device, each with its own communicator of one ncclCommInitAll clique and its own host thread
    std::vector<jaicov_handle *> sub;
    std::vector<DistContext *> sub_ctx;
    bool group_ready = false;
    jaicov_handle() : dist(g_dist) {}
    explicit jaicov_handle(DistContext &d) : dist(d) {}
    int panel_tiles = 16;                // block-column panel width of the distributed Cholesky, in 128-tiles.  2048 columns since the
                                         // int8 digit products became the default: their rate grows with the contraction length
                                         // (36 / 42 / 45 TFLOP/s FP64-equivalent at K = 1024 / 2048 / 4096, profiles/r02_gemm_k_sweep.log);
                                         // config 5: 3388 -> 3163 ms on 2 GPUs, 1807 -> 1733 ms on 4 (32 tiles: 1790).  JAICOV_PANEL_TILES
    bool owner_only = false;             // distributed dense route: M holds only this rank's own block-column panels (np x ldo),
                                         // every rank evaluates all observations and keeps what it owns (no all-reduce of N), the
                                         // factor is streamed panel by panel (DenseSchedule::factor_solve_invert_streamed)
    bool obs_sharded = false;            // the observation sweeps run on this rank's image shard only (replicated layout)
    int64_t ldo = 0;                     // leading dimension of M with owner-only storage (128 * own tiles)
    SysView sys;                         // where the entries of N live on this rank
    DevBuf<int32_t> d_tile_lcol;
    DevBuf<double> solve_partial;        // owner-only storage: partial sums of the column-oriented skinny backward solve
    DevBuf<double> gather;               // single-process handle, device 0: the gathered Qxx (np x np lower) when M is owner-only
    cudaEvent_t ev_phase = nullptr;
    DevBuf<double> Xl;                   // np x (128 * ntc): this rank's column tiles of the inverse
    DevBuf<int32_t> d_ktab, d_col_local, d_ptab;
    std::vector<int32_t> ktab;           // first columns of the 128-wide tiles of the INVERSE this rank computes
    std::vector<int32_t> ptab;           // first columns of the tiles of the FACTOR this rank updates (its Cholesky panels)
    DevBuf<double> d_cam_sum;
    int64_t strip_row0 = -1;             // first row of the EO strip of N that is all-reduced
    // MatrixInversion.REDUCED / PRE_ELIMINATION (BA:261-291): numRows of the reduced system (BA:262).  The Schur
    // complement on the EO blocks is not formed: its solution and cofactor matrix are the leading numRows block of the
    // full solution / inverse, which the tensor-core path computes anyway; only that block is exposed.
    int reduced_rows = -1;
    // Levenberg-Marquardt state of estimateModel (BA:89, :93-97, :207-211)
    double adapted_damping = 0.0, lm_omega = 0.0, last_valid_max_abs_dx = 0.0;
    bool derive_first_damping = false;
    // structured route (structured.cu): block-diagonal object-point block, reduced camera system
    struct Structured {
        bool on = false;
        StructDims D{};
        int nBlk = 0;
        int ntp = 0;     // multi-GPU: number of this rank's inverse tiles that hold object-coordinate columns
        DevBuf<int32_t> blk_start, blk_size, col_blk;
        DevBuf<double> Pinv, Zt, Yt, T1t, Kp, Sm, Wm, Dinv, Eb, ED, Fb, zp, rp, yr, ys;
    } st;
    bool wants_inverse() const { return opt.invert_mode != JAICOV_INVERT_NONE; }
    int qxx_rows() const {
        const int n = P.u + P.d;
        return (opt.invert_mode == JAICOV_INVERT_REDUCED || opt.invert_mode == JAICOV_INVERT_PRE_ELIMINATION) ? std::min(n, reduced_rows) : n;
    }
};

namespace {

int fail(jaicov_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    return code;
}

int usable_devices() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

#define API_GUARD_BEGIN try {
#define API_GUARD_END(h)                                                                                      \
    }                                                                                                         \
    catch (const CudaError &e) {                                                                              \
        char buf[512];                                                                                        \
        snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e.code, cudaGetErrorString(e.code), e.what, e.file, e.line); \
        cudaGetLastError();                                                                                   \
        return fail(h, e.code == cudaErrorMemoryAllocation ? JAICOV_OUT_OF_MEMORY : JAICOV_NOT_INITIALISED, buf); \
    }                                                                                                         \
    catch (const std::bad_alloc &) { return fail(h, JAICOV_OUT_OF_MEMORY, "host allocation failed"); }        \
    catch (const std::exception &e) { return fail(h, JAICOV_ILLEGAL_ARGUMENT, e.what()); }

bool active(int32_t c) { return c >= 0 && c != JAICOV_COL_FIXED; }

// MathExtension.binomial, MathExtension.java:53-64
long long binomial(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    long long r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}

// Zernike radial terms of coefficient index j (ZernikeCoefficient.ZernikePolynomial, parameter/ZernikeCoefficient.java:40-56)
void zernike_terms(int order, int &m, std::vector<int32_t> &p, std::vector<double> &c) {
    const int n = (int)std::ceil((-3 + std::sqrt((double)(9 + 8 * order))) / 2);
    m = 2 * order - n * (n + 2);
    const int halfnm = (n - std::abs(m)) / 2;
    const double length = std::sqrt((double)((1 + ((m != 0) ? 1 : 0)) * (n + 1)) / 3.14159265358979323846);
    for (int k = 0; k <= halfnm; k++) {
        p.push_back(n - 2 * k);
        c.push_back(length * (double)(((k % 2 == 0) ? 1 : -1) * binomial(n - k, k) * binomial(n - 2 * k, halfnm - k)));
    }
}

// Decides whether the structured route applies and builds its tables.  It needs the object coordinates to be the
// leading columns (true whenever no coordinate gets its column late through a scale bar or an observed group,
// BA:724-771), every point's unknown components in consecutive columns, and no observation that couples two
// points (scale bars PDF:210-283, observed groups with point targets PDF:447-473) -- then N[p, p] is block diagonal.
void select_solver(jaicov_handle *h) {
    jaicov_handle::Structured &st = h->st;
    st.on = false;
    int want = h->opt.solver;
    if (const char *e = getenv("JAICOV_SOLVER")) {
        if (!strcmp(e, "dense")) want = JAICOV_SOLVER_DENSE;
        else if (!strcmp(e, "structured")) want = JAICOV_SOLVER_STRUCTURED;
    }
    const char *why = nullptr;
    const DevProblem &P = h->P;
    const int d = P.d;
    std::vector<int32_t> blk_start, blk_size, col_blk;
    int up = 0;
    if (want == JAICOV_SOLVER_DENSE) why = "dense route requested";
    else if (!h->bar_a.empty()) why = "scale bars couple object points";
    if (!why && !h->img_sigma.empty()) why = "an image with a fully populated dispersion couples its object points";
    if (!why)
        for (const Group &g : h->groups)
            for (int k : g.kind)
                if (k == 0) { why = "a directly observed group targets object coordinates"; break; }
    if (!why) {
        const size_t nPt = h->pt_col.size() / 3;
        int64_t cnt = 0, mx = -1;
        std::vector<std::pair<int32_t, int32_t>> blocks;   // (start, size)
        for (size_t p = 0; p < nPt && !why; p++) {
            int32_t first = -1, prev = -1, sz = 0;
            for (int c = 0; c < 3; c++) {
                const int32_t col = h->pt_col[3 * p + c];
                if (!active(col)) continue;
                const int32_t e = col - d;
                if (first < 0) first = e;
                else if (e != prev + 1) why = "components of an object point are not in consecutive columns";
                prev = e;
                sz++;
                cnt++;
                mx = std::max<int64_t>(mx, e);
            }
            if (sz) blocks.emplace_back(first, sz);
        }
        if (!why && (cnt == 0 || mx + 1 != cnt)) why = "object coordinates are not the leading columns";
        up = (int)cnt;
        if (!why) {
            auto below = [&](const std::vector<int32_t> &cols) {
                for (int32_t c : cols)
                    if (active(c) && c - d < up) return true;
                return false;
            };
            if (below(h->io_col) || below(h->coef_col) || below(h->eo_col)) why = "a camera or image parameter precedes an object coordinate";
        }
        if (!why && P.u - up <= 0) why = "no camera or image unknowns";
        if (!why && round_up(P.u - up + d, kBlk) > 65535) why = "reduced system too large for one grid dimension";
        if (!why) {
            std::sort(blocks.begin(), blocks.end());
            col_blk.assign(up, -1);
            for (size_t b = 0; b < blocks.size(); b++) {
                blk_start.push_back(blocks[b].first);
                blk_size.push_back(blocks[b].second);
                for (int k = 0; k < blocks[b].second; k++) col_blk[blocks[b].first + k] = (int32_t)b;
            }
            for (int32_t b : col_blk)
                if (b < 0) { why = "object coordinate columns are not contiguous"; break; }
        }
    }
    if (why) {
        if (want == JAICOV_SOLVER_STRUCTURED) throw std::runtime_error(std::string("structured solver not applicable: ") + why);
        return;
    }
    st.on = true;
    StructDims &D = st.D;
    D.up = up; D.nc = P.u - up; D.d = d; D.u = P.u;
    D.Tp = round_up(up, kBlk); D.mp = round_up(D.nc + d, kBlk); D.ncp = round_up(D.nc, kBlk); D.np = P.np;
    st.nBlk = (int)blk_start.size();
    st.blk_start.upload(blk_start); st.blk_size.upload(blk_size); st.col_blk.upload(col_blk);
    st.Pinv.alloc((size_t)st.nBlk * 9);
    const size_t zt = (size_t)D.mp * D.Tp;
    st.Zt.alloc(zt); st.Yt.alloc(zt);
    st.Kp.alloc((size_t)D.mp * D.mp);
    st.Sm.alloc((size_t)D.ncp * D.ncp); st.Wm.alloc((size_t)D.ncp * D.ncp); st.Dinv.alloc((size_t)D.ncp * kBlk);
    st.Eb.alloc(8 * (size_t)D.ncp); st.ED.alloc(8 * (size_t)D.ncp); st.Fb.alloc(8 * (size_t)D.ncp);
    st.zp.alloc((size_t)D.Tp); st.rp.alloc((size_t)D.mp); st.yr.alloc((size_t)D.mp); st.ys.alloc((size_t)P.np);
}

void prepare(jaicov_handle *h) {
    if (h->prepared) return;
    NvtxRange nvtx_("jaicov: prepare (upload, index structures, buffers)");
    if (usable_devices() == 0) throw CudaError{cudaErrorNoDevice, "no sm_100 device: jaicov_b200 has no CPU path", __FILE__, __LINE__};
    JCHECK(cudaSetDevice(h->opt.device));
    if (!h->stream) {
        JCHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        for (auto &e : h->ev) JCHECK(cudaEventCreate(&e));
        for (auto &e : h->evk) JCHECK(cudaEventCreate(&e));
    }
    if (!h->has_datum_call || h->n_unknowns < 0) throw std::runtime_error("jaicov_set_datum has not been called");
    DevProblem &P = h->P;
    P = DevProblem();
    int d = 0, mask = 0;
    for (int i = 0; i < 7; i++) if (h->free_flags[i]) { d++; mask |= 1 << i; }
    h->free_mask = mask;
    P.d = d;
    P.u = h->n_unknowns;
    P.np = round_up(std::max(P.u, 1), kBlk);
    P.sigma2 = h->opt.sigma2apriori > 0 ? h->opt.sigma2apriori : 1.0;   // BA:221
    // ---- cameras -----------------------------------------------------------------------------------------------
    P.nCam = (int)h->r0.size();
    P.nCoef = (int)h->coef_val.size();
    if (h->coef_ptr.empty()) h->coef_ptr.assign(1, 0);
    if ((int)h->coef_ptr.size() != P.nCam + 1 || h->coef_ptr[0] != 0 || h->coef_ptr.back() != P.nCoef)
        throw std::runtime_error("coef_ptr must have n_cam + 1 entries from 0 to the number of coefficients");
    for (int c = 0; c < P.nCam; c++)
        if (h->coef_ptr[c + 1] < h->coef_ptr[c]) throw std::runtime_error("coef_ptr is not monotone");
    for (int k = 0; k < P.nCoef; k++) {
        const int t = h->coef_type[k];
        if (t != JAICOV_PT_RADIAL_A && t != JAICOV_PT_TANGENTIAL_B && t != JAICOV_PT_TANGENTIAL_BX && t != JAICOV_PT_TANGENTIAL_BY &&
            t != JAICOV_PT_AFFINITY_CX && t != JAICOV_PT_AFFINITY_CY && t != JAICOV_PT_DISTANCE_D && t != JAICOV_PT_ZERNIKE_X &&
            t != JAICOV_PT_ZERNIKE_Y && t != JAICOV_PT_ZERNIKE_Z)
            throw std::runtime_error("coef_type: not a distortion ParameterType id");
        if (h->coef_order[k] < 0 || h->coef_order[k] > 100000) throw std::runtime_error("coef_order out of range");
        // the evaluation consumes Cx,Cy and Bx,By as pairs (AffinityShear... / Tangential...Factory): each must be followed by its partner
    }
    for (int c = 0; c < P.nCam; c++)
        for (int k = h->coef_ptr[c]; k < h->coef_ptr[c + 1]; k++) {
            const int t = h->coef_type[k];
            if ((t == JAICOV_PT_AFFINITY_CX || t == JAICOV_PT_TANGENTIAL_BX) && (k + 1 >= h->coef_ptr[c + 1] || h->coef_type[k + 1] != t + 1))
                throw std::runtime_error("coef_type: Cx / Bx must be followed by Cy / By of the same camera");
        }
    std::vector<int32_t> zm(P.nCoef, 0), zptr(P.nCoef + 1, 0), zp;
    std::vector<double> zc;
    int maxcoef = 0;
    for (int k = 0; k < P.nCoef; k++) {
        zptr[k] = (int32_t)zp.size();
        const int t = h->coef_type[k];
        if (t == JAICOV_PT_ZERNIKE_X || t == JAICOV_PT_ZERNIKE_Y || t == JAICOV_PT_ZERNIKE_Z) {
            int m;
            zernike_terms(h->coef_order[k], m, zp, zc);
            zm[k] = m;
        }
    }
    zptr[P.nCoef] = (int32_t)zp.size();
    if (zp.empty()) { zp.push_back(0); zc.push_back(0.0); }
    std::vector<int32_t> kbase(P.nCam + 1, 0), campos;
    for (int c = 0; c < P.nCam; c++) {
        const int nc = h->coef_ptr[c + 1] - h->coef_ptr[c];
        if (nc > kMaxCoef) throw std::runtime_error("more than 64 distortion coefficients per camera are not supported");
        maxcoef = std::max(maxcoef, nc);
        kbase[c + 1] = kbase[c] + 3 + nc;
        for (int i = 0; i < 3; i++) campos.push_back(h->io_col[3 * c + i]);
        for (int k = h->coef_ptr[c]; k < h->coef_ptr[c + 1]; k++) campos.push_back(h->coef_col[k]);
    }
    P.kRaw = kbase[P.nCam];
    h->S = AssemblyScratch();
    h->S.ntImg = (9 + maxcoef + 1 + 7) / 8;
    h->S.kcMax = 3 + maxcoef;
    if (h->S.ntImg > 9) throw std::runtime_error("more than 62 distortion coefficients in one camera are not supported (by-image Gram row of 72 columns)");
    // camera groups of the by-point sweep: consecutive cameras while [X Y Z | raw parameters | w] fits 72 Gram columns
    h->pt_groups.clear();
    h->S.ntPt = 1;
    for (int c = 0; c < P.nCam;) {
        PtGroup g{c, c, 0, 0};
        while (g.cam1 < P.nCam && (g.cam1 == g.cam0 || 3 + g.kraw + (kbase[g.cam1 + 1] - kbase[g.cam1]) + 1 <= 72)) {
            g.kraw += kbase[g.cam1 + 1] - kbase[g.cam1];
            g.cam1++;
        }
        g.nt = (3 + g.kraw + 1 + 7) / 8;
        h->S.ntPt = std::max(h->S.ntPt, g.nt);
        h->pt_groups.push_back(g);
        c = g.cam1;
    }
    if (h->pt_groups.empty()) h->pt_groups.push_back(PtGroup{0, 0, 0, 1});
    // canonical coefficient lists ([Cx Cy]? [Bx By B1..]? [A1..] [D1..], orders consecutive from 1 -- what the reference's readers
    // build, AICONReportFileReader.java:308): the sweeps then evaluate them straight-line (model.cuh, STD) instead of interpreting
    std::vector<int32_t> cam_std((size_t)std::max(P.nCam, 1) * 5, 0);
    bool all_std = P.nCam > 0 && getenv("JAICOV_SWEEP_GENERIC") == nullptr;
    for (int c = 0; c < P.nCam; c++) {
        int k = h->coef_ptr[c];
        const int k1 = h->coef_ptr[c + 1];
        int32_t *t = &cam_std[5 * (size_t)c];
        auto type_at = [&](int i) { return i < k1 ? h->coef_type[i] : -1; };
        if (type_at(k) == JAICOV_PT_AFFINITY_CX && type_at(k + 1) == JAICOV_PT_AFFINITY_CY) { t[0] = 1; k += 2; }
        if (type_at(k) == JAICOV_PT_TANGENTIAL_BX && type_at(k + 1) == JAICOV_PT_TANGENTIAL_BY) {
            t[1] = 1; k += 2;
            while (type_at(k) == JAICOV_PT_TANGENTIAL_B && h->coef_order[k] == t[2] + 1) { t[2]++; k++; }
        }
        while (type_at(k) == JAICOV_PT_RADIAL_A && h->coef_order[k] == t[3] + 1) { t[3]++; k++; }
        while (type_at(k) == JAICOV_PT_DISTANCE_D && h->coef_order[k] == t[4] + 1) { t[4]++; k++; }
        if (k != k1) { t[0] = -1; all_std = false; }
    }
    // r0^(2 order) of every coefficient (the constant of the radial / distance polynomials; repeated multiplication like the kernels)
    std::vector<double> r0pow(std::max(P.nCoef, 1), 0.0);
    for (int c = 0; c < P.nCam; c++)
        for (int k = h->coef_ptr[c]; k < h->coef_ptr[c + 1]; k++) r0pow[k] = ipow(h->r0[c] * h->r0[c], h->coef_order[k]);
    // ---- images / observations -----------------------------------------------------------------------------------
    P.nImg = (int)h->cam_of_img.size();
    P.m = h->m_obs;
    if (h->pt_ptr.empty()) h->pt_ptr.assign(1, 0);
    if (h->pt_ptr.back() != P.m) throw std::runtime_error("pt_ptr does not cover the image points");
    P.nPt = (int)(h->xyz.size() / 3);
    // a foreign caller's bad index must come back as JAICOV_ILLEGAL_ARGUMENT, not as an out-of-bounds device read (which is
    // sticky for the CUDA context): O(nImg + nBar + nCoef) range / monotonicity checks on everything a kernel indexes with
    if ((int)h->pt_ptr.size() != P.nImg + 1 || h->pt_ptr[0] != 0) throw std::runtime_error("pt_ptr must have n_img + 1 entries starting at 0");
    for (int i = 0; i < P.nImg; i++) {
        if (h->pt_ptr[i + 1] < h->pt_ptr[i]) throw std::runtime_error("pt_ptr is not monotone");
        if (h->cam_of_img[i] < 0 || h->cam_of_img[i] >= P.nCam) throw std::runtime_error("cam_of_img: camera index out of range");
    }
    for (size_t b = 0; b < h->bar_a.size(); b++)
        if (h->bar_a[b] < 0 || h->bar_a[b] >= P.nPt || h->bar_b[b] < 0 || h->bar_b[b] >= P.nPt)
            throw std::runtime_error("scale bar end point index out of range");
    for (const Group &g : h->groups)
        for (int i = 0; i < g.r; i++) {
            const int k = g.kind[i], ix = g.index[i], cp = g.comp[i];
            const bool ok = (k == 0 && ix >= 0 && ix < P.nPt && cp >= 0 && cp < 3) || (k == 1 && ix >= 0 && ix < P.nCam && cp >= 0 && cp < 3) ||
                            (k == 2 && ix >= 0 && ix < P.nCoef) || (k == 3 && ix >= 0 && ix < P.nImg && cp >= 0 && cp < 6);
            if (!ok) throw std::runtime_error("observed group: target kind / index / component out of range");
        }
    // image shard of this rank: contiguous image ranges balanced by observation count.  Owner-only storage (the
    // distributed dense route): every rank sweeps ALL observations and keeps the entries it owns -- 3.4 ms of replicated sweeps
    // per pass at config 5 instead of all-reducing 1.5 GB of N.
    select_solver(h);       // needs only host data; decides the route (and with it the storage layout) before anything is sharded
    h->owner_only = false;
    if (h->dist_on && h->dist.world > 1 && !h->st.on) {
        const char *e = getenv("JAICOV_DIST_STORAGE");
        h->owner_only = !(e && !strcmp(e, "replica"));
    }
    h->obs_sharded = h->dist_on && h->dist.world > 1 && !h->owner_only;
    {
        int32_t i0 = 0, i1 = P.nImg;
        jaicov_shard_images(P.nImg, h->pt_ptr.data(), h->obs_sharded ? h->dist.world : 1, h->obs_sharded ? h->dist.rank : 0, &i0, &i1);
        P.img0 = i0;
        P.img1 = i1;
        P.obs0 = h->pt_ptr[P.img0];
        P.obs1 = h->pt_ptr[P.img1];
    }
    std::vector<WorkItem> work;
    std::vector<int32_t> img_work_ptr(P.img1 - P.img0 + 1, 0);
    const int64_t chunk = 1024;
    for (int i = P.img0; i < P.img1; i++) {
        img_work_ptr[i - P.img0] = (int32_t)work.size();
        for (int64_t b = h->pt_ptr[i]; b < h->pt_ptr[i + 1]; b += chunk)
            work.push_back(WorkItem{i, 0, b, std::min(b + chunk, h->pt_ptr[i + 1])});
    }
    img_work_ptr[P.img1 - P.img0] = (int32_t)work.size();
    // observations: already on the device (jaicov_set_image_points), or host copies taken while no device was visible
    if (!h->obs_on_device) {
        h->d_obj_idx.upload(h->obj_idx); h->d_xy.upload(h->xy); h->d_var.upload(h->var); h->d_rho.upload(h->rho);
    }
    h->d_pt_ptr.upload(h->pt_ptr);
    // (cudaMemcpy from pageable memory may return while the last staging buffer is still in flight on the legacy stream,
    // and the handle's stream is non-blocking: order the kernels below after those copies explicitly)
    JCHECK(cudaStreamSynchronize(0));
    // index structures on the device: image of every observation, observations of every object point (stable sort)
    h->d_img_of_obs.alloc((size_t)std::max<int64_t>(P.m, 1));
    launch_img_of_obs(h->d_pt_ptr.p, P.nImg, h->d_img_of_obs.p, h->stream);
    {
        const int64_t nloc = P.obs1 - P.obs0;
        DevBuf<int32_t> keys;
        DevBuf<int64_t> iota;
        DevBuf<unsigned char> temp;
        keys.alloc((size_t)std::max<int64_t>(nloc, 1));
        iota.alloc((size_t)std::max<int64_t>(nloc, 1));
        const size_t tb = csc_temp_bytes(std::max<int64_t>(nloc, 1));
        temp.alloc(std::max<size_t>(tb, 1));
        h->d_pt_obs_ptr.alloc((size_t)P.nPt + 1);
        h->d_pt_obs.alloc((size_t)std::max<int64_t>(nloc, 1));
        int32_t mm[2];
        launch_build_csc(h->d_obj_idx.p, P.obs0, nloc, P.nPt, keys.p, iota.p, temp.p, tb, h->d_pt_obs_ptr.p, h->d_pt_obs.p, mm, h->stream);
        if (nloc > 0 && (mm[0] < 0 || mm[1] >= P.nPt)) throw std::runtime_error("object point index out of range");
    }
    // ---- validate columns ----------------------------------------------------------------------------------------
    auto check_cols = [&](const std::vector<int32_t> &c) {
        for (int32_t x : c)
            if (active(x) && (x < d || x >= P.u + d)) throw std::runtime_error("column index outside [d, u+d)");
    };
    check_cols(h->io_col); check_cols(h->coef_col); check_cols(h->eo_col); check_cols(h->pt_col);
    // datum points: isDatum and no fixed component (BA:501-513)
    std::vector<int32_t> datum_pts;
    for (int p = 0; p < P.nPt; p++) {
        if (!h->is_datum.empty() && !h->is_datum[p]) continue;
        if (h->is_datum.empty()) continue;
        const int32_t *c = &h->pt_col[3 * (size_t)p];
        if (c[0] == JAICOV_COL_FIXED || c[1] == JAICOV_COL_FIXED || c[2] == JAICOV_COL_FIXED) continue;
        if (!active(c[0]) || !active(c[1]) || !active(c[2])) continue;
        datum_pts.push_back(p);
    }
    h->nDatumPts = (int)datum_pts.size();
    if (d > 0 && h->nDatumPts < 3) throw std::runtime_error("not enough object points to realise the frame datum (BA:515-516)");
    // ---- upload --------------------------------------------------------------------------------------------------
    h->d_io_val.upload(h->io_val); h->d_io_col.upload(h->io_col); h->d_r0.upload(h->r0);
    h->d_coef_ptr.upload(h->coef_ptr); h->d_coef_type.upload(h->coef_type); h->d_coef_order.upload(h->coef_order);
    h->d_coef_val.upload(h->coef_val); h->d_coef_col.upload(h->coef_col); h->d_coef_r0pow.upload(r0pow);
    h->d_cam_std.upload(cam_std);
    P.cam_std = h->d_cam_std.p;
    h->d_zern_m.upload(zm); h->d_zern_ptr.upload(zptr); h->d_zern_p.upload(zp); h->d_zern_c.upload(zc);
    h->d_cam_kbase.upload(kbase); h->d_campos_col.upload(campos);
    h->kbase_host = kbase;
    h->d_cam_of_img.upload(h->cam_of_img); h->d_eo_val.upload(h->eo_val); h->d_eo_col.upload(h->eo_col);
    h->d_pose.alloc((size_t)std::max(P.nImg, 1) * 16);
    h->d_xyz.upload(h->xyz); h->d_pt_col.upload(h->pt_col);
    h->d_bar_a.upload(h->bar_a); h->d_bar_b.upload(h->bar_b); h->d_bar_len.upload(h->bar_len); h->d_bar_var.upload(h->bar_var);
    h->d_work.upload(work); h->d_img_work_ptr.upload(img_work_ptr); h->d_datum_pts.upload(datum_pts);
    P.io_val = h->d_io_val.p; P.io_col = h->d_io_col.p; P.r0 = h->d_r0.p; P.coef_ptr = h->d_coef_ptr.p;
    P.coef_type = h->d_coef_type.p; P.coef_order = h->d_coef_order.p; P.coef_col = h->d_coef_col.p; P.coef_val = h->d_coef_val.p; P.coef_r0pow = h->d_coef_r0pow.p;
    P.zern_m = h->d_zern_m.p; P.zern_ptr = h->d_zern_ptr.p; P.zern_p = h->d_zern_p.p; P.zern_c = h->d_zern_c.p;
    P.cam_kbase = h->d_cam_kbase.p; P.campos_col = h->d_campos_col.p;
    P.cam_of_img = h->d_cam_of_img.p; P.eo_val = h->d_eo_val.p; P.eo_col = h->d_eo_col.p; P.pt_ptr = h->d_pt_ptr.p;
    P.pose = h->d_pose.p; P.obj_idx = h->d_obj_idx.p; P.xy = h->d_xy.p; P.var = h->d_var.p; P.rho = h->d_rho.p;
    P.img_of_obs = h->d_img_of_obs.p; P.pt_obs_ptr = h->d_pt_obs_ptr.p; P.pt_obs = h->d_pt_obs.p;
    P.xyz = h->d_xyz.p; P.pt_col = h->d_pt_col.p;
    P.nBar = (int)h->bar_a.size();
    P.bar_a = h->d_bar_a.p; P.bar_b = h->d_bar_b.p; P.bar_len = h->d_bar_len.p; P.bar_var = h->d_bar_var.p;
    AssemblyScratch &S = h->S;
    S.std_eval = all_std ? 1 : 0;
    S.nWork = (int)work.size();
    S.work = h->d_work.p; S.img_work_ptr = h->d_img_work_ptr.p;
    const int NCi = 8 * S.ntImg, NCp = 8 * S.ntPt;
    h->d_img_partial.alloc((size_t)std::max(S.nWork, 1) * NCi * NCi);
    h->d_cam_partial.alloc((size_t)std::max(P.nImg, 1) * S.kcMax * (S.kcMax + 1));
    h->d_cam_sum.alloc((size_t)std::max(P.nCam, 1) * S.kcMax * (S.kcMax + 1));
    S.cam_sum = h->d_cam_sum.p;
    h->d_pt_partial.alloc((size_t)std::max(P.nPt, 1) * 3 * NCp);
    h->d_omega_partial.alloc((size_t)std::max(S.nWork, 1) + 2);
    h->d_dxp.alloc((size_t)std::max(P.nPt, 1) * 3);
    S.dxp = h->d_dxp.p;
    h->d_rw.alloc((size_t)std::max<int64_t>(P.m, 1) * 3);
    P.rw = h->d_rw.p;
    launch_obs_weights(P, h->d_rw.p, h->stream);
    // ---- images with a fully populated dispersion (extension): P = (Sigma / sigma0^2)^-1 once, the standard sweeps skip them -----------
    {
        std::vector<uint8_t> dense(std::max(P.nImg, 1), 0);
        int64_t rp_max = 0;
        if (!h->img_sigma.empty() && h->dist_on) throw std::runtime_error("images with a fully populated dispersion need a single device (their point blocks are not sharded)");
        for (ImgSigma &is : h->img_sigma) {
            if (is.img < 0 || is.img >= P.nImg) throw std::runtime_error("jaicov_set_image_dispersion: image index out of range");
            const int64_t m2 = 2 * (h->pt_ptr[is.img + 1] - h->pt_ptr[is.img]);
            if ((int64_t)is.sigma.size() != m2 * (m2 + 1) / 2) throw std::runtime_error("jaicov_set_image_dispersion: needs (2m)(2m+1)/2 entries for the image's m points");
            if (m2 == 0) continue;
            dense[is.img] = 1;
            is.rp = round_up(m2, kBlk);
            rp_max = std::max(rp_max, is.rp);
            launch_zero_weights(h->d_rw.p, h->pt_ptr[is.img], h->pt_ptr[is.img + 1], h->stream);
            DevBuf<double> ap, Wg, Dg;
            DevBuf<int> inf;
            ap.upload(is.sigma);
            is.Pw.alloc((size_t)is.rp * is.rp);
            Wg.alloc((size_t)is.rp * is.rp);
            Dg.alloc((size_t)is.rp * kBlk);
            inf.alloc(1);
            launch_unpack_scaled(ap.p, (int)m2, 1.0 / P.sigma2, is.Pw.p, is.rp, is.rp, h->stream);
            JCHECK(cudaMemsetAsync(inf.p, 0, sizeof(int), h->stream));
            CudaBackend be{h->stream, inf.p};
            DenseSchedule<CudaBackend> ds{be, is.Pw.p, is.rp, is.rp, Dg.p};
            ds.potrf();
            ds.invert_from_factor(Wg.p);
            launch_symmetrize(is.Pw.p, is.rp, (int)m2, h->stream);
            int info = 0;
            JCHECK(cudaMemcpyAsync(&info, inf.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            JCHECK(cudaStreamSynchronize(h->stream));
            if (info != 0) throw std::runtime_error("dispersion matrix of an image is not positive definite");
        }
        h->d_img_dense.upload(dense);
        P.img_dense = h->img_sigma.empty() ? nullptr : h->d_img_dense.p;
        if (rp_max) {
            h->ds_Ac.alloc((size_t)rp_max * kBlk); h->ds_T.alloc((size_t)rp_max * kBlk); h->ds_G.alloc((size_t)kBlk * kBlk);
            h->ds_v.alloc((size_t)rp_max); h->ds_t.alloc((size_t)rp_max);
        }
    }
    S.img_partial = h->d_img_partial.p; S.cam_partial = h->d_cam_partial.p; S.pt_partial = h->d_pt_partial.p;
    S.omega_partial = h->d_omega_partial.p;
    // ---- system buffers ------------------------------------------------------------------------------------------
    const size_t np = (size_t)P.np;
    if (h->owner_only && h->st.on) throw std::runtime_error("internal: owner-only storage chosen for the structured route");
    if (!h->owner_only) h->M.alloc(np * np);
    h->sys = SysView{h->M.p, P.np, nullptr};
    if (h->dist_on) {
        // Factor: block-column panels of panel_tiles tiles, owner = panel % world.
        // Inverse: tile c costs ~ (nb - c)^2, so its tiles are dealt out one by one in snake order
        // (0..W-1, W-1..0, ...) which balances the quadratic cost to within a fraction of a percent.
        const int nb = (int)(P.np / kBlk), W = h->dist.world, R = h->dist.rank;
        h->ktab.clear();
        h->ptab.clear();
        std::vector<int32_t> col_local(nb, -1);
        for (int c = 0; c < nb; c++) {
            const int q = c % (2 * W);
            const int owner = q < W ? q : 2 * W - 1 - q;
            if (owner == R) {
                col_local[c] = (int32_t)h->ktab.size();
                h->ktab.push_back(c * kBlk);
            }
            if ((c / h->panel_tiles) % W == R) h->ptab.push_back(c * kBlk);
        }
        h->d_col_local.upload(col_local);
        {
            std::vector<int32_t> pt = h->ptab;
            if (pt.empty()) pt.push_back(0);
            h->d_ptab.upload(pt);
        }
        if (h->owner_only) {
            // own tiles side by side in ascending order: local tile index = position in ptab
            std::vector<int32_t> lcol(nb, -1);
            for (size_t lt = 0; lt < h->ptab.size(); lt++) lcol[h->ptab[lt] / kBlk] = (int32_t)(lt * kBlk);
            h->d_tile_lcol.upload(lcol);
            h->ldo = (int64_t)std::max<size_t>(h->ptab.size(), 1) * kBlk;
            h->M.alloc(np * (size_t)h->ldo);
            h->sys = SysView{h->M.p, h->ldo, h->d_tile_lcol.p};
            if (!h->ev_phase) JCHECK(cudaEventCreate(&h->ev_phase));
            h->solve_partial.alloc(((size_t)P.np / 256 + 2) * 8 * 128);
        }
        std::vector<int32_t> kt = h->ktab;
        if (kt.empty()) kt.push_back(0);
        h->d_ktab.upload(kt);
        if (h->wants_inverse()) h->Xl.alloc(np * (size_t)kBlk * std::max<size_t>(h->ktab.size(), 1));
        // EO strip of N: every entry the image sweeps write lies in a row >= the smallest EO row
        int64_t r0 = -1;
        for (int32_t c : h->eo_col)
            if (active(c)) r0 = (r0 < 0) ? c - d : std::min<int64_t>(r0, c - d);
        h->strip_row0 = h->owner_only ? -1 : r0;
    } else if (h->wants_inverse() && !h->st.on) {
        h->W.alloc(np * np);
    }
    if (h->st.on && h->wants_inverse()) {
        // Q'Y': all object-coordinate columns on one GPU, this rank's object-coordinate tiles (a prefix of ktab) otherwise
        const StructDims &D = h->st.D;
        h->st.ntp = 0;
        for (int32_t c : h->ktab)
            if (c < D.Tp) h->st.ntp++;
        h->st.T1t.alloc((size_t)D.mp * (h->dist_on ? (size_t)std::max(h->st.ntp, 1) * kBlk : (size_t)D.Tp));
    }
    if ((h->opt.invert_mode == JAICOV_INVERT_REDUCED || h->opt.invert_mode == JAICOV_INVERT_PRE_ELIMINATION) && h->reduced_rows < 0)
        throw std::runtime_error("invert_mode REDUCED / PRE_ELIMINATION needs jaicov_set_reduced_rows (numRows of BA:262)");
    h->Dinv.alloc(np * kBlk);
    h->rhs.alloc(np); h->V.alloc(np);
    h->Bt.alloc(8 * np); h->Btv.alloc(8 * np); h->H.alloc(8 * np); h->Tq.alloc(8 * np);
    h->Rt.alloc((size_t)kRhsRows * np);   // 8 right-hand-side rows: n, datum rows
    h->small.alloc(128); h->dxref.alloc(np + 8); h->omega_parts.alloc(4 + h->groups.size());
    h->info.alloc(1); h->upd.alloc(2);
    JCHECK(cudaMemset(h->dxref.p, 0, (np + 8) * sizeof(double)));
    // ---- directly observed groups ----------------------------------------------------------------------------------
    int gi = 0;
    for (Group &g : h->groups) {
        std::vector<const double *> tp(g.r);
        std::vector<int32_t> col(g.r);
        for (int i = 0; i < g.r; i++) {
            const int k = g.kind[i], ix = g.index[i], cp = g.comp[i];
            if (k == 0) { tp[i] = P.xyz + 3 * (size_t)ix + cp; col[i] = h->pt_col.at(3 * (size_t)ix + cp); }
            else if (k == 1) { tp[i] = P.io_val + 3 * (size_t)ix + cp; col[i] = h->io_col.at(3 * (size_t)ix + cp); }
            else if (k == 2) { tp[i] = P.coef_val + ix; col[i] = h->coef_col.at(ix); }
            else if (k == 3) { tp[i] = P.eo_val + 6 * (size_t)ix + cp; col[i] = h->eo_col.at(6 * (size_t)ix + cp); }
            else throw std::runtime_error("unknown target kind in observed group");
        }
        g.tptr.upload(tp); g.col.upload(col); g.d_obs.upload(g.obs); g.w.alloc(g.r);
        if (g.sigma.empty()) {
            g.d_var.upload(g.var);
        } else {
            // P = (Sigma / sigma0^2)^-1 (DOPG:80-90) with the same blocked Cholesky + inverse as the main system
            const int64_t rp = round_up(g.r, kBlk);
            DevBuf<double> ap, Wg, Dg;
            ap.upload(g.sigma);
            g.Pw.alloc((size_t)rp * rp);
            Wg.alloc((size_t)rp * rp);
            Dg.alloc((size_t)rp * kBlk);
            g.ldp = rp;
            launch_unpack_scaled(ap.p, g.r, 1.0 / P.sigma2, g.Pw.p, rp, rp, h->stream);
            JCHECK(cudaMemsetAsync(h->info.p, 0, sizeof(int), h->stream));
            CudaBackend be{h->stream, h->info.p};
            DenseSchedule<CudaBackend> ds{be, g.Pw.p, rp, rp, Dg.p};
            ds.potrf();
            ds.invert_from_factor(Wg.p);
            launch_symmetrize(g.Pw.p, rp, g.r, h->stream);
            int info = 0;
            JCHECK(cudaMemcpyAsync(&info, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            JCHECK(cudaStreamSynchronize(h->stream));
            if (info != 0) throw std::runtime_error("dispersion matrix of a directly observed group is not positive definite");
        }
        gi++;
    }
    (void)gi;
    JCHECK(cudaStreamSynchronize(0));     // every upload of this function has landed before the first pass starts
    h->prepared = true;
    h->have_qxx = false;
    h->have_neq = false;
}

// stages 1-2 of one pass: N, n of the current values in M (lower) / rhs, datum rows in Bt
// (sparse_clear: the structured route clears only the point blocks and the rows of the camera / image unknowns)
void assemble(jaicov_handle *h, bool sparse_clear = false) {
    NvtxRange nvtx_("jaicov: assembly N = A'PA, n = A'Pw");
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    const size_t np = (size_t)P.np;
    if (sparse_clear && h->st.on) {
        const size_t up = (size_t)h->st.D.up;
        JCHECK(cudaMemsetAsync(h->M.p + up * np, 0, (np - up) * np * sizeof(double), s));
        launch_zero_point_blocks(h->M.p, P.np, h->st.blk_start.p, h->st.blk_size.p, h->st.nBlk, s);
    } else {
        JCHECK(cudaMemsetAsync(h->M.p, 0, np * (h->owner_only ? (size_t)h->ldo : np) * sizeof(double), s));
    }
    JCHECK(cudaMemsetAsync(h->rhs.p, 0, np * sizeof(double), s));
    JCHECK(cudaMemsetAsync(h->Bt.p, 0, 8 * np * sizeof(double), s));
    launch_pose(P, s);
    JCHECK(cudaEventRecord(h->evk[0], s));
    launch_assemble_images(P, h->S, h->sys, h->rhs.p, s);
    JCHECK(cudaEventRecord(h->evk[1], s));
    const bool multi = h->obs_sharded;      // image shards: the shared pieces are summed over the ranks
    const AssemblyScratch &S = h->S;
    for (size_t gi = 0; gi < h->pt_groups.size(); gi++) {
        const PtGroup &g = h->pt_groups[gi];
        if (gi == 0) JCHECK(cudaEventRecord(h->evk[2], s));
        launch_by_point(P, S, g, s);
        if (gi == 0) JCHECK(cudaEventRecord(h->evk[3], s));
        // the only exchange of the assembly: sum the shared pieces over the image shards
        if (multi) h->dist.allreduce_sum(S.pt_partial, (size_t)P.nPt * 3 * 8 * g.nt, s);
        if (gi == 0) {
            if (multi) {
                h->dist.allreduce_sum(S.cam_sum, (size_t)P.nCam * S.kcMax * (S.kcMax + 1), s);
                h->dist.allreduce_sum(h->rhs.p, (size_t)P.np, s);
                if (h->strip_row0 >= 0)
                    h->dist.allreduce_sum(h->M.p + (size_t)h->strip_row0 * np, ((size_t)P.np - h->strip_row0) * np, s);
            }
            launch_camera_scatter(P, S, h->sys, h->rhs.p, s);
        }
        launch_point_scatter(P, S, g, h->kbase_host[g.cam0], h->sys, h->rhs.p, s);
    }
    for (ImgSigma &is : h->img_sigma)
        if (is.rp)
            launch_dense_image_assemble(P, is.img, h->pt_ptr[is.img + 1] - h->pt_ptr[is.img], is.Pw.p, is.rp, is.rp, h->ds_Ac.p, h->ds_T.p,
                                        h->ds_G.p, h->M.p, h->rhs.p, s);
    launch_scale_bars(P, h->sys, h->rhs.p, s);
    for (Group &g : h->groups) {
        launch_group_w(g.r, g.tptr.p, g.d_obs.p, g.w.p, s);
        launch_group_stack(g.r, g.col.p, g.d_var.p, g.Pw.p, g.ldp, P.sigma2, g.w.p, P.d, h->sys, h->rhs.p, s);
    }
    if (P.d > 0) launch_datum_rows(P.xyz, P.pt_col, h->d_datum_pts.p, h->nDatumPts, h->free_mask, P.d, P.np, h->Bt.p, s);
    JCHECK(cudaGetLastError());
    h->have_neq = true;
    h->have_qxx = false;
    h->gathered_qxx = false;
}

struct PassResult {
    int info = 0;
    double max_abs_dx = 0.0;
    bool bad = false;
    double omega = 0.0;
    bool lm_step = false, lm_accepted = true;
    double lm_last = 0.0;
};

// Omega part of the images with a fully populated dispersion, added to omega_parts[0] (after launch_omega has written it)
void omega_dense_images(jaicov_handle *h) {
    for (ImgSigma &is : h->img_sigma)
        if (is.rp)
            launch_dense_image_omega(h->P, is.img, is.Pw.p, is.rp, is.rp, h->ds_Ac.p, h->dxref.p, h->ds_v.p, h->ds_t.p, h->omega_parts.p, h->stream);
}

// sum of the Omega parts of all observation groups for the dx currently in dxref (device -> host, synchronises)
double omega_now(jaicov_handle *h, bool multi) {
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    launch_omega(P, h->S, h->dxref.p, h->omega_parts.p, s);
    if (multi) h->dist.allreduce_sum(h->omega_parts.p, 1, s);
    omega_dense_images(h);
    if (P.nBar) launch_omega_bars(P, h->dxref.p, h->omega_parts.p + 1, s);
    int gi = 0;
    for (Group &g : h->groups) {
        launch_group_omega(g.r, g.col.p, g.d_var.p, g.Pw.p, g.ldp, P.sigma2, g.w.p, h->dxref.p, h->omega_parts.p + 2 + gi, s);
        gi++;
    }
    std::vector<double> om(2 + h->groups.size(), 0.0);
    JCHECK(cudaMemcpyAsync(om.data(), h->omega_parts.p, om.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    JCHECK(cudaStreamSynchronize(s));
    double t = om[0];
    if (P.nBar) t += om[1];
    for (size_t g = 0; g < h->groups.size(); g++) t += om[2 + g];
    return t;
}

PassResult run_pass(jaicov_handle *h, bool final_pass, bool apply_update) {
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    const size_t np = (size_t)P.np;
    const bool invert = final_pass && h->wants_inverse();
    NvtxRange nvtx_pass(final_pass ? "jaicov: final pass" : "jaicov: pass");
    JCHECK(cudaEventRecord(h->ev[0], s));
    assemble(h, true);
    nvtxRangePushA("jaicov: precondition + datum");
    // Levenberg-Marquardt: N_cc += lambda N_cc on every unknown column, before the preconditioner (BA:801-822)
    if (h->derive_first_damping) {
        h->adapted_damping = h->opt.damping_value;
        h->derive_first_damping = false;
    }
    if (h->adapted_damping > 0) launch_damp_diag(h->sys, P.u, h->adapted_damping, s);
    // preconditioner and SPD reformulation (K4)
    launch_precond_diag(h->sys, P.u, P.np, h->V.p, s);
    if (h->owner_only) h->dist.allreduce_sum(h->V.p, (size_t)P.np, s);      // every rank contributed the entries of its own diagonal
    // (the structured route keeps the datum rows as a border of its reduced system instead of folding B'B into N)
    if (h->owner_only)
        launch_scale_system_own(h->M.p, h->ldo, h->d_ptab.p, (int)h->ptab.size(), P.u, h->V.p, h->Bt.p, P.d, P.np, s);
    else
        launch_scale_system(h->M.p, P.np, P.u, h->V.p, h->Bt.p, h->st.on ? 0 : P.d, P.np, h->st.on ? h->st.D.up : 0, s);
    JCHECK(cudaMemsetAsync(h->Rt.p, 0, (size_t)kRhsRows * np * sizeof(double), s));
    launch_build_rhs(h->Rt.p, h->Btv.p, P.np, P.u, h->V.p, h->rhs.p, h->Bt.p, P.d, h->opt.estimation_type == JAICOV_SIMULATION, s);
    h->have_neq = false;
    JCHECK(cudaEventRecord(h->ev[1], s));
    nvtxRangePop();
    nvtxRangePushA("jaicov: factor");
    // factor (K5)
    JCHECK(cudaMemsetAsync(h->info.p, 0, sizeof(int), s));
    CudaBackend be{s, h->info.p};
    DenseSchedule<CudaBackend> ds{be, h->M.p, P.np, P.np, h->Dinv.p};
    const bool multi = h->dist_on && h->dist.world > 1;
    jaicov_handle::Structured &st = h->st;
    if (st.on) {
        // reduced system K' = [[R, B_r'],[B_r, 0]] - Z'Y and its inverse Q' (structured.cu)
        const StructDims &D = st.D;
        launch_point_block_inv(h->M.p, P.np, st.blk_start.p, st.blk_size.p, st.nBlk, h->V.p, st.Pinv.p, h->info.p, s);
        launch_build_zy(h->M.p, h->Btv.p, D, st.col_blk.p, st.blk_start.p, st.blk_size.p, st.Pinv.p, st.Zt.p, st.Yt.p, s);
        {
            // Z'Y contracts over the object coordinates: with several GPUs every rank takes a slice of that range and the
            // m x m partial products are summed (GEMM -> all-reduce; the result is bitwise the same on every rank)
            const int64_t nkt = D.Tp / kBlk;
            const int64_t kt0 = multi ? h->dist.rank * nkt / h->dist.world : 0, kt1 = multi ? (h->dist.rank + 1) * nkt / h->dist.world : nkt;
            if (!multi || h->dist.rank == 0) launch_init_kp(h->M.p, h->Btv.p, D, st.Kp.p, s);
            else JCHECK(cudaMemsetAsync(st.Kp.p, 0, (size_t)D.mp * D.mp * sizeof(double), s));
            if (kt1 > kt0) {
                GemmDesc g;
                g.al = 0; g.bl = 0; g.mt = g.nt = (int)(D.mp / kBlk); g.K = (kt1 - kt0) * kBlk; g.alpha = -1.0; g.beta = 1.0;
                g.A = st.Zt.p + kt0 * kBlk; g.lda = D.Tp; g.B = st.Yt.p + kt0 * kBlk; g.ldb = D.Tp; g.C = st.Kp.p; g.ldc = D.mp; g.tri_out = 1;
                be.gemm(g);
            }
            if (multi) h->dist.allreduce_sum(st.Kp.p, (size_t)D.mp * D.mp, s);
        }
        if (D.d > 0) launch_border_prep(st.Kp.p, D, st.Eb.p, st.ED.p, h->small.p, s);
        launch_form_stilde(st.Kp.p, D, st.Eb.p, st.ED.p, st.Sm.p, s);
        DenseSchedule<CudaBackend> dr{be, st.Sm.p, D.ncp, D.ncp, st.Dinv.p};
        dr.potrf();
        dr.invert_from_factor(st.Wm.p);
        launch_symmetrize(st.Sm.p, D.ncp, D.nc, s);
        launch_border_f(st.Sm.p, D, st.ED.p, st.Fb.p, h->small.p, s);
        launch_fill_qprime(st.Sm.p, D, st.Fb.p, h->small.p, st.Kp.p, s);
    } else if (multi && h->owner_only) {
        // owner-only storage: factorisation, the solves for n and the datum rows, and (final pass) both triangular sweeps of this
        // rank's columns of the inverse in ONE schedule that consumes every panel where the broadcast lands it
        const int ntc = invert ? (int)h->ktab.size() : 0;
        const int64_t ldx = (int64_t)std::max(ntc, 1) * kBlk;
        if (ntc > 0) {
            JCHECK(cudaMemsetAsync(h->Xl.p, 0, np * (size_t)ldx * sizeof(double), s));
            launch_identity_columns(h->Xl.p, ldx, P.np, h->d_ktab.p, ntc, s);
        }
        StreamPanelComm sc{&h->dist, s, h->M.p, h->ldo, P.np, h->Dinv.p, h->panel_tiles, h->ev_phase};
        sc.ensure_stage((size_t)P.np * h->panel_tiles * kBlk + (size_t)h->panel_tiles * kBlk * kBlk);
        be.solve_partial = h->solve_partial.p;
        DenseSchedule<CudaBackend> dso{be, h->M.p, h->ldo, P.np, h->Dinv.p};
        dso.factor_solve_invert_streamed(sc, h->dist.rank, h->dist.world, h->panel_tiles, h->d_ptab.p, (int)h->ptab.size(), h->ptab.data(),
                                         h->Rt.p, h->Rt.p + 8 * np, ntc > 0 ? h->Xl.p : nullptr, ldx, ntc, h->d_ktab.p, true);
        JCHECK(cudaEventRecord(h->dist.ev_tmp, h->dist.net));
        JCHECK(cudaStreamWaitEvent(s, h->dist.ev_tmp, 0));
    } else if (multi) {
        PanelComm pc{&h->dist, s, h->M.p, P.np, P.np, h->Dinv.p, h->panel_tiles};
        pc.ensure_stage((size_t)P.np * h->panel_tiles * kBlk + (size_t)h->panel_tiles * kBlk * kBlk);
        ds.potrf_distributed(pc, h->dist.rank, h->dist.world, h->panel_tiles, h->ptab.empty() ? nullptr : h->d_ptab.p,
                             (int)h->ptab.size(), h->ptab.data());
        // the last broadcasts this rank rooted are still on the network stream: later stages (and the next pass,
        // which re-uses the staging buffers) are ordered after them
        JCHECK(cudaEventRecord(h->dist.ev_tmp, h->dist.net));
        JCHECK(cudaStreamWaitEvent(s, h->dist.ev_tmp, 0));
    } else {
        ds.potrf();
    }
    JCHECK(cudaEventRecord(h->ev[2], s));
    nvtxRangePop();
    nvtxRangePushA("jaicov: solve + datum correction");
    // solve for n and the datum rows, datum correction, dx (K5/K9)
    if (st.on) {
        // (H doubles as the d null-space rows K^-1[lambda, x]; small[100..107) holds the datum residual B y)
        launch_structured_solution(h->Rt.p, st.D, st.col_blk.p, st.blk_start.p, st.blk_size.p, st.Pinv.p, st.Zt.p, st.Yt.p, st.Kp.p,
                                   h->Btv.p, h->V.p, st.zp.p, st.rp.p, st.yr.p, st.ys.p, h->H.p, h->small.p + 100, h->dxref.p, h->Tq.p, s);
    } else {
        if (!h->owner_only) launch_solve_rows8(h->M.p, P.np, h->Dinv.p, h->Rt.p, h->Rt.p + 8 * np, P.np, s);
        launch_datum_solve(h->Rt.p, h->Btv.p, P.d, P.np, P.u, h->V.p, h->dxref.p, h->H.p, h->Tq.p, h->small.p, s);
    }
    // Levenberg-Marquardt step control (updateModel, BA:390-426): shorten the step, compare Omega, accept or reject
    PassResult r;
    bool apply_dx = apply_update;
    if (h->adapted_damping > 0) {
        const double SQRT_EPS = std::sqrt(kEps);
        const double alpha = std::min(0.25 * std::pow(h->adapted_damping, -0.05), 0.75);
        launch_scale_vector(h->dxref.p, (int64_t)P.u + P.d, alpha, s);
        double prev = h->lm_omega;
        const double cur = omega_now(h, h->obs_sharded);
        prev = prev <= 0 ? 1.7976931348623157e308 : prev;
        const bool converge = prev >= cur;
        h->lm_omega = cur;
        r.lm_step = true;
        r.lm_last = h->adapted_damping;
        r.lm_accepted = converge;
        if (converge) {
            h->adapted_damping *= 0.2;
        } else {
            h->adapted_damping *= 5.0;
            if (h->adapted_damping > 1.0 / SQRT_EPS) { h->adapted_damping = 1.0 / SQRT_EPS; h->lm_omega = 0.0; }
            JCHECK(cudaMemsetAsync(h->dxref.p, 0, ((size_t)P.u + P.d) * sizeof(double), s));   // dx.zero(), BA:423
            apply_dx = false;
        }
    }
    JCHECK(cudaEventRecord(h->ev[3], s));
    nvtxRangePop();
    nvtxRangePushA("jaicov: inverse Qxx");
    // inverse (K6/K7)
    if (invert && st.on && h->dist_on) {
        // the same two products restricted to this rank's column tiles of the inverse (no communication): Q'Y' for its
        // object-coordinate tiles, then the rows on and below each tile's diagonal of Y (Q'Y')
        const StructDims &D = st.D;
        const int ntc = (int)h->ktab.size(), ntp = st.ntp;
        if (ntc > 0) {
            const int64_t ldx = (int64_t)ntc * kBlk, ldt = (int64_t)std::max(ntp, 1) * kBlk;
            JCHECK(cudaMemsetAsync(h->Xl.p, 0, np * (size_t)ldx * sizeof(double), s));
            launch_scale_yt(st.Yt.p, D, h->V.p, s);
            if (ntp > 0) {
                GemmDesc g;
                g.al = 0; g.bl = 1; g.mt = (int)(D.mp / kBlk); g.nt = ntp; g.K = D.mp; g.alpha = 1.0; g.beta = 0.0;
                g.A = st.Kp.p; g.lda = D.mp; g.B = st.Yt.p; g.ldb = D.Tp; g.C = st.T1t.p; g.ldc = ldt;
                g.coltab = h->d_ktab.p; g.ncoltab = ntp; g.coltab_full = 1; g.c_local = 1;
                be.gemm(g);
                GemmDesc q;
                q.al = 1; q.bl = 1; q.mt = (int)(D.Tp / kBlk); q.nt = ntp; q.K = D.mp; q.alpha = 1.0; q.beta = 0.0;
                q.A = st.Yt.p; q.lda = D.Tp; q.B = st.T1t.p; q.ldb = ldt; q.C = h->Xl.p; q.ldc = ldx;
                q.kmode = K_ROW_MASK; q.ktab = h->d_ktab.p; q.roff = 0;
                be.gemm(q);
            }
            launch_structured_place_cols(h->Xl.p, ldx, ntc, h->d_ktab.p, h->d_col_local.p, D, st.T1t.p, ldt, st.Kp.p, st.blk_start.p,
                                         st.blk_size.p, st.nBlk, st.Pinv.p, h->V.p, s);
        }
    } else if (invert && st.on) {
        // K^-1[p, r|lambda] = -(Q' Y')', K^-1[p, p] = P^-1 + Y (Q' Y'): two tensor-core products, then placement and V scaling
        // (Yt is scaled by V first, so both products carry the V (.) V scaling of K7 and no pass over the 16 GB result is needed)
        const StructDims &D = st.D;
        launch_scale_yt(st.Yt.p, D, h->V.p, s);
        GemmDesc g;
        g.al = 0; g.bl = 1; g.mt = (int)(D.mp / kBlk); g.nt = (int)(D.Tp / kBlk); g.K = D.mp; g.alpha = 1.0; g.beta = 0.0;
        g.A = st.Kp.p; g.lda = D.mp; g.B = st.Yt.p; g.ldb = D.Tp; g.C = st.T1t.p; g.ldc = D.Tp;
        be.gemm(g);
        GemmDesc q;
        q.al = 1; q.bl = 1; q.mt = q.nt = (int)(D.Tp / kBlk); q.K = D.mp; q.alpha = 1.0; q.beta = 0.0;
        q.A = st.Yt.p; q.lda = D.Tp; q.B = st.T1t.p; q.ldb = D.Tp; q.C = h->M.p; q.ldc = P.np; q.tri_out = 1;
        be.gemm(q);
        launch_structured_place(h->M.p, D, st.T1t.p, st.Kp.p, st.blk_start.p, st.blk_size.p, st.nBlk, st.Pinv.p, h->V.p, s);
    } else if (invert && h->dist_on) {
        // every rank inverts its own column tiles from the replicated factor: no communication
        const int ntc = (int)h->ktab.size();
        if (ntc > 0) {
            const int64_t ldx = (int64_t)ntc * kBlk;
            if (!h->owner_only) {          // (owner-only storage: both sweeps already ran inside the streamed schedule)
                JCHECK(cudaMemsetAsync(h->Xl.p, 0, np * (size_t)ldx * sizeof(double), s));
                launch_identity_columns(h->Xl.p, ldx, P.np, h->d_ktab.p, ntc, s);
                ds.inverse_columns(h->Xl.p, ldx, ntc, h->d_ktab.p);
            }
            launch_qxx_epilogue_cols(h->Xl.p, ldx, ntc, h->d_ktab.p, P.u, h->V.p, h->H.p, h->Rt.p + np, P.d, P.np, s);
        }
    } else if (invert) {
        ds.invert_from_factor(h->W.p);
        launch_qxx_epilogue(h->M.p, P.np, P.u, h->V.p, h->H.p, h->Rt.p + np, P.d, P.np, s);
    }
    JCHECK(cudaEventRecord(h->ev[4], s));
    nvtxRangePop();
    nvtxRangePushA("jaicov: omega + update");
    // Omega at the pre-update point (K8), BA:429-430
    const bool want_omega = final_pass && h->opt.estimation_type != JAICOV_SIMULATION && (!r.lm_step || r.lm_accepted);
    h->evk_omega = want_omega;
    if (want_omega) {
        JCHECK(cudaEventRecord(h->evk[4], s));
        launch_omega(P, h->S, h->dxref.p, h->omega_parts.p, s);
        JCHECK(cudaEventRecord(h->evk[5], s));
        if (h->obs_sharded) h->dist.allreduce_sum(h->omega_parts.p, 1, s);
        omega_dense_images(h);
        if (P.nBar) launch_omega_bars(P, h->dxref.p, h->omega_parts.p + 1, s);
        int gi = 0;
        for (Group &g : h->groups) {
            launch_group_omega(g.r, g.col.p, g.d_var.p, g.Pw.p, g.ldp, P.sigma2, g.w.p, h->dxref.p, h->omega_parts.p + 2 + gi, s);
            gi++;
        }
    }
    JCHECK(cudaEventRecord(h->ev[5], s));
    // update + max|dx| (K9), BA:450-462
    JCHECK(cudaMemsetAsync(h->upd.p, 0, 2 * sizeof(unsigned long long), s));
    launch_update(P.xyz, P.pt_col, 3 * (int64_t)P.nPt, h->dxref.p, apply_dx, h->upd.p, s);
    launch_update(P.io_val, P.io_col, 3 * (int64_t)P.nCam, h->dxref.p, apply_dx, h->upd.p, s);
    launch_update(P.coef_val, P.coef_col, P.nCoef, h->dxref.p, apply_dx, h->upd.p, s);
    launch_update(P.eo_val, P.eo_col, 6 * (int64_t)P.nImg, h->dxref.p, apply_dx, h->upd.p, s);
    JCHECK(cudaEventRecord(h->ev[6], s));
    nvtxRangePop();
    JCHECK(cudaGetLastError());
    // the one host read per pass: status word, max|dx|, Omega
    unsigned long long upd[2];
    std::vector<double> om(2 + h->groups.size(), 0.0);
    double small99 = 0.0;
    JCHECK(cudaMemcpyAsync(&r.info, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    JCHECK(cudaMemcpyAsync(upd, h->upd.p, sizeof upd, cudaMemcpyDeviceToHost, s));
    if (want_omega) JCHECK(cudaMemcpyAsync(om.data(), h->omega_parts.p, om.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (P.d > 0) JCHECK(cudaMemcpyAsync(&small99, h->small.p + 98, sizeof(double), cudaMemcpyDeviceToHost, s));
    JCHECK(cudaStreamSynchronize(s));
    if (small99 != 0.0 && r.info == 0) r.info = -1;
    if (multi) {   // a failed pivot is seen by the panel's owner only: agree on the status
        double flag = r.info != 0 ? 1.0 : 0.0, *dflag = h->small.p + 120;
        JCHECK(cudaMemcpyAsync(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice, s));
        h->dist.allreduce_sum(dflag, 1, s);
        JCHECK(cudaMemcpyAsync(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost, s));
        JCHECK(cudaStreamSynchronize(s));
        if (flag != 0.0 && r.info == 0) r.info = -2;
    }
    memcpy(&r.max_abs_dx, &upd[0], sizeof(double));
    r.bad = upd[1] != 0;
    if (r.lm_step && !r.lm_accepted) r.max_abs_dx = h->last_valid_max_abs_dx;   // BA:420-425
    else h->last_valid_max_abs_dx = r.max_abs_dx;                               // BA:433
    if (want_omega) {
        r.omega = om[0];
        if (P.nBar) r.omega += om[1];
        for (size_t g = 0; g < h->groups.size(); g++) r.omega += om[2 + g];
    }
    float ms[6];
    for (int i = 0; i < 6; i++) cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]);
    if (h->owner_only && multi) {   // one fused schedule: forward phase = factor (+ forward substitutions), backward phase counts as inverse
        float fwd = 0, bwd = 0;
        cudaEventElapsedTime(&fwd, h->ev[1], h->ev_phase);
        cudaEventElapsedTime(&bwd, h->ev_phase, h->ev[2]);
        ms[1] = fwd;
        ms[3] += bwd;
    }
    h->stats.ms_assembly = ms[0]; h->stats.ms_factor = ms[1]; h->stats.ms_solve = ms[2]; h->stats.ms_inverse = ms[3];
    h->stats.ms_omega = ms[4];
    float tot;
    cudaEventElapsedTime(&tot, h->ev[0], h->ev[6]);
    h->stats.ms_total = tot;
    h->have_qxx = invert && r.info == 0;
    h->stats.solver_used = st.on ? JAICOV_SOLVER_STRUCTURED : JAICOV_SOLVER_DENSE;
    return r;
}

void download_values(jaicov_handle *h) {
    const DevProblem &P = h->P;
    auto dl = [&](std::vector<double> &v, const double *p) {
        if (!v.empty()) JCHECK(cudaMemcpy(v.data(), p, v.size() * sizeof(double), cudaMemcpyDeviceToHost));
    };
    dl(h->xyz, P.xyz); dl(h->io_val, P.io_val); dl(h->coef_val, P.coef_val); dl(h->eo_val, P.eo_val);
}

void upload_values(jaicov_handle *h) {
    const DevProblem &P = h->P;
    auto ul = [&](const std::vector<double> &v, double *p) {
        if (!v.empty()) JCHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
    };
    ul(h->xyz, P.xyz); ul(h->io_val, P.io_val); ul(h->coef_val, P.coef_val); ul(h->eo_val, P.eo_val);
    for (Group &g : h->groups)
        if (g.r) JCHECK(cudaMemcpy(g.d_obs.p, g.obs.data(), g.r * sizeof(double), cudaMemcpyHostToDevice));
}

// centroidCoordinates (BA:115-201) on the host copies; sign = -1 before the loop, +1 after
bool centroid_shift(jaicov_handle *h, bool invert) {
    const size_t nPt = h->xyz.size() / 3, nImg = h->eo_val.size() / 6;
    if (!invert) {
        double s[3] = {0, 0, 0};
        long cnt[3] = {0, 0, 0};
        for (size_t p = 0; p < nPt; p++)
            for (int c = 0; c < 3; c++)
                if (active(h->pt_col[3 * p + c])) { s[c] += h->xyz[3 * p + c]; cnt[c]++; }
        for (size_t i = 0; i < nImg; i++)
            for (int c = 0; c < 3; c++)
                if (active(h->eo_col[6 * i + c])) { s[c] += h->eo_val[6 * i + c]; cnt[c]++; }
        if (!(cnt[0] == cnt[1] && cnt[0] == cnt[2] && cnt[0] > 0)) return false;
        for (int c = 0; c < 3; c++) h->centroid[c] = s[c] / (double)cnt[c];
    }
    const double sign = invert ? 1.0 : -1.0;
    for (size_t p = 0; p < nPt; p++)
        for (int c = 0; c < 3; c++)
            if (active(h->pt_col[3 * p + c])) h->xyz[3 * p + c] += sign * h->centroid[c];
    for (size_t i = 0; i < nImg; i++)
        for (int c = 0; c < 3; c++)
            if (active(h->eo_col[6 * i + c])) h->eo_val[6 * i + c] += sign * h->centroid[c];
    for (Group &g : h->groups)
        for (int i = 0; i < g.r; i++) {
            if (g.kind[i] == 0) g.obs[i] += sign * h->centroid[g.comp[i]];
            else if (g.kind[i] == 3 && g.comp[i] < 3) g.obs[i] += sign * h->centroid[g.comp[i]];
        }
    return true;
}

// BundleAdjustment.interrupt (:240-245, :320-325): reads and clears the caller's flag.  Several GPUs: the decision must be the same
// on every rank (a rank that leaves the loop alone would leave the others waiting in a collective), so the ranks agree on it
// -- one scalar all-reduce; either every rank passes a flag or none does.
bool interrupted(jaicov_handle *h, volatile int32_t *flag) {
    if (!flag) return false;
    int local = 0;
    if (*flag) { *flag = 0; local = 1; }
    if (h->dist_on && h->dist.world > 1) {
        double v = local, *dflag = h->small.p + 121;
        JCHECK(cudaMemcpyAsync(dflag, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->dist.allreduce_sum(dflag, 1, h->stream);
        JCHECK(cudaMemcpyAsync(&v, dflag, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        JCHECK(cudaStreamSynchronize(h->stream));
        local = v != 0.0;
    }
    return local != 0;
}

// ---- single-process multi-GPU handle -----------------------------------------------------------------------------------------
bool is_group(const jaicov_handle *h) { return h && !h->sub.empty(); }

// f(sub handle, index) on every device's handle: index 0 on the CALLING thread (progress callbacks stay on it, as the ABI
// promises), the others on one host thread each.  Result: the common return code, or the first differing (failing) one with its
// error text copied to the group handle.
template <class F>
int group_run(jaicov_handle *g, F &&f) {
    const size_t n = g->sub.size();
    std::vector<int> rc(n, JAICOV_OK);
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t i = 1; i < n; i++) th.emplace_back([&rc, &f, g, i] { rc[i] = f(g->sub[i], (int)i); });
    rc[0] = f(g->sub[0], 0);
    for (auto &t : th) t.join();
    for (size_t i = 0; i < n; i++)
        if (rc[i] != rc[0] || (rc[i] < 0 && rc[i] != JAICOV_NO_CONVERGENCE && rc[i] != JAICOV_SINGULAR_MATRIX && rc[i] != JAICOV_INTERRUPT)) {
            const size_t k = rc[i] != rc[0] ? (rc[i] < rc[0] ? i : 0) : i;
            g->err = "device " + std::to_string(g->sub[k]->opt.device) + ": " + g->sub[k]->err;
            return rc[k];
        }
    return rc[0];
}

// communicators of the clique, created by the first computing call (one ncclCommInitAll from the calling thread)
void group_ensure(jaicov_handle *g) {
    if (g->group_ready) return;
    const int n = (int)g->sub.size();
    if (usable_devices() < g->opt.device + n)
        throw CudaError{cudaErrorNoDevice, "not enough sm_100 devices for jaicov_options.n_devices (jaicov_b200 has no CPU path)", __FILE__, __LINE__};
    std::vector<int> devs(n);
    std::vector<void *> comms(n, nullptr);
    for (int i = 0; i < n; i++) devs[i] = g->opt.device + i;
    nccl_comm_init_all(comms.data(), n, devs.data());
    for (int i = 0; i < n; i++) {
        JCHECK(cudaSetDevice(devs[i]));
        g->sub_ctx[i]->adopt(i, n, comms[i]);
    }
    JCHECK(cudaSetDevice(devs[0]));
    g->group_ready = true;
}

#define GROUP_FORWARD(h, CALL)                                        \
    if (is_group(h)) {                                                \
        for (jaicov_handle *s_ : (h)->sub) {                          \
            const int rc_ = CALL;                                     \
            if (rc_ != JAICOV_OK) { (h)->err = s_->err; return rc_; } \
        }                                                             \
        return JAICOV_OK;                                             \
    }

}  // namespace

// =====================================================================================================================
extern "C" {

int32_t jaicov_default_options(jaicov_options *opt) {
    if (!opt) return JAICOV_ILLEGAL_ARGUMENT;
    opt->invert_mode = JAICOV_INVERT_FULL;
    opt->estimation_type = JAICOV_L2NORM;
    opt->max_iterations = 5000;
    opt->use_centroid = 1;
    opt->apply_aposteriori = 1;
    opt->device = 0;
    opt->solver = JAICOV_SOLVER_AUTO;
    opt->n_devices = 1;
    opt->sigma2apriori = 1.0;
    opt->damping_value = 0.0;
    return JAICOV_OK;
}

int32_t jaicov_device_count(void) { return usable_devices(); }

int64_t jaicov_release_cached_memory(void) { return (int64_t)(g_cache.purge() + ozaki_release_scratch()); }

int64_t jaicov_launch_count(void) { return (int64_t)g_launch_count.load(); }

int32_t jaicov_set_gemm_digits(int32_t digits) { return ozaki_set_digits(digits); }

int32_t jaicov_create(const jaicov_options *opt, jaicov_handle **out) {
    if (!opt || !out) return JAICOV_ILLEGAL_ARGUMENT;
    *out = nullptr;
    if (!(opt->damping_value >= 0.0)) return JAICOV_ILLEGAL_ARGUMENT;
    if (opt->invert_mode < JAICOV_INVERT_NONE || opt->invert_mode > JAICOV_INVERT_REDUCED) return JAICOV_ILLEGAL_ARGUMENT;
    if (opt->estimation_type != JAICOV_L2NORM && opt->estimation_type != JAICOV_SIMULATION) return JAICOV_ILLEGAL_ARGUMENT;
    if (opt->solver < JAICOV_SOLVER_AUTO || opt->solver > JAICOV_SOLVER_STRUCTURED) return JAICOV_ILLEGAL_ARGUMENT;
    if (opt->n_devices < 0 || opt->n_devices > 64 || opt->device < 0) return JAICOV_ILLEGAL_ARGUMENT;
    jaicov_handle *h = new (std::nothrow) jaicov_handle();
    if (!h) return JAICOV_OUT_OF_MEMORY;
    h->opt = *opt;
    if (opt->n_devices > 1) {
        // single-process multi-GPU: one ordinary distributed handle per device [device, device + n_devices)
        try {
            for (int i = 0; i < opt->n_devices; i++) {
                DistContext *ctx = new DistContext();
                ctx->rank = i;
                ctx->world = opt->n_devices;
                h->sub_ctx.push_back(ctx);
                jaicov_handle *sh = new jaicov_handle(*ctx);
                sh->opt = *opt;
                sh->opt.device = opt->device + i;
                sh->opt.n_devices = 1;
                sh->dist_on = true;
                if (const char *e = getenv("JAICOV_PANEL_TILES")) sh->panel_tiles = std::max(1, atoi(e));
                h->sub.push_back(sh);
            }
        } catch (const std::bad_alloc &) {
            jaicov_destroy(h);
            return JAICOV_OUT_OF_MEMORY;
        }
    }
    *out = h;
    return JAICOV_OK;
}

void jaicov_destroy(jaicov_handle *h) {
    if (!h) return;
    if (is_group(h) || !h->sub_ctx.empty()) {
        for (jaicov_handle *sh : h->sub) jaicov_destroy(sh);
        for (DistContext *ctx : h->sub_ctx) {
            if (ctx->comm && usable_devices() > 0) { cudaSetDevice(h->opt.device + ctx->rank); cudaDeviceSynchronize(); ctx->destroy(); }
            delete ctx;
        }
        delete h;
        return;
    }
    if (usable_devices() > 0) cudaSetDevice(h->opt.device);      // buffers may exist without a stream (set_* before the first pass)
    if (h->stream) {
        cudaStreamSynchronize(h->stream);
        if (h->dist_on) cudaDeviceSynchronize();   // the process-wide communicator stays
        for (auto &e : h->ev) if (e) cudaEventDestroy(e);
        for (auto &e : h->evk) if (e) cudaEventDestroy(e);
        if (h->ev_phase) cudaEventDestroy(h->ev_phase);
        cudaStreamDestroy(h->stream);
    }
    delete h;
}

const char *jaicov_last_error(const jaicov_handle *h) { return h ? h->err.c_str() : "null handle"; }

int32_t jaicov_shard_images(int32_t n_img, const int64_t *pt_ptr, int32_t world, int32_t rank, int32_t *img_begin, int32_t *img_end) {
    if (n_img < 0 || !pt_ptr || world < 1 || rank < 0 || rank >= world || !img_begin || !img_end) return JAICOV_ILLEGAL_ARGUMENT;
    const int64_t m = pt_ptr[n_img];
    auto bound = [&](int r) -> int32_t {
        if (r <= 0) return 0;
        if (r >= world) return n_img;
        const int64_t target = m * r / world;
        return (int32_t)(std::lower_bound(pt_ptr, pt_ptr + n_img + 1, target) - pt_ptr);
    };
    const int32_t b0 = std::min(bound(rank), n_img);
    *img_begin = b0;
    *img_end = std::min(std::max(bound(rank + 1), b0), n_img);
    return JAICOV_OK;
}

int32_t jaicov_nccl_unique_id(void *out128) {
    if (!out128) return JAICOV_ILLEGAL_ARGUMENT;
    jaicov_handle *h = nullptr;
    API_GUARD_BEGIN
    NcclUniqueId id;
    nccl_unique_id(&id);
    memcpy(out128, &id, sizeof id);
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_dist_init(jaicov_handle *h, int32_t rank, int32_t world, const void *nccl_id128) {
    if (!h || world < 1 || rank < 0 || rank >= world || !nccl_id128) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h) || &h->dist != &g_dist) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "handle already spans several devices of this process (jaicov_options.n_devices)");
    API_GUARD_BEGIN
    if (usable_devices() == 0) throw CudaError{cudaErrorNoDevice, "no sm_100 device: jaicov_b200 has no CPU path", __FILE__, __LINE__};
    JCHECK(cudaSetDevice(h->opt.device));
    if (g_dist_ready && (g_dist.rank != rank || g_dist.world != world))
        return fail(h, JAICOV_ILLEGAL_ARGUMENT, "this process already joined a communicator with another rank/world");
    if (!g_dist_ready) {
        NcclUniqueId id;
        memcpy(&id, nccl_id128, sizeof id);
        g_dist.init(rank, world, id);
        g_dist_ready = true;
    }
    h->dist_on = true;
    if (const char *e = getenv("JAICOV_PANEL_TILES")) h->panel_tiles = std::max(1, atoi(e));
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_set_cameras(jaicov_handle *h, int32_t n_cam, const double *io_val, const int32_t *io_col, const double *r0,
                           const int32_t *coef_ptr, const int32_t *coef_type, const int32_t *coef_order, const double *coef_val,
                           const int32_t *coef_col) {
    if (!h || n_cam < 0) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_set_cameras(s_, n_cam, io_val, io_col, r0, coef_ptr, coef_type, coef_order, coef_val, coef_col))
    if (n_cam > 0 && (!io_val || !io_col || !r0 || !coef_ptr)) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "jaicov_set_cameras: null array");
    if (n_cam > 0 && coef_ptr[n_cam] > 0 && (!coef_type || !coef_order || !coef_val || !coef_col))
        return fail(h, JAICOV_ILLEGAL_ARGUMENT, "jaicov_set_cameras: null coefficient array");
    if (n_cam == 0) {
        h->io_val.clear(); h->io_col.clear(); h->r0.clear(); h->coef_ptr.assign(1, 0);
        h->coef_type.clear(); h->coef_order.clear(); h->coef_val.clear(); h->coef_col.clear();
        h->prepared = false; h->resident = false;
        return JAICOV_OK;
    }
    API_GUARD_BEGIN
    h->io_val.assign(io_val, io_val + 3 * (size_t)n_cam);
    h->io_col.assign(io_col, io_col + 3 * (size_t)n_cam);
    h->r0.assign(r0, r0 + n_cam);
    h->coef_ptr.assign(coef_ptr, coef_ptr + n_cam + 1);
    const size_t nc = (size_t)h->coef_ptr[n_cam];
    h->coef_type.assign(coef_type, coef_type + nc);
    h->coef_order.assign(coef_order, coef_order + nc);
    h->coef_val.assign(coef_val, coef_val + nc);
    h->coef_col.assign(coef_col, coef_col + nc);
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_set_images(jaicov_handle *h, int32_t n_img, const int32_t *cam_of_img, const double *eo_val, const int32_t *eo_col,
                          const int64_t *pt_ptr) {
    if (!h || n_img < 0) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_set_images(s_, n_img, cam_of_img, eo_val, eo_col, pt_ptr))
    if (!pt_ptr || (n_img > 0 && (!cam_of_img || !eo_val || !eo_col))) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "jaicov_set_images: null array");
    API_GUARD_BEGIN
    h->cam_of_img.assign(cam_of_img, cam_of_img + n_img);
    h->eo_val.assign(eo_val, eo_val + 6 * (size_t)n_img);
    h->eo_col.assign(eo_col, eo_col + 6 * (size_t)n_img);
    h->pt_ptr.assign(pt_ptr, pt_ptr + n_img + 1);
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_set_image_points(jaicov_handle *h, int64_t m, const int32_t *obj_idx, const double *xy, const double *var,
                                const double *rho) {
    if (!h || m < 0) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h))   // every device takes its copy over its own PCIe link, in parallel
        return group_run(h, [&](jaicov_handle *sh, int) { return (int)jaicov_set_image_points(sh, m, obj_idx, xy, var, rho); });
    API_GUARD_BEGIN
    if (m > 0 && (!obj_idx || !xy || !var)) return JAICOV_ILLEGAL_ARGUMENT;
    h->m_obs = m;
    if (usable_devices() > 0) {
        // the observations are only ever read by kernels: copy them from the caller's buffers straight to the device
        JCHECK(cudaSetDevice(h->opt.device));
        h->d_obj_idx.upload(obj_idx, (size_t)m);
        h->d_xy.upload(xy, (size_t)(2 * m));
        h->d_var.upload(var, (size_t)(2 * m));
        if (rho) h->d_rho.upload(rho, (size_t)m);
        else { h->d_rho.alloc((size_t)m); if (m) JCHECK(cudaMemset(h->d_rho.p, 0, (size_t)m * sizeof(double))); }
        h->obs_on_device = true;
        h->obj_idx.clear(); h->xy.clear(); h->var.clear(); h->rho.clear();
    } else {
        // no device visible: keep host copies so that the failure surfaces at the first computing call (no CPU path)
        h->obj_idx.assign(obj_idx, obj_idx + m);
        h->xy.assign(xy, xy + 2 * m);
        h->var.assign(var, var + 2 * m);
        if (rho) h->rho.assign(rho, rho + m); else h->rho.assign(m, 0.0);
        h->obs_on_device = false;
    }
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_set_object_points(jaicov_handle *h, int32_t n_pt, const double *xyz, const int32_t *col, const uint8_t *is_datum) {
    if (!h || n_pt < 0) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_set_object_points(s_, n_pt, xyz, col, is_datum))
    if (n_pt > 0 && (!xyz || !col)) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "jaicov_set_object_points: null array");
    API_GUARD_BEGIN
    h->xyz.assign(xyz, xyz + 3 * (size_t)n_pt);
    h->pt_col.assign(col, col + 3 * (size_t)n_pt);
    if (is_datum) h->is_datum.assign(is_datum, is_datum + n_pt); else h->is_datum.assign(n_pt, 0);
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_set_scale_bars(jaicov_handle *h, int32_t n_bar, const int32_t *a, const int32_t *b, const double *length,
                              const double *var) {
    if (!h || n_bar < 0) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_set_scale_bars(s_, n_bar, a, b, length, var))
    if (n_bar > 0 && (!a || !b || !length || !var)) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "jaicov_set_scale_bars: null array");
    API_GUARD_BEGIN
    h->bar_a.assign(a, a + n_bar); h->bar_b.assign(b, b + n_bar);
    h->bar_len.assign(length, length + n_bar); h->bar_var.assign(var, var + n_bar);
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_add_observed_group(jaicov_handle *h, int32_t r, const int32_t *target_kind, const int32_t *target_index,
                                  const int32_t *target_comp, const double *obs, const double *var, const double *sigma_packed_upper) {
    if (!h || r <= 0 || (!var && !sigma_packed_upper)) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_add_observed_group(s_, r, target_kind, target_index, target_comp, obs, var, sigma_packed_upper))
    if (!target_kind || !target_index || !target_comp || !obs) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "jaicov_add_observed_group: null array");
    API_GUARD_BEGIN
    h->groups.emplace_back();
    Group &g = h->groups.back();
    g.r = r;
    g.kind.assign(target_kind, target_kind + r);
    g.index.assign(target_index, target_index + r);
    g.comp.assign(target_comp, target_comp + r);
    g.obs.assign(obs, obs + r);
    if (sigma_packed_upper) g.sigma.assign(sigma_packed_upper, sigma_packed_upper + (size_t)r * (r + 1) / 2);
    else g.var.assign(var, var + r);
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_set_image_dispersion(jaicov_handle *h, int32_t image, int64_t n_rows, const double *sigma_packed_upper) {
    if (!h || image < 0 || n_rows < 0 || (n_rows > 0 && !sigma_packed_upper)) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_set_image_dispersion(s_, image, n_rows, sigma_packed_upper))
    API_GUARD_BEGIN
    for (size_t i = 0; i < h->img_sigma.size(); i++)
        if (h->img_sigma[i].img == image) { h->img_sigma.erase(h->img_sigma.begin() + i); break; }
    if (n_rows > 0) {
        h->img_sigma.emplace_back();
        h->img_sigma.back().img = image;
        h->img_sigma.back().sigma.assign(sigma_packed_upper, sigma_packed_upper + (size_t)n_rows * (n_rows + 1) / 2);
    }
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_set_datum(jaicov_handle *h, const int32_t free_flags[7], int32_t n_unknowns, int32_t n_observations) {
    if (!h || !free_flags || n_unknowns < 0) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_set_datum(s_, free_flags, n_unknowns, n_observations))
    for (int i = 0; i < 7; i++) h->free_flags[i] = free_flags[i] != 0;
    h->n_unknowns = n_unknowns;
    h->n_observations = n_observations;
    h->has_datum_call = true;
    h->prepared = false; h->resident = false;
    return JAICOV_OK;
}

int32_t jaicov_set_reduced_rows(jaicov_handle *h, int32_t num_rows) {
    if (!h || num_rows < 0) return JAICOV_ILLEGAL_ARGUMENT;
    GROUP_FORWARD(h, jaicov_set_reduced_rows(s_, num_rows))
    h->reduced_rows = num_rows;
    return JAICOV_OK;
}

int32_t jaicov_iterate(jaicov_handle *h, int32_t final_pass, int32_t apply_update) {
    if (!h) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {
        API_GUARD_BEGIN
        group_ensure(h);
        return group_run(h, [&](jaicov_handle *sh, int) { return (int)jaicov_iterate(sh, final_pass, apply_update); });
        API_GUARD_END(h)
    }
    API_GUARD_BEGIN
    prepare(h);
    JCHECK(cudaSetDevice(h->opt.device));
    PassResult r = run_pass(h, final_pass != 0, apply_update != 0);
    h->stats.max_abs_dx = r.max_abs_dx;
    if (final_pass) h->stats.omega = r.omega;
    h->stats.iterations++;
    if (r.info != 0 || r.bad) { h->stats.status = JAICOV_SINGULAR_MATRIX; return JAICOV_SINGULAR_MATRIX; }
    h->stats.status = JAICOV_OK;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_estimate(jaicov_handle *h, jaicov_progress_cb cb, void *user, volatile int32_t *interrupt_flag) {
    if (!h) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {
        API_GUARD_BEGIN
        group_ensure(h);
        // the listener and the caller's interrupt flag belong to device 0's handle, which runs on the calling thread; the other
        // devices learn of an interrupt through the agreement step of the loop (agree_interrupt)
        std::vector<int32_t> dummy(h->sub.size(), 0);
        return group_run(h, [&](jaicov_handle *sh, int i) {
            return (int)jaicov_estimate(sh, i == 0 ? cb : nullptr, i == 0 ? user : nullptr, i == 0 || !interrupt_flag ? interrupt_flag : &dummy[i]);
        });
        API_GUARD_END(h)
    }
    API_GUARD_BEGIN
    const double SQRT_EPS = std::sqrt(kEps);                 // BA:77
    auto fire = [&](int st, double a, double b) { if (cb) cb(user, st, a, b); };
    fire(JAICOV_STATE_BUSY, 0, 1);
    const int maxIter = h->opt.max_iterations;
    int runs = maxIter - 1;                                  // BA:213
    bool isEstimated = false, complete = false, isConverge = true;
    if (maxIter == 0) complete = isEstimated = true;         // BA:216-219
    h->derive_first_damping = h->opt.damping_value > 0;      // BA:207-208
    h->adapted_damping = 0.0;
    h->lm_omega = 0.0;
    h->last_valid_max_abs_dx = 0.0;
    h->prepared = false; h->resident = false;                 // (re)upload the caller's values
    if (h->opt.use_centroid && !centroid_shift(h, false))
        return fail(h, JAICOV_ILLEGAL_ARGUMENT, "numbers of coordinate components are un-equal or zero (BA:151)");
    prepare(h);
    JCHECK(cudaSetDevice(h->opt.device));
    upload_values(h);
    h->stats = jaicov_stats();
    int status = JAICOV_OK;
    do {
        h->stats.iteration_step = maxIter - runs;            // BA:230
        fire(JAICOV_STATE_ITERATE, maxIter, h->stats.iteration_step);
        if (interrupted(h, interrupt_flag)) { status = JAICOV_INTERRUPT; break; }   // BA:240-245
        complete = isEstimated;                              // BA:250
        if (complete && h->opt.invert_mode != JAICOV_INVERT_NONE) fire(JAICOV_STATE_INVERT_NORMAL_EQUATION_MATRIX, 0, 1);
        PassResult r = run_pass(h, complete, true);
        h->stats.iterations++;
        if (r.info != 0) { status = JAICOV_SINGULAR_MATRIX; break; }                                        // BA:304-309
        if (r.lm_step) fire(JAICOV_STATE_LEVENBERG_MARQUARDT_STEP, r.lm_last, h->adapted_damping);       // BA:417-418
        if (complete) {
            fire(JAICOV_STATE_ESTIMATE_STOCHASTIC_PARAMETERS, 0, 1);
            if (!r.lm_step || r.lm_accepted) { h->stats.omega = r.omega; h->lm_omega = r.omega; }
        }
        h->stats.max_abs_dx = r.max_abs_dx;
        if (interrupted(h, interrupt_flag)) { status = JAICOV_INTERRUPT; break; }   // BA:320-325
        if (r.bad || std::isnan(r.max_abs_dx) || std::isinf(r.max_abs_dx)) { status = JAICOV_SINGULAR_MATRIX; break; }  // BA:327-330
        else if (r.max_abs_dx <= SQRT_EPS && runs > 0 && h->adapted_damping == 0) {     // BA:332-337
            isEstimated = true;
            fire(JAICOV_STATE_CONVERGENCE, SQRT_EPS, r.max_abs_dx);
        } else if (runs-- <= 1) {                            // BA:338-345
            if (complete) isConverge = false;
            isEstimated = true;
        } else {
            fire(JAICOV_STATE_CONVERGENCE, SQRT_EPS, r.max_abs_dx);
        }
        if (isEstimated || h->adapted_damping <= SQRT_EPS || runs < maxIter * 0.5 + 1) h->adapted_damping = 0.0;   // BA:352-353
    } while (!complete);
    download_values(h);
    if (status == JAICOV_OK) {
        if (h->opt.use_centroid) centroid_shift(h, true);    // BA:357-358
        status = isConverge ? JAICOV_ERROR_FREE_ESTIMATION : JAICOV_NO_CONVERGENCE;   // BA:377-384
    }
    h->prepared = false;   // device values are centred; a later call starts from the host copies
    h->resident = (status == JAICOV_ERROR_FREE_ESTIMATION || status == JAICOV_NO_CONVERGENCE);
    h->stats.status = status;
    return status;
    API_GUARD_END(h)
}

int32_t jaicov_get_stats(jaicov_handle *h, jaicov_stats *out) {
    if (!h || !out) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {      // device 0's view; stage times: the slowest device
        int32_t rc = jaicov_get_stats(h->sub[0], out);
        for (size_t i = 1; i < h->sub.size() && rc == JAICOV_OK; i++) {
            jaicov_stats t;
            rc = jaicov_get_stats(h->sub[i], &t);
            out->ms_assembly = std::max(out->ms_assembly, t.ms_assembly); out->ms_factor = std::max(out->ms_factor, t.ms_factor);
            out->ms_solve = std::max(out->ms_solve, t.ms_solve); out->ms_inverse = std::max(out->ms_inverse, t.ms_inverse);
            out->ms_omega = std::max(out->ms_omega, t.ms_omega); out->ms_total = std::max(out->ms_total, t.ms_total);
        }
        return rc;
    }
    int d = 0;
    for (int i = 0; i < 7; i++) d += h->free_flags[i];
    h->stats.n_unknowns = h->n_unknowns;
    h->stats.n_datum = d;
    h->stats.n_observations = h->n_observations;
    h->stats.dof = h->n_observations - h->n_unknowns + d;      // BA:1080-1082
    const double s2 = h->opt.sigma2apriori > 0 ? h->opt.sigma2apriori : 1.0;
    h->stats.sigma2apriori = s2;
    // getVarianceFactorAposteriori, BA:1090-1093
    h->stats.sigma2aposteriori = (h->stats.dof > 0 && h->stats.omega > 0 && h->opt.estimation_type != JAICOV_SIMULATION &&
                                  h->opt.apply_aposteriori)
                                     ? std::fabs(h->stats.omega / (double)h->stats.dof)
                                     : s2;
    *out = h->stats;
    return JAICOV_OK;
}

int32_t jaicov_get_values(jaicov_handle *h, double *xyz, double *io_val, double *coef_val, double *eo_val) {
    if (!h) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) return jaicov_get_values(h->sub[0], xyz, io_val, coef_val, eo_val);     // identical on every device
    API_GUARD_BEGIN
    if (h->prepared) { JCHECK(cudaSetDevice(h->opt.device)); download_values(h); }
    if (xyz && !h->xyz.empty()) memcpy(xyz, h->xyz.data(), h->xyz.size() * sizeof(double));
    if (io_val && !h->io_val.empty()) memcpy(io_val, h->io_val.data(), h->io_val.size() * sizeof(double));
    if (coef_val && !h->coef_val.empty()) memcpy(coef_val, h->coef_val.data(), h->coef_val.size() * sizeof(double));
    if (eo_val && !h->eo_val.empty()) memcpy(eo_val, h->eo_val.data(), h->eo_val.size() * sizeof(double));
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_get_dx(jaicov_handle *h, double *dx) {
    if (is_group(h)) return jaicov_get_dx(h->sub[0], dx);
    if (!h || !dx || !h->dxref.p) return JAICOV_ILLEGAL_ARGUMENT;
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    JCHECK(cudaMemcpy(dx, h->dxref.p, (size_t)(h->P.u + h->P.d) * sizeof(double), cudaMemcpyDeviceToHost));
    return JAICOV_OK;
    API_GUARD_END(h)
}

static int32_t pack_to_host(jaicov_handle *h, const double *border, const double *q11, double *dst, int64_t n_limit = -1,
                            const double *lower = nullptr) {
    if (!lower) lower = h->M.p;
    NvtxRange nvtx_("jaicov: packed Qxx -> host");
    const DevProblem &P = h->P;
    const int64_t n = n_limit >= 0 ? n_limit : (int64_t)P.u + P.d;   // the packed leading block is a prefix of the packed matrix
    const int64_t budget = (int64_t)1 << 25;   // doubles per staging chunk (256 MiB)
    DevBuf<double> stage[2];
    int64_t c0 = 0;
    int which = 0;
    cudaStream_t s = h->stream;
    stage[0].alloc((size_t)std::min<int64_t>(budget + n, n * (n + 1) / 2 + 1));
    stage[1].alloc(stage[0].n);
    while (c0 < n) {
        int64_t c1 = c0;
        int64_t cnt = 0;
        while (c1 < n && (cnt == 0 || cnt + c1 + 1 <= (int64_t)stage[0].n)) { cnt += c1 + 1; c1++; }
        launch_pack_columns(lower, P.np, border, P.np, q11, P.d, c0, c1, stage[which].p, s);
        JCHECK(cudaMemcpyAsync(dst + c0 * (c0 + 1) / 2, stage[which].p, cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
        which ^= 1;
        c0 = c1;
        if (which == 0) JCHECK(cudaStreamSynchronize(s));   // both staging buffers in flight at most
    }
    JCHECK(cudaStreamSynchronize(s));
    return JAICOV_OK;
}

// Single-process multi-GPU handle: the column tiles of Qxx live on their owners.  They are gathered over NVLink (peer-to-peer 2-D
// copies, one per tile) into device 0's system buffer -- whose factor is no longer needed once the pass is over -- in the
// single-GPU layout (row-major lower triangle), and leave through ONE packed DMA stream like on a single GPU.
static int32_t group_gather_qxx(jaicov_handle *g) {
    jaicov_handle *h0 = g->sub[0];
    if (h0->gathered_qxx) return JAICOV_OK;
    const int64_t np = h0->P.np;
    JCHECK(cudaSetDevice(h0->opt.device));
    double *dstM = h0->M.p;
    if (h0->owner_only) {            // device 0 holds only its own panels: the gathered matrix gets a buffer of its own
        if (h0->gather.n < (size_t)np * np) h0->gather.alloc((size_t)np * np);
        dstM = h0->gather.p;
    }
    for (size_t r = 1; r < g->sub.size(); r++) {
        int can = 0;
        JCHECK(cudaDeviceCanAccessPeer(&can, h0->opt.device, g->sub[r]->opt.device));
        if (can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->sub[r]->opt.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) JCHECK(e);
            cudaGetLastError();
        }   // without peer access the copies below are staged through the host by the driver: slower, still correct
    }
    for (jaicov_handle *sh : g->sub) {          // everybody's pass has finished (group_run joined its threads); drain the streams
        JCHECK(cudaSetDevice(sh->opt.device));
        JCHECK(cudaStreamSynchronize(sh->stream));
    }
    JCHECK(cudaSetDevice(h0->opt.device));
    for (jaicov_handle *sh : g->sub) {
        const int64_t ldx = (int64_t)sh->ktab.size() * kBlk;
        for (size_t jl = 0; jl < sh->ktab.size(); jl++) {
            const int64_t e0 = sh->ktab[jl];
            JCHECK(cudaMemcpy2DAsync(dstM + e0 * np + e0, (size_t)np * sizeof(double), sh->Xl.p + e0 * ldx + (int64_t)jl * kBlk,
                                     (size_t)ldx * sizeof(double), kBlk * sizeof(double), (size_t)(np - e0), cudaMemcpyDeviceToDevice,
                                     h0->stream));
        }
    }
    JCHECK(cudaStreamSynchronize(h0->stream));
    h0->gathered_qxx = true;
    return JAICOV_OK;
}

int32_t jaicov_get_qxx_packed(jaicov_handle *h, double *dst) {
    if (!h || !dst) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {
        jaicov_handle *h0 = h->sub[0];
        for (jaicov_handle *sh : h->sub)
            if (!sh->have_qxx) return fail(h, JAICOV_NOT_INITIALISED, "no cofactor matrix: run a final pass with invert_mode FULL");
        API_GUARD_BEGIN
        group_gather_qxx(h);
        return pack_to_host(h0, h0->Tq.p, h0->small.p + 49, dst, h0->qxx_rows(), h0->owner_only ? h0->gather.p : h0->M.p);
        API_GUARD_END(h)
    }
    if (!h->have_qxx) return fail(h, JAICOV_NOT_INITIALISED, "no cofactor matrix: run a final pass with invert_mode FULL");
    if (h->dist_on) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "distributed handle: Qxx is spread over the ranks, use jaicov_get_qxx_block (partial sums) or jaicov_get_qxx_local");
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    return pack_to_host(h, h->Tq.p, h->small.p + 49, dst, h->qxx_rows());
    API_GUARD_END(h)
}

int32_t jaicov_get_qxx_block(jaicov_handle *h, int32_t r0, int32_t r1, int32_t c0, int32_t c1, double *dst, int64_t ld) {
    if (!h || !dst) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h) && (r1 < r0 || c1 < c0 || ld < c1 - c0)) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {      // the devices hold disjoint parts: the block is the sum of their partial blocks
        API_GUARD_BEGIN
        const int64_t rows = r1 - r0, cols = c1 - c0;
        std::vector<double> part((size_t)std::max<int64_t>(rows * cols, 1));
        for (size_t i = 0; i < h->sub.size(); i++) {
            const int32_t rc = jaicov_get_qxx_block(h->sub[i], r0, r1, c0, c1, i == 0 ? dst : part.data(), i == 0 ? ld : cols);
            if (rc != JAICOV_OK) { h->err = h->sub[i]->err; return rc; }
            if (i > 0)
                for (int64_t r = 0; r < rows; r++)
                    for (int64_t c = 0; c < cols; c++) dst[r * ld + c] += part[(size_t)(r * cols + c)];
        }
        return JAICOV_OK;
        API_GUARD_END(h)
    }
    if (!h->have_qxx) return fail(h, JAICOV_NOT_INITIALISED, "no cofactor matrix: run a final pass with invert_mode FULL");
    const int n = h->qxx_rows();
    if (r0 < 0 || c0 < 0 || r1 > n || c1 > n || r1 < r0 || c1 < c0 || ld < c1 - c0) return JAICOV_ILLEGAL_ARGUMENT;
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    DevBuf<double> tmp;
    const int64_t rows_per = std::max<int64_t>(1, ((int64_t)1 << 25) / std::max(1, c1 - c0));
    tmp.alloc((size_t)std::min<int64_t>(rows_per, std::max(1, r1 - r0)) * std::max(1, c1 - c0));
    for (int64_t r = r0; r < r1; r += rows_per) {
        const int rr1 = (int)std::min<int64_t>(r1, r + rows_per);
        if (h->dist_on)
            launch_get_block_dist(h->Xl.p, (int64_t)h->ktab.size() * kBlk, h->d_col_local.p, h->Tq.p, h->P.np, h->small.p + 49, h->P.d,
                                  h->dist.rank, (int)r, rr1, c0, c1, tmp.p, h->stream);
        else
            launch_get_block(h->M.p, h->P.np, h->Tq.p, h->P.np, h->small.p + 49, h->P.d, (int)r, rr1, c0, c1, tmp.p, h->stream);
        JCHECK(cudaMemcpy2DAsync(dst + (r - r0) * ld, ld * sizeof(double), tmp.p, (size_t)(c1 - c0) * sizeof(double),
                                 (size_t)(c1 - c0) * sizeof(double), rr1 - r, cudaMemcpyDeviceToHost, h->stream));
        JCHECK(cudaStreamSynchronize(h->stream));
    }
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_get_qxx_submatrix(jaicov_handle *h, int32_t n_idx, const int32_t *idx, double scale, double *dst) {
    if (!h || n_idx < 0 || (n_idx > 0 && (!idx || !dst))) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {
        API_GUARD_BEGIN
        std::vector<double> part((size_t)n_idx * n_idx);
        for (size_t i = 0; i < h->sub.size(); i++) {
            const int32_t rc = jaicov_get_qxx_submatrix(h->sub[i], n_idx, idx, scale, i == 0 ? dst : part.data());
            if (rc != JAICOV_OK) { h->err = h->sub[i]->err; return rc; }
            if (i > 0)
                for (size_t k = 0; k < part.size(); k++) dst[k] += part[k];
        }
        return JAICOV_OK;
        API_GUARD_END(h)
    }
    if (!h->have_qxx) return fail(h, JAICOV_NOT_INITIALISED, "no cofactor matrix: run a final pass with invert_mode FULL");
    const int n = h->qxx_rows();
    for (int i = 0; i < n_idx; i++)
        if (idx[i] < 0 || idx[i] >= n) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "index outside the cofactor matrix");
    if (n_idx == 0) return JAICOV_OK;
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    DevBuf<int32_t> didx;
    didx.upload(std::vector<int32_t>(idx, idx + n_idx));
    const int64_t rows_per = std::max<int64_t>(1, ((int64_t)1 << 25) / n_idx);
    DevBuf<double> tmp;
    tmp.alloc((size_t)std::min<int64_t>(rows_per, n_idx) * n_idx);
    for (int64_t r = 0; r < n_idx; r += rows_per) {
        const int nr = (int)std::min<int64_t>(rows_per, n_idx - r);
        // rows r .. r+nr of the gathered matrix: entry (i, j) = scale * Qxx[idx[r + i], idx[j]]
        launch_get_submatrix(h->dist_on ? nullptr : h->M.p, h->P.np, h->dist_on ? h->Xl.p : nullptr, (int64_t)h->ktab.size() * kBlk,
                             h->dist_on ? h->d_col_local.p : nullptr, h->Tq.p, h->P.np, h->small.p + 49, h->P.d,
                             h->dist_on ? h->dist.rank : 0, didx.p + r, nr, didx.p, n_idx, scale, tmp.p, h->stream);
        JCHECK(cudaMemcpyAsync(dst + r * n_idx, tmp.p, (size_t)nr * n_idx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        JCHECK(cudaStreamSynchronize(h->stream));
    }
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_get_qxx_diag(jaicov_handle *h, double *dst) {
    if (!h || !dst) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "several devices: use jaicov_get_qxx_submatrix / jaicov_get_qxx_packed");
    if (!h->have_qxx) return fail(h, JAICOV_NOT_INITIALISED, "no cofactor matrix: run a final pass with invert_mode FULL");
    if (h->dist_on) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "distributed handle: use jaicov_get_qxx_block");
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    const DevProblem &P = h->P;
    std::vector<double> small(128);
    JCHECK(cudaMemcpy(small.data(), h->small.p, 128 * sizeof(double), cudaMemcpyDeviceToHost));
    const int nq = h->qxx_rows();
    for (int a = 0; a < std::min(P.d, nq); a++) dst[a] = small[49 + a * kMaxDatum + a];
    if (nq > P.d)
        JCHECK(cudaMemcpy2D(dst + P.d, sizeof(double), h->M.p, (size_t)(P.np + 1) * sizeof(double), sizeof(double), nq - P.d,
                            cudaMemcpyDeviceToHost));
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_get_qxx_local(jaicov_handle *h, int32_t *n_tiles, int32_t *tile_first_col, int32_t tile_cap, double *dst) {
    if (!h || !n_tiles) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "single-process handle: every getter returns complete results, use jaicov_get_qxx_packed / _block");
    if (!h->dist_on) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "not a distributed handle");
    API_GUARD_BEGIN
    *n_tiles = (int32_t)h->ktab.size();
    if (tile_first_col)
        for (int i = 0; i < std::min<int>(tile_cap, *n_tiles); i++) tile_first_col[i] = h->ktab[i] + h->P.d;
    if (dst) {
        if (!h->have_qxx) return fail(h, JAICOV_NOT_INITIALISED, "no cofactor matrix: run a final pass with invert_mode FULL");
        JCHECK(cudaSetDevice(h->opt.device));
        // tile after tile: gather the rows on and below the tile's diagonal into a contiguous staging block, ship it
        const int64_t np = h->P.np, ldx = (int64_t)h->ktab.size() * kBlk;
        const size_t cap = (size_t)1 << 26;   // doubles per staging buffer (512 MiB)
        DevBuf<double> stage[2];
        stage[0].alloc(cap); stage[1].alloc(cap);
        int which = 0;
        size_t fill = 0, out_off = 0;
        cudaStream_t s = h->stream;
        cudaEvent_t done[2];
        JCHECK(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
        JCHECK(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
        bool pending[2] = {false, false};
        auto flush = [&]() {
            if (!fill) return;
            JCHECK(cudaMemcpyAsync(dst + out_off, stage[which].p, fill * sizeof(double), cudaMemcpyDeviceToHost, s));
            JCHECK(cudaEventRecord(done[which], s));
            pending[which] = true;
            out_off += fill;
            fill = 0;
            which ^= 1;
            if (pending[which]) { JCHECK(cudaEventSynchronize(done[which])); pending[which] = false; }
        };
        for (size_t jl = 0; jl < h->ktab.size(); jl++) {
            int64_t r0 = h->ktab[jl];
            while (r0 < np) {
                const int64_t rows = std::min<int64_t>(np - r0, (int64_t)((cap - fill) / kBlk));
                if (rows <= 0) { flush(); continue; }
                launch_copy2d(stage[which].p + fill, kBlk, h->Xl.p + r0 * ldx + (int64_t)jl * kBlk, ldx, rows, kBlk, s);
                fill += (size_t)rows * kBlk;
                r0 += rows;
                if (fill + kBlk > cap) flush();
            }
        }
        flush();
        JCHECK(cudaStreamSynchronize(s));
        cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
    }
    return JAICOV_OK;
    API_GUARD_END(h)
}

// ---- covariance propagation of transformed coordinates ------------------------------------------------------
namespace {

// R = Rx(omega) Ry(phi) Rz(kappa), the rotation of ExteriorOrientation as written out in
// CoordinateTransformationExteriorOrientation.java:172-184, and its derivatives with respect to the three angles
void rotation_with_derivatives(double om, double ph, double ka, double R[3][3], double dR[3][3][3]) {
    const double so = std::sin(om), co = std::cos(om), sp = std::sin(ph), cp = std::cos(ph), sk = std::sin(ka), ck = std::cos(ka);
    const double Rx[3][3] = {{1, 0, 0}, {0, co, -so}, {0, so, co}}, dRx[3][3] = {{0, 0, 0}, {0, -so, -co}, {0, co, -so}};
    const double Ry[3][3] = {{cp, 0, sp}, {0, 1, 0}, {-sp, 0, cp}}, dRy[3][3] = {{-sp, 0, cp}, {0, 0, 0}, {-cp, 0, -sp}};
    const double Rz[3][3] = {{ck, -sk, 0}, {sk, ck, 0}, {0, 0, 1}}, dRz[3][3] = {{-sk, -ck, 0}, {ck, -sk, 0}, {0, 0, 0}};
    auto mul3 = [](const double A[3][3], const double B[3][3], const double C[3][3], double out[3][3]) {
        double T[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) T[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) out[i][j] = T[i][0] * C[0][j] + T[i][1] * C[1][j] + T[i][2] * C[2][j];
    };
    mul3(Rx, Ry, Rz, R);
    mul3(dRx, Ry, Rz, dR[0]);
    mul3(Rx, dRy, Rz, dR[1]);
    mul3(Rx, Ry, dRz, dR[2]);
}

}  // namespace

int32_t jaicov_propagate_eo_transform(jaicov_handle *h, int32_t n_points, const int32_t *point, const int32_t *src_image,
                                      const int32_t *trg_image, double sigma2, double *xyz_out, double *cov_packed) {
    if (!h || n_points < 0 || (n_points > 0 && (!point || !src_image || !trg_image))) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {
        API_GUARD_BEGIN
        const size_t npk = (size_t)(3 * (int64_t)n_points) * (3 * (int64_t)n_points + 1) / 2;
        std::vector<double> part(cov_packed ? npk : 0);
        for (size_t i = 0; i < h->sub.size(); i++) {
            const int32_t rc = jaicov_propagate_eo_transform(h->sub[i], n_points, point, src_image, trg_image, sigma2, i == 0 ? xyz_out : nullptr,
                                                             cov_packed ? (i == 0 ? cov_packed : part.data()) : nullptr);
            if (rc != JAICOV_OK) { h->err = h->sub[i]->err; return rc; }
            if (i > 0 && cov_packed)
                for (size_t k = 0; k < npk; k++) cov_packed[k] += part[k];
            if (!cov_packed) break;
        }
        return JAICOV_OK;
        API_GUARD_END(h)
    }
    if (cov_packed && !h->have_qxx) return fail(h, JAICOV_NOT_INITIALISED, "no cofactor matrix: run a final pass with invert_mode FULL");
    const int nPt = (int)(h->xyz.size() / 3), nImg = (int)(h->eo_val.size() / 6);
    for (int i = 0; i < n_points; i++)
        if (point[i] < 0 || point[i] >= nPt || src_image[i] < 0 || src_image[i] >= nImg || trg_image[i] < 0 || trg_image[i] >= nImg)
            return fail(h, JAICOV_ILLEGAL_ARGUMENT, "point or image index out of range");
    if (n_points == 0) return JAICOV_OK;
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    if (h->prepared) download_values(h);       // jaicov_iterate users: the host copies follow the device values
    const int nq = h->qxx_rows();
    // X_trg = X0_trg + R_trg R_src' (X - X0_src)  (:209-215) and its Jacobian with respect to
    // [X0_trg, angles_trg, X0_src, angles_src, X] (:223-279), on the host copies of the adjusted values
    std::vector<double> Jv((size_t)n_points * 45, 0.0);
    std::vector<int32_t> Jc((size_t)n_points * 15, -1);
    auto col_or_skip = [&](int32_t c) { return (active(c) && c < nq) ? c : -1; };
    for (int i = 0; i < n_points; i++) {
        const double *X = &h->xyz[3 * (size_t)point[i]];
        double *J = &Jv[(size_t)i * 45];
        int32_t *C = &Jc[(size_t)i * 15];
        for (int k = 0; k < 3; k++) C[12 + k] = col_or_skip(h->pt_col[3 * (size_t)point[i] + k]);
        if (src_image[i] == trg_image[i]) {     // the reference image itself: identity (:141-149)
            for (int k = 0; k < 3; k++) {
                J[k * 15 + 12 + k] = 1.0;
                if (xyz_out) xyz_out[3 * (size_t)i + k] = X[k];
            }
            continue;
        }
        const double *eT = &h->eo_val[6 * (size_t)trg_image[i]], *eS = &h->eo_val[6 * (size_t)src_image[i]];
        for (int k = 0; k < 6; k++) {
            C[k] = col_or_skip(h->eo_col[6 * (size_t)trg_image[i] + k]);
            C[6 + k] = col_or_skip(h->eo_col[6 * (size_t)src_image[i] + k]);
        }
        double RT[3][3], dRT[3][3][3], RS[3][3], dRS[3][3][3];
        rotation_with_derivatives(eT[3], eT[4], eT[5], RT, dRT);
        rotation_with_derivatives(eS[3], eS[4], eS[5], RS, dRS);
        const double dv[3] = {X[0] - eS[0], X[1] - eS[1], X[2] - eS[2]};
        double q[3], M3[3][3];
        for (int a = 0; a < 3; a++) q[a] = RS[0][a] * dv[0] + RS[1][a] * dv[1] + RS[2][a] * dv[2];        // R_src' d
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) M3[r][c] = RT[r][0] * RS[c][0] + RT[r][1] * RS[c][1] + RT[r][2] * RS[c][2];   // R_trg R_src'
        for (int r = 0; r < 3; r++) {
            if (xyz_out) xyz_out[3 * (size_t)i + r] = eT[r] + RT[r][0] * q[0] + RT[r][1] * q[1] + RT[r][2] * q[2];
            J[r * 15 + r] = 1.0;                                                                               // d / d X0_trg
            for (int a = 0; a < 3; a++) {
                J[r * 15 + 3 + a] = dRT[a][r][0] * q[0] + dRT[a][r][1] * q[1] + dRT[a][r][2] * q[2];         // d / d angle_trg
                J[r * 15 + 6 + a] = -M3[r][a];                                                                 // d / d X0_src
                double s = 0.0;                                                                                // d / d angle_src
                for (int k = 0; k < 3; k++) {
                    const double qa = dRS[a][0][k] * dv[0] + dRS[a][1][k] * dv[1] + dRS[a][2][k] * dv[2];    // (dR_src' d)_k
                    s += RT[r][k] * qa;
                }
                J[r * 15 + 9 + a] = s;
                J[r * 15 + 12 + a] = M3[r][a];                                                                 // d / d X
            }
        }
    }
    if (!cov_packed) return JAICOV_OK;
    DevBuf<double> dJ, dout;
    DevBuf<int32_t> dC;
    dJ.upload(Jv);
    dC.upload(Jc);
    const int64_t n3 = 3 * (int64_t)n_points, npk = n3 * (n3 + 1) / 2;
    dout.alloc((size_t)npk);
    launch_propagate(h->dist_on ? nullptr : h->M.p, h->P.np, h->dist_on ? h->Xl.p : nullptr, (int64_t)h->ktab.size() * kBlk,
                     h->dist_on ? h->d_col_local.p : nullptr, h->Tq.p, h->P.np, h->small.p + 49, h->P.d,
                     h->dist_on ? h->dist.rank : 0, n_points, dJ.p, dC.p, sigma2, dout.p, h->stream);
    JCHECK(cudaGetLastError());
    JCHECK(cudaMemcpyAsync(cov_packed, dout.p, (size_t)npk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JCHECK(cudaStreamSynchronize(h->stream));
    return JAICOV_OK;
    API_GUARD_END(h)
}

// ---- batched direct linear transformation -------------------------------------------------------------------------
int32_t jaicov_dlt_batch(int32_t device, int32_t n_img, const int64_t *pt_ptr, const double *xy, const double *xyz, const double *io,
                         int32_t n_restrictions, const int32_t *restrictions, int32_t max_iterations, double *out20,
                         int32_t *status, int32_t *passes) {
    if (n_img < 0 || n_restrictions < 0 || max_iterations < 0 || (n_img > 0 && (!pt_ptr || !io || !out20 || !status))) return JAICOV_ILLEGAL_ARGUMENT;
    if (n_restrictions > 0 && !restrictions) return JAICOV_ILLEGAL_ARGUMENT;
    if (n_img == 0) return JAICOV_OK;
    if (pt_ptr[0] != 0) return JAICOV_ILLEGAL_ARGUMENT;
    for (int i = 0; i < n_img; i++)
        if (pt_ptr[i + 1] < pt_ptr[i]) return JAICOV_ILLEGAL_ARGUMENT;
    const int64_t m = pt_ptr[n_img];
    if (m > 0 && (!xy || !xyz)) return JAICOV_ILLEGAL_ARGUMENT;
    // validateRestrictions (DLT:268-277): distinct, insertion order; the two fixed principal distances make the identical one redundant
    std::vector<int32_t> restr;
    bool has[6] = {false, false, false, false, false, false};
    for (int i = 0; i < n_restrictions; i++) {
        const int32_t r = restrictions[i];
        if (r < 0 || r > 5) return JAICOV_ILLEGAL_ARGUMENT;
        if (!has[r]) { has[r] = true; restr.push_back(r); }
    }
    if (has[JAICOV_DLT_FIXED_PRINCIPLE_DISTANCE_X] && has[JAICOV_DLT_FIXED_PRINCIPLE_DISTANCE_Y] && has[JAICOV_DLT_IDENTICAL_PRINCIPLE_DISTANCE])
        restr.erase(std::find(restr.begin(), restr.end(), (int32_t)JAICOV_DLT_IDENTICAL_PRINCIPLE_DISTANCE));
    jaicov_handle *h = nullptr;
    API_GUARD_BEGIN
    if (usable_devices() == 0) throw CudaError{cudaErrorNoDevice, "no sm_100 device: jaicov_b200 has no CPU path", __FILE__, __LINE__};
    JCHECK(cudaSetDevice(device));
    DevBuf<int64_t> d_ptr;
    DevBuf<double> d_xy, d_xyz, d_io, d_out;
    DevBuf<int32_t> d_restr, d_status, d_passes;
    d_ptr.upload(std::vector<int64_t>(pt_ptr, pt_ptr + n_img + 1));
    d_xy.upload(std::vector<double>(xy, xy + 2 * m));
    d_xyz.upload(std::vector<double>(xyz, xyz + 3 * m));
    d_io.upload(std::vector<double>(io, io + 3 * (size_t)n_img));
    std::vector<int32_t> rr = restr;
    if (rr.empty()) rr.push_back(0);
    d_restr.upload(rr);
    d_out.alloc(20 * (size_t)n_img); d_status.alloc(n_img); d_passes.alloc(n_img);
    launch_dlt_batch(n_img, d_ptr.p, d_xy.p, d_xyz.p, d_io.p, (int)restr.size(), d_restr.p, max_iterations, d_out.p, d_status.p, d_passes.p, nullptr);
    JCHECK(cudaGetLastError());
    JCHECK(cudaMemcpy(out20, d_out.p, 20 * (size_t)n_img * sizeof(double), cudaMemcpyDeviceToHost));
    JCHECK(cudaMemcpy(status, d_status.p, (size_t)n_img * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (passes) JCHECK(cudaMemcpy(passes, d_passes.p, (size_t)n_img * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_eval_residual_jacobian(jaicov_handle *h, int32_t ns_max, double *a, double *w, double *p) {
    if (!h || !a || !w || !p) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {      // no collective in this stage: the first device evaluates all image points
        const int32_t rc = jaicov_eval_residual_jacobian(h->sub[0], ns_max, a, w, p);
        if (rc != JAICOV_OK) h->err = h->sub[0]->err;
        return rc;
    }
    API_GUARD_BEGIN
    prepare(h);
    JCHECK(cudaSetDevice(h->opt.device));
    const DevProblem &P = h->P;
    int maxcoef = 0;
    for (int c = 0; c < P.nCam; c++) maxcoef = std::max(maxcoef, h->coef_ptr[c + 1] - h->coef_ptr[c]);
    if (ns_max < 12 + maxcoef) return fail(h, JAICOV_ILLEGAL_ARGUMENT, "ns_max too small");
    DevBuf<double> da, dw, dp;
    da.alloc((size_t)P.m * 2 * ns_max); dw.alloc((size_t)P.m * 2); dp.alloc((size_t)P.m * 3);
    launch_pose(P, h->stream);
    launch_eval_k1(P, ns_max, da.p, dw.p, dp.p, h->stream);
    JCHECK(cudaGetLastError());
    JCHECK(cudaStreamSynchronize(h->stream));
    if (P.m) {
        JCHECK(cudaMemcpy(a, da.p, da.n * sizeof(double), cudaMemcpyDeviceToHost));
        JCHECK(cudaMemcpy(w, dw.p, dw.n * sizeof(double), cudaMemcpyDeviceToHost));
        JCHECK(cudaMemcpy(p, dp.p, dp.n * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_get_normal_equations(jaicov_handle *h, double *n_packed, double *rhs) {
    if (!h) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {      // the assembly all-reduces: every device takes part, the first one delivers
        API_GUARD_BEGIN
        group_ensure(h);
        return group_run(h, [&](jaicov_handle *sh, int i) { return (int)jaicov_get_normal_equations(sh, i == 0 ? n_packed : nullptr, i == 0 ? rhs : nullptr); });
        API_GUARD_END(h)
    }
    API_GUARD_BEGIN
    prepare(h);
    JCHECK(cudaSetDevice(h->opt.device));
    if (h->owner_only && n_packed)
        return fail(h, JAICOV_ILLEGAL_ARGUMENT, "the distributed dense route stores only this rank's panels of N (JAICOV_DIST_STORAGE=replica keeps a whole copy per rank)");
    assemble(h);
    JCHECK(cudaStreamSynchronize(h->stream));
    const DevProblem &P = h->P;
    if (n_packed) pack_to_host(h, h->Bt.p, nullptr, n_packed);
    if (rhs) {
        for (int a = 0; a < P.d; a++) rhs[a] = 0.0;
        if (P.u) JCHECK(cudaMemcpy(rhs + P.d, h->rhs.p, (size_t)P.u * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_omega(jaicov_handle *h, const double *dx, double *omega) {
    if (!h || !dx || !omega) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {
        API_GUARD_BEGIN
        group_ensure(h);
        std::vector<double> om(h->sub.size(), 0.0);
        const int rc = group_run(h, [&](jaicov_handle *sh, int i) { return (int)jaicov_omega(sh, dx, &om[i]); });
        *omega = om[0];
        return rc;
        API_GUARD_END(h)
    }
    API_GUARD_BEGIN
    prepare(h);
    JCHECK(cudaSetDevice(h->opt.device));
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    JCHECK(cudaMemcpy(h->dxref.p, dx, (size_t)(P.u + P.d) * sizeof(double), cudaMemcpyHostToDevice));
    launch_pose(P, s);
    launch_omega(P, h->S, h->dxref.p, h->omega_parts.p, s);
    if (h->obs_sharded) h->dist.allreduce_sum(h->omega_parts.p, 1, s);
    omega_dense_images(h);
    JCHECK(cudaMemsetAsync(h->omega_parts.p + 1, 0, sizeof(double), s));
    if (P.nBar) launch_omega_bars(P, h->dxref.p, h->omega_parts.p + 1, s);
    int gi = 0;
    for (Group &g : h->groups) {
        launch_group_w(g.r, g.tptr.p, g.d_obs.p, g.w.p, s);
        launch_group_omega(g.r, g.col.p, g.d_var.p, g.Pw.p, g.ldp, P.sigma2, g.w.p, h->dxref.p, h->omega_parts.p + 2 + gi, s);
        gi++;
    }
    JCHECK(cudaGetLastError());
    std::vector<double> om(2 + h->groups.size());
    JCHECK(cudaMemcpyAsync(om.data(), h->omega_parts.p, om.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    JCHECK(cudaStreamSynchronize(s));
    double t = 0.0;
    for (double v : om) t += v;
    *omega = t;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_normal_product(jaicov_handle *h, int32_t nvec, const double *x, double *y, double *rhs, double *wpw) {
    if (!h || nvec < 0 || (nvec > 0 && (!x || !y))) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {      // every device computes its image shard and receives the sum; the first one delivers
        API_GUARD_BEGIN
        group_ensure(h);
        const size_t n = h->sub[0]->n_unknowns >= 0 ? (size_t)h->sub[0]->n_unknowns + 7 : 0;
        std::vector<std::vector<double>> ys(h->sub.size()), rs(h->sub.size());
        std::vector<double> ws(h->sub.size(), 0.0);
        for (size_t i = 1; i < h->sub.size(); i++) { ys[i].resize(std::max<size_t>(1, (size_t)nvec * n)); rs[i].resize(std::max<size_t>(1, n)); }
        return group_run(h, [&](jaicov_handle *sh, int i) {
            return (int)jaicov_normal_product(sh, nvec, x, i == 0 ? y : ys[i].data(), rhs ? (i == 0 ? rhs : rs[i].data()) : nullptr,
                                              wpw ? (i == 0 ? wpw : &ws[i]) : nullptr);
        });
        API_GUARD_END(h)
    }
    API_GUARD_BEGIN
    if (!h->prepared && !h->resident) prepare(h);
    JCHECK(cudaSetDevice(h->opt.device));
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    const int64_t n = (int64_t)P.u + P.d;
    const size_t np = (size_t)P.np;
    DevBuf<double> dX, dY, dR, dW, dB;
    dX.alloc(std::max<size_t>(1, (size_t)nvec * n)); dY.alloc(std::max<size_t>(1, (size_t)nvec * n)); dR.alloc((size_t)n); dW.alloc(1);
    dB.alloc(8 * np);
    if (nvec) JCHECK(cudaMemcpyAsync(dX.p, x, (size_t)nvec * n * sizeof(double), cudaMemcpyHostToDevice, s));
    JCHECK(cudaMemsetAsync(dY.p, 0, dY.n * sizeof(double), s));
    JCHECK(cudaMemsetAsync(dR.p, 0, (size_t)n * sizeof(double), s));
    JCHECK(cudaMemsetAsync(dW.p, 0, sizeof(double), s));
    JCHECK(cudaMemsetAsync(dB.p, 0, 8 * np * sizeof(double), s));
    const bool want_rhs = rhs != nullptr || wpw != nullptr;
    launch_pose(P, s);
    launch_normal_product_points(P, nvec, n, dX.p, dY.p, want_rhs ? dR.p : nullptr, want_rhs ? dW.p : nullptr, s);
    if (h->obs_sharded) {      // image shards: sum over the ranks; everything below is replicated
        if (nvec) h->dist.allreduce_sum(dY.p, (size_t)nvec * n, s);
        if (want_rhs) { h->dist.allreduce_sum(dR.p, (size_t)n, s); h->dist.allreduce_sum(dW.p, 1, s); }
    }
    for (ImgSigma &is : h->img_sigma) {      // images with a fully populated dispersion (skipped by the kernel above)
        if (!is.rp) continue;
        launch_image_rows(P, is.img, h->ds_Ac.p, is.rp, s);
        for (int v = 0; v < nvec; v++)
            launch_dense_image_product(P, is.img, is.Pw.p, is.rp, is.rp, h->ds_Ac.p, dX.p + (size_t)v * n, 1, h->ds_v.p, h->ds_t.p,
                                       dY.p + (size_t)v * n, nullptr, s);
        if (want_rhs)
            launch_dense_image_product(P, is.img, is.Pw.p, is.rp, is.rp, h->ds_Ac.p, nullptr, 2, h->ds_v.p, h->ds_t.p, dR.p, dW.p, s);
    }
    if (P.d > 0) launch_datum_rows(P.xyz, P.pt_col, h->d_datum_pts.p, h->nDatumPts, h->free_mask, P.d, P.np, dB.p, s);
    launch_normal_product_rest(P, nvec, n, dX.p, dY.p, want_rhs ? dR.p : nullptr, want_rhs ? dW.p : nullptr, dB.p, P.np, s);
    for (Group &g : h->groups) {
        launch_group_w(g.r, g.tptr.p, g.d_obs.p, g.w.p, s);
        launch_group_product(g.r, g.col.p, g.d_var.p, g.Pw.p, g.ldp, P.sigma2, g.w.p, nvec, n, dX.p, dY.p, want_rhs ? dR.p : nullptr,
                             want_rhs ? dW.p : nullptr, s);
    }
    JCHECK(cudaGetLastError());
    if (nvec) JCHECK(cudaMemcpyAsync(y, dY.p, (size_t)nvec * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (rhs) JCHECK(cudaMemcpyAsync(rhs, dR.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (wpw) JCHECK(cudaMemcpyAsync(wpw, dW.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    JCHECK(cudaStreamSynchronize(s));
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_get_sweep_times(jaicov_handle *h, double *ms_by_image, double *ms_by_point, double *ms_omega) {
    if (!h) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) return jaicov_get_sweep_times(h->sub[0], ms_by_image, ms_by_point, ms_omega);
    if (!h->stream || !h->have_neq && !h->prepared && !h->resident) return fail(h, JAICOV_NOT_INITIALISED, "no pass has run");
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    JCHECK(cudaStreamSynchronize(h->stream));
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, h->evk[0], h->evk[1]);
    cudaEventElapsedTime(&b, h->evk[2], h->evk[3]);
    if (h->evk_omega) cudaEventElapsedTime(&c, h->evk[4], h->evk[5]);
    cudaGetLastError();
    if (ms_by_image) *ms_by_image = a;
    if (ms_by_point) *ms_by_point = b;
    if (ms_omega) *ms_omega = c;
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_get_device_bytes(jaicov_handle *h, int64_t out[4]) {
    if (!h || !out) return JAICOV_ILLEGAL_ARGUMENT;
    if (is_group(h)) {      // the largest device's figures
        out[0] = out[1] = out[2] = out[3] = 0;
        for (jaicov_handle *sh : h->sub) {
            int64_t t[4];
            const int32_t rc = jaicov_get_device_bytes(sh, t);
            if (rc != JAICOV_OK) return rc;
            for (int i = 0; i < 4; i++) out[i] = std::max(out[i], t[i]);
        }
        return JAICOV_OK;
    }
    out[0] = (int64_t)(h->M.n * sizeof(double));
    out[1] = (int64_t)(h->W.n * sizeof(double));
    out[2] = (int64_t)(h->Xl.n * sizeof(double));
    out[3] = h->dist_on ? (int64_t)(2 * h->dist.stage_elems * sizeof(double)) : 0;
    return JAICOV_OK;
}

int32_t jaicov_get_preconditioner(jaicov_handle *h, double *v) {
    if (is_group(h)) return jaicov_get_preconditioner(h->sub[0], v);
    if (!h || !v || !h->V.p) return JAICOV_ILLEGAL_ARGUMENT;
    API_GUARD_BEGIN
    JCHECK(cudaSetDevice(h->opt.device));
    for (int a = 0; a < h->P.d; a++) v[a] = 1.0;      // border diagonal is 0 <= EPS (BA:825-828)
    if (h->P.u) JCHECK(cudaMemcpy(v + h->P.d, h->V.p, (size_t)h->P.u * sizeof(double), cudaMemcpyDeviceToHost));
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_gemm_tiles(int32_t device, int32_t a_layout, int32_t b_layout, int32_t mt, int32_t nt, int64_t K, double alpha,
                          double beta, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
                          int32_t tri_out, int32_t kmode, int32_t reps, double *ms) {
    if (mt <= 0 || nt <= 0 || K <= 0 || K % kBlk || !A || !B || !C || kmode < K_FULL || kmode > K_MAX_IJ || (tri_out && mt != nt) ||
        (a_layout != 0 && a_layout != 1) || (b_layout != 0 && b_layout != 1))
        return JAICOV_ILLEGAL_ARGUMENT;
    const int64_t Mr = (int64_t)mt * kBlk, Nr = (int64_t)nt * kBlk;
    if (lda < (a_layout ? Mr : K) || ldb < (b_layout ? Nr : K) || ldc < Nr || (lda | ldb | ldc) % 2) return JAICOV_ILLEGAL_ARGUMENT;
    jaicov_handle *h = nullptr;
    API_GUARD_BEGIN
    if (usable_devices() == 0) return JAICOV_NOT_INITIALISED;
    JCHECK(cudaSetDevice(device));
    DevBuf<double> dA, dB, dC;
    const size_t na = (size_t)(a_layout ? K : Mr) * lda, nb = (size_t)(b_layout ? K : Nr) * ldb, nc = (size_t)Mr * ldc;
    dA.upload(A, na); dB.upload(B, nb); dC.upload(C, nc);
    cudaStream_t s;
    JCHECK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    GemmDesc g;
    g.al = a_layout; g.bl = b_layout; g.mt = mt; g.nt = nt; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = dA.p; g.lda = lda; g.B = dB.p; g.ldb = ldb; g.C = dC.p; g.ldc = ldc; g.tri_out = tri_out ? 1 : 0; g.kmode = kmode;
    const int n_rep = (beta != 0.0 || reps < 1) ? 1 : reps;
    JCHECK(cudaStreamSynchronize(0));
    cudaEventRecord(e0, s);
    for (int r = 0; r < n_rep; r++) launch_gemm(g, s);
    cudaEventRecord(e1, s);
    JCHECK(cudaGetLastError());
    JCHECK(cudaStreamSynchronize(s));
    float f = 0;
    cudaEventElapsedTime(&f, e0, e1);
    if (ms) *ms = f / n_rep;
    JCHECK(cudaMemcpy(C, dC.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    return JAICOV_OK;
    API_GUARD_END(h)
}

int32_t jaicov_spd_solve_invert(int32_t device, int64_t n, double *a, int32_t nrhs, double *b, int32_t invert, double *ms_factor,
                                double *ms_inverse) {
    if (n <= 0 || !a || nrhs < 0 || nrhs > kRhsRows || (nrhs > 0 && !b)) return JAICOV_ILLEGAL_ARGUMENT;
    jaicov_handle *h = nullptr;
    API_GUARD_BEGIN
    if (usable_devices() == 0) return JAICOV_NOT_INITIALISED;
    JCHECK(cudaSetDevice(device));
    const int64_t np = round_up(n, kBlk);
    DevBuf<double> M, W, Dinv, Rt;
    DevBuf<int> info;
    M.alloc((size_t)np * np); Dinv.alloc((size_t)np * kBlk); Rt.alloc((size_t)kRhsRows * np); info.alloc(1);
    if (invert) W.alloc((size_t)np * np);
    cudaStream_t s;
    JCHECK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    JCHECK(cudaMemsetAsync(M.p, 0, (size_t)np * np * sizeof(double), s));
    JCHECK(cudaMemcpy2DAsync(M.p, np * sizeof(double), a, n * sizeof(double), n * sizeof(double), n, cudaMemcpyHostToDevice, s));
    if (np > n) {   // identity padding
        std::vector<double> ones(np - n, 1.0);
        JCHECK(cudaMemcpy2DAsync(M.p + n * np + n, (np + 1) * sizeof(double), ones.data(), sizeof(double), sizeof(double), np - n,
                                 cudaMemcpyHostToDevice, s));
        JCHECK(cudaStreamSynchronize(s));
    }
    JCHECK(cudaMemsetAsync(Rt.p, 0, (size_t)kRhsRows * np * sizeof(double), s));
    if (nrhs) JCHECK(cudaMemcpy2DAsync(Rt.p, np * sizeof(double), b, n * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyHostToDevice, s));
    JCHECK(cudaMemsetAsync(info.p, 0, sizeof(int), s));
    CudaBackend be{s, info.p};
    DenseSchedule<CudaBackend> ds{be, M.p, np, np, Dinv.p};
    cudaEventRecord(e0, s);
    ds.potrf();
    cudaEventRecord(e1, s);
    if (nrhs > 0 && nrhs <= 8) launch_solve_rows8(M.p, np, Dinv.p, Rt.p, Rt.p + 8 * np, np, s);
    else if (nrhs) ds.solve_rows(Rt.p, np, 1);
    cudaEventRecord(e2, s);
    if (invert) { ds.invert_from_factor(W.p); launch_symmetrize(M.p, np, (int)n, s); }
    cudaEventRecord(e3, s);
    JCHECK(cudaGetLastError());
    int hinfo = 0;
    JCHECK(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (invert) JCHECK(cudaMemcpy2DAsync(a, n * sizeof(double), M.p, np * sizeof(double), n * sizeof(double), n, cudaMemcpyDeviceToHost, s));
    if (nrhs) JCHECK(cudaMemcpy2DAsync(b, n * sizeof(double), Rt.p, np * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToHost, s));
    JCHECK(cudaStreamSynchronize(s));
    float f1 = 0, f2 = 0;
    cudaEventElapsedTime(&f1, e0, e1);
    cudaEventElapsedTime(&f2, e2, e3);
    if (ms_factor) *ms_factor = f1;
    if (ms_inverse) *ms_inverse = f2;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    cudaStreamDestroy(s);
    return hinfo != 0 ? JAICOV_SINGULAR_MATRIX : JAICOV_OK;
    API_GUARD_END(h)
}

}  // extern "C"
