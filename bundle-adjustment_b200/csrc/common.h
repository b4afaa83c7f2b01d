// common.h -- shared declarations of the jaicov_b200 CUDA library (host side).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/jaicov_b200.h"

namespace jaicov {

constexpr int kBlk = 128;          // base block / tile edge of every dense kernel
constexpr int kMaxCoef = 64;       // coefficients per camera staged in shared memory
constexpr int kMaxDatum = 7;
constexpr int kRhsRows = 128;      // rows of the right-hand-side buffer (row 0 = n, rows 1..d = datum rows; 8 used by the loop)
constexpr double kEps = 1.1102230246251565e-16;  // Constant.EPS = 2^-53 (Constant.java:61-75)

extern std::atomic<long long> g_launch_count;   // kernels launched by this library (diagnostic, jaicov_launch_count); handles may live on several host threads

struct CudaError {
    cudaError_t code;
    const char *what;
    const char *file;
    int line;
};

#define JCHECK(expr)                                                          \
    do {                                                                      \
        cudaError_t e__ = (expr);                                             \
        if (e__ != cudaSuccess) throw ::jaicov::CudaError{e__, #expr, __FILE__, __LINE__}; \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: one flag per device ordinal, so that handles on several
// GPUs of one process (and on several host threads) each opt in on their own device
struct PerDeviceOnce {
    std::atomic<bool> done[64];
    PerDeviceOnce() { for (auto &d : done) d.store(false); }
    template <class F>
    void run(F &&f) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) { f(); return; }
        if (!done[dev].load(std::memory_order_acquire)) {
            f();            // idempotent: two threads racing here both set the same value
            done[dev].store(true, std::memory_order_release);
        }
    }
};

// NVTX range of one stage (nsys / ncu --nvtx timelines name the stages of a pass; header-only NVTX3, a no-op without a tool attached)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// Where the entries of the symmetric system matrix live on this rank.  tile_lcol == nullptr: the whole lower triangle, row-major,
// ld = np (one GPU, or the replicated layout).  Otherwise owner-only storage (distributed dense route): only the 128-column tiles
// this rank owns, side by side, full height; tile_lcol[t] = first LOCAL column of global tile t, -1 for tiles stored elsewhere
// (writes to those are dropped -- every rank evaluates all observations and keeps what it owns, no all-reduce of N).
struct SysView {
    double *p = nullptr;
    int64_t ld = 0;
    const int32_t *tile_lcol = nullptr;
};
#if defined(__CUDACC__)
// address of entry (a, b) / (b, a) of the lower triangle (internal indices), or nullptr if this rank does not store it
__device__ __forceinline__ double *sys_at(const SysView &v, int64_t a, int64_t b) {
    const int64_t r = a >= b ? a : b, c = a >= b ? b : a;
    if (!v.tile_lcol) return v.p + r * v.ld + c;
    const int32_t lc = v.tile_lcol[c >> 7];
    return lc < 0 ? nullptr : v.p + r * v.ld + lc + (c & 127);
}
__device__ __forceinline__ void sys_add(const SysView &v, int64_t a, int64_t b, double x) {
    double *q = sys_at(v, a, b);
    if (q) *q += x;
}
__device__ __forceinline__ void sys_set(const SysView &v, int64_t a, int64_t b, double x) {
    double *q = sys_at(v, a, b);
    if (q) *q = x;
}
#endif

// Flattened problem resident in HBM (structure-of-arrays; all pointers are device pointers).
struct DevProblem {
    // cameras
    int nCam = 0, nCoef = 0;
    double *io_val = nullptr;       // [3 nCam]
    int32_t *io_col = nullptr;      // [3 nCam]
    double *r0 = nullptr;           // [nCam]
    int32_t *coef_ptr = nullptr;    // [nCam+1]
    int32_t *coef_type = nullptr, *coef_order = nullptr, *coef_col = nullptr;  // [nCoef]
    double *coef_val = nullptr;     // [nCoef]
    const int32_t *cam_std = nullptr;  // [5 nCam] hasC, hasB, nB, nA, nD of a canonical coefficient list (hasC = -1: not canonical)
    double *coef_r0pow = nullptr;   // [nCoef] r0^(2 order) of the coefficient's camera (radial / distance polynomials)
    int32_t *zern_m = nullptr;      // [nCoef]
    int32_t *zern_ptr = nullptr;    // [nCoef+1]
    int32_t *zern_p = nullptr;
    double *zern_c = nullptr;
    int32_t *cam_kbase = nullptr;   // [nCam+1] start of camera's raw parameter block (x0,y0,c,coefs) in the camera-parameter list
    int32_t *campos_col = nullptr;  // [kRaw] column of every raw camera parameter (io then coefs, per camera)
    int kRaw = 0;                   // sum over cameras of (3 + ncoef)
    // images
    int nImg = 0;
    int32_t *cam_of_img = nullptr;
    double *eo_val = nullptr;       // [6 nImg]
    int32_t *eo_col = nullptr;      // [6 nImg]
    int64_t *pt_ptr = nullptr;      // [nImg+1]
    double *pose = nullptr;         // [16 nImg]
    // image points
    int64_t m = 0;
    int32_t *obj_idx = nullptr;
    double *xy = nullptr, *var = nullptr, *rho = nullptr;
    double *rw = nullptr;           // [3 m] r00, r01, r11 with P = R'R (weights are constant over the passes)
    int32_t *img_of_obs = nullptr;  // [m]
    const uint8_t *img_dense = nullptr;  // [nImg] or null: image has a fully populated dispersion (dense_sigma.cu), its points carry zero per-point weights
    int64_t *pt_obs_ptr = nullptr;  // [nPt+1] CSC: observations of every object point
    int64_t *pt_obs = nullptr;      // [m]
    // object points
    int nPt = 0;
    double *xyz = nullptr;          // [3 nPt]
    int32_t *pt_col = nullptr;      // [3 nPt]
    // scale bars
    int nBar = 0;
    int32_t *bar_a = nullptr, *bar_b = nullptr;
    double *bar_len = nullptr, *bar_var = nullptr;
    // image shard of this rank (multi-GPU: contiguous image range, hence contiguous observation range)
    int img0 = 0, img1 = 0;
    int64_t obs0 = 0, obs1 = 0;
    // system
    int d = 0;                      // datum defect (border size)
    int u = 0;                      // unknowns
    int64_t np = 0;                 // u padded to a multiple of kBlk (leading dimension of M)
    double sigma2 = 1.0;
};

struct WorkItem {
    int32_t img;
    int32_t pad;
    int64_t begin, end;
};

// ---- assembly.cu ---------------------------------------------------------------------------------------------------
struct AssemblyScratch {
    WorkItem *work = nullptr;       // image chunks
    int nWork = 0;
    int32_t *img_work_ptr = nullptr;  // [nImg+1] work items of every image
    int ntImg = 0;                  // 8-column tiles of the by-image Gram
    int ntPt = 0;                   // 8-column tiles of the by-point Gram
    double *img_partial = nullptr;  // [nWork][(8 ntImg)^2]
    double *cam_partial = nullptr;  // [nImg][kc*(kc+1)] camera block + rhs of every image (kc = max raw params per camera)
    double *cam_sum = nullptr;      // [nCam][kc*(kc+1)] sum over this rank's images (all-reduced across ranks)
    int kcMax = 0;
    int std_eval = 0;               // every camera has a canonical coefficient list: the sweeps use the straight-line evaluation
    double *pt_partial = nullptr;   // [nPt][3][8 ntPt]
    double *omega_partial = nullptr;  // [nWork + 2]
    double *dxp = nullptr;          // [3 nPt] dx of the object coordinates (0 where fixed), gathered per pass for Omega
};

void launch_pose(const DevProblem &P, cudaStream_t s);
void launch_obs_weights(const DevProblem &P, double *rw, cudaStream_t s);
// Camera group of the by-point sweep: cameras [cam0, cam1) with kraw raw parameters in total, Gram rows of 8 nt columns
struct PtGroup { int cam0, cam1, kraw, nt; };
// assembly of the image points in pieces, so that a multi-GPU run can all-reduce in between (api.cu: assemble):
//   launch_assemble_images: sweeps over this rank's images (unique EO blocks into M, per-camera sums);
//   launch_by_point:        per-point partial sums of one camera group;
//   launch_camera_scatter / launch_point_scatter: scatter of the (all-reduced) sums into M / rhs
void launch_assemble_images(const DevProblem &P, const AssemblyScratch &S, const SysView &M, double *rhs, cudaStream_t s);
void launch_by_point(const DevProblem &P, const AssemblyScratch &S, const PtGroup &g, cudaStream_t s);
void launch_camera_scatter(const DevProblem &P, const AssemblyScratch &S, const SysView &M, double *rhs, cudaStream_t s);
void launch_point_scatter(const DevProblem &P, const AssemblyScratch &S, const PtGroup &g, int kbase0, const SysView &M, double *rhs,
                          cudaStream_t s);
void launch_omega(const DevProblem &P, const AssemblyScratch &S, const double *dxref, double *omega_out, cudaStream_t s);
void launch_eval_k1(const DevProblem &P, int ns_max, double *a, double *w, double *p, cudaStream_t s);
void launch_scale_bars(const DevProblem &P, const SysView &M, double *rhs, cudaStream_t s);
void launch_omega_bars(const DevProblem &P, const double *dxref, double *omega_out, cudaStream_t s);

// ---- structured.cu: point-block solver (DESIGN.md "structured route") ------------------------------------------------------
struct StructDims {
    int up;        // object-coordinate columns (leading, block diagonal)
    int nc;        // remaining columns (interior orientation, distortion, exterior orientations)
    int d;         // datum rows
    int u;         // up + nc
    int64_t Tp;    // up padded to a multiple of 128 (leading dimension of Zt, Yt, T1t)
    int64_t mp;    // nc + d padded to a multiple of 128 (leading dimension of Kp)
    int64_t ncp;   // nc padded to a multiple of 128 (leading dimension of Sm)
    int64_t np;    // leading dimension of M, Btv, Tq
};
void launch_point_block_inv(const double *M, int64_t ld, const int32_t *blk_start, const int32_t *blk_size, int nBlk, const double *V,
                            double *Pinv, int *info, cudaStream_t s);
void launch_zero_point_blocks(double *M, int64_t ld, const int32_t *blk_start, const int32_t *blk_size, int nBlk, cudaStream_t s);
void launch_scale_yt(double *Yt, const StructDims &D, const double *V, cudaStream_t s);
void launch_build_zy(const double *M, const double *Btv, const StructDims &D, const int32_t *col_blk, const int32_t *blk_start,
                     const int32_t *blk_size, const double *Pinv, double *Zt, double *Yt, cudaStream_t s);
void launch_init_kp(const double *M, const double *Btv, const StructDims &D, double *Kp, cudaStream_t s);
void launch_border_prep(const double *Kp, const StructDims &D, double *Eb, double *ED, double *sb, cudaStream_t s);
void launch_form_stilde(const double *Kp, const StructDims &D, const double *Eb, const double *ED, double *Sm, cudaStream_t s);
void launch_border_f(const double *Sm, const StructDims &D, const double *ED, double *Fb, double *sb, cudaStream_t s);
void launch_fill_qprime(const double *Sm, const StructDims &D, const double *Fb, const double *sb, double *Kp, cudaStream_t s);
void launch_structured_solution(const double *nrm, const StructDims &D, const int32_t *col_blk, const int32_t *blk_start,
                                const int32_t *blk_size, const double *Pinv, const double *Zt, const double *Yt, const double *Kp,
                                const double *Btv, const double *V, double *zp, double *rp, double *yr, double *ys, double *Nt,
                                double *t, double *dxref, double *Tq, cudaStream_t s);
void launch_structured_place(double *M, const StructDims &D, const double *T1t, const double *Kp, const int32_t *blk_start,
                             const int32_t *blk_size, int nBlk, const double *Pinv, const double *V, cudaStream_t s);
void launch_structured_place_cols(double *X, int64_t ldx, int ntc, const int32_t *ktab, const int32_t *col_local, const StructDims &D,
                                  const double *T1l, int64_t ldt, const double *Kp, const int32_t *blk_start, const int32_t *blk_size,
                                  int nBlk, const double *Pinv, const double *V, cudaStream_t s);

}  // namespace jaicov
