// dense_kernels.cu -- FP64 tensor-core tile kernels behind dense_driver.hpp (sm_100a).
//
//  k_gemm<AL,BL>   C(128x128 tile) = alpha * op(A) op(B) + beta * C on FP64 tensor cores: mma.sync.m8n8k4.f64
//                  (SASS DMMA.8x8x4 -- the only FP64 tensor instruction of sm_100; tcgen05/TMEM have no f64 kind).
//                  256 threads = 8 warps (2 x 4), warp tile 64 x 32, 4-stage cp.async (LDGSTS) pipeline of
//                  128 x 16 operand slabs in padded shared memory (pitch == 4 mod 16 doubles => conflict-free
//                  LDS.64 fragment loads for both operand orientations).  Triangular operands skip their zero
//                  k-slabs per output tile; tri_out launches only the lower tiles; tiles are issued longest-first.
//  k_potrf_diag    one CTA: Cholesky of a 128 x 128 diagonal block in shared memory (left-looking), the factor is
//                  written back and then inverted in place (dtrti2-style) to give Dinv, which turns every
//                  triangular solve of the schedule into a GEMM.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.h"
#include "dense_driver.hpp"

namespace jaicov {

constexpr int GB = 128;       // tile edge
constexpr int GK = 16;        // k slab
constexpr int GSTAGES = 4;
constexpr int GTHREADS = 256;
constexpr int PITCH_K = 20;   // [128][16] slab stored with pitch 20 doubles
constexpr int PITCH_M = 132;  // [16][128] slab stored with pitch 132 doubles
constexpr int SLAB = 128 * PITCH_K;  // 2560 doubles >= 16 * 132
constexpr size_t GEMM_SMEM = (size_t)GSTAGES * 2 * SLAB * sizeof(double);

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// load one ROWS x 16 operand slab: LAYOUT 0 = rows are the m (or n) index, k contiguous in global memory;
// LAYOUT 1 = rows are k, the m (or n) index contiguous in global memory (pitch ROWS + 4)
template <int LAYOUT, int ROWS, int NTHREADS>
__device__ __forceinline__ void load_slab(double *s, const double *g, int64_t ld, int tid) {
    if (LAYOUT == 0) {
#pragma unroll
        for (int c = tid; c < ROWS * 8; c += NTHREADS) {
            const int row = c >> 3, cc = c & 7;
            cp_async16(s + row * PITCH_K + cc * 2, g + (int64_t)row * ld + cc * 2);
        }
    } else {
        constexpr int CPR = ROWS / 2;   // 16-byte chunks per k row
#pragma unroll
        for (int c = tid; c < 16 * CPR; c += NTHREADS) {
            const int kr = c / CPR, cc = c % CPR;
            cp_async16(s + kr * (ROWS + 4) + cc * 2, g + (int64_t)kr * ld + cc * 2);
        }
    }
}

// BM = rows of C per CTA (128, 64 or 32; the columns are always 128).  Small BM spreads a 128 x 128 tile over 2 or
// 4 SMs: a single tile with K = 128 is bound by ONE SM's DMMA rate (~17 us), which is what the many small launches
// of the recursion's deep levels pay; the host picks BM from the number of tiles of the launch.
// CTA shapes.  BMT = 129: the 128-row tile with 8 warps (2 x 4, warp tile 64 x 32) -- the default; BMT = 128: the same tile
// with 16 warps (4 x 4, warp tile 32 x 32), tried because "wait" (fixed DMMA issue latency) was the top stall reason with two
// warps per scheduler, but measured slower (more LDS per DMMA); 64- and 32-row tiles use 8 warps.
template <int BMT> struct GemmShape {
    static constexpr int ROWS = (BMT == 129) ? 128 : BMT;
    static constexpr int WARPS_M = (BMT == 128) ? 4 : ((BMT == 129 || BMT == 64) ? 2 : 1);
    static constexpr int WARPS_N = (BMT == 128) ? 4 : ((BMT == 129 || BMT == 64) ? 4 : 8);
    static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
};

template <int AL, int BL, int BMT>
__global__ void __launch_bounds__(GemmShape<BMT>::THREADS, 1) k_gemm(GemmDesc g) {
    constexpr int BM = GemmShape<BMT>::ROWS;
    constexpr int WARPS_M = GemmShape<BMT>::WARPS_M;
    constexpr int WARPS_N = GemmShape<BMT>::WARPS_N;
    constexpr int NTHREADS = GemmShape<BMT>::THREADS;
    constexpr int WT_M = BM / WARPS_M, WT_N = 128 / WARPS_N;
    constexpr int MI = WT_M / 8, NJ = WT_N / 8;
    constexpr int SPLIT = 128 / BM;
    extern __shared__ __align__(16) double gsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int grp = lane >> 2, tig = lane & 3;

    // ---- tile decode (longest contraction first) -----------------------------------------------------------------
    int it, jt;
    const int sub = blockIdx.x % SPLIT;
    if (g.coltab) {
        it = blockIdx.x / SPLIT;
        jt = g.coltab[blockIdx.y] / GB;
        if (it < jt) return;    // above the diagonal of this column tile
    } else {
        const int l = blockIdx.x / SPLIT;
        if (g.tri_out) {
            it = (int)((sqrt(8.0 * (double)l + 1.0) - 1.0) * 0.5);
            while ((int64_t)(it + 1) * (it + 2) / 2 <= l) it++;
            while ((int64_t)it * (it + 1) / 2 > l) it--;
            jt = l - (int)((int64_t)it * (it + 1) / 2);
        } else {
            it = l / g.nt;
            jt = l - it * g.nt;
            if (g.kmode == K_A_LOWER) it = g.mt - 1 - it;
        }
    }
    int64_t kbeg = 0, kend = g.K;
    if (g.kmode == K_B_LOWER) kbeg = (int64_t)jt * GB;
    else if (g.kmode == K_A_LOWER) kend = min(g.K, (int64_t)(it + 1) * GB);
    else if (g.kmode == K_MAX_IJ) kbeg = (int64_t)max(it, jt) * GB;
    else if (g.kmode == K_COL_BEG) {
        kbeg = max((int64_t)0, (int64_t)g.ktab[jt] - g.koff);
        if (kbeg >= g.K) return;            // this column tile is still zero in the whole contraction range
    } else if (g.kmode == K_ROW_MASK) {
        if (g.roff + (int64_t)it * GB < (int64_t)g.ktab[jt]) return;   // above the column's diagonal: not wanted
    }
    const int nk = (int)((kend - kbeg) / GK);
    const int64_t mrow0 = (int64_t)it * GB + sub * BM;   // first row of C (and of op(A)) of this CTA

    const double *Ag = (AL == 0) ? g.A + mrow0 * g.lda + kbeg : g.A + kbeg * g.lda + mrow0;
    const double *Bg = (BL == 0) ? g.B + (int64_t)jt * GB * g.ldb + kbeg : g.B + kbeg * g.ldb + (int64_t)jt * GB;
    const int64_t a_step = (AL == 0) ? GK : (int64_t)GK * g.lda;
    const int64_t b_step = (BL == 0) ? GK : (int64_t)GK * g.ldb;

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // ---- pipeline prologue ---------------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        if (s < nk) {
            load_slab<AL, BM, NTHREADS>(gsm + (size_t)(2 * s) * SLAB, Ag + s * a_step, g.lda, tid);
            load_slab<BL, 128, NTHREADS>(gsm + (size_t)(2 * s + 1) * SLAB, Bg + s * b_step, g.ldb, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        {   // prefetch slab kt + STAGES - 1 into the buffer consumed at iteration kt - 1
            const int kn = kt + GSTAGES - 1;
            if (kn < nk) {
                const int sb = kn % GSTAGES;
                load_slab<AL, BM, NTHREADS>(gsm + (size_t)(2 * sb) * SLAB, Ag + kn * a_step, g.lda, tid);
                load_slab<BL, 128, NTHREADS>(gsm + (size_t)(2 * sb + 1) * SLAB, Bg + kn * b_step, g.ldb, tid);
            }
            cp_async_commit();
        }
        const double *As = gsm + (size_t)(2 * (kt % GSTAGES)) * SLAB;
        const double *Bs = As + SLAB;
#pragma unroll
        for (int kk = 0; kk < GK / 4; kk++) {
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; i++)
                a[i] = (AL == 0) ? As[(wm * WT_M + 8 * i + grp) * PITCH_K + kk * 4 + tig]
                                 : As[(kk * 4 + tig) * (BM + 4) + wm * WT_M + 8 * i + grp];
#pragma unroll
            for (int j = 0; j < NJ; j++)
                b[j] = (BL == 0) ? Bs[(wn * WT_N + 8 * j + grp) * PITCH_K + kk * 4 + tig]
                                 : Bs[(kk * 4 + tig) * PITCH_M + wn * WT_N + 8 * j + grp];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // every load of this CTA has landed before any store: in-place strips (C == A) are safe

    // ---- epilogue ------------------------------------------------------------------------------------------------
    double *Cg = g.C + (mrow0 + wm * WT_M + grp) * g.ldc + (int64_t)jt * GB + wn * WT_N + 2 * tig;
    const double alpha = g.alpha, beta = g.beta;
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            double2 *p = reinterpret_cast<double2 *>(Cg + (int64_t)(8 * i) * g.ldc + 8 * j);
            double2 v;
            if (beta != 0.0) {
                const double2 c = *p;
                v.x = alpha * acc[i][j][0] + beta * c.x;
                v.y = alpha * acc[i][j][1] + beta * c.y;
            } else {
                v.x = alpha * acc[i][j][0];
                v.y = alpha * acc[i][j][1];
            }
            *p = v;
        }
}

template <int AL, int BL, int BMT>
static void launch_gemm_t(const GemmDesc &g, int64_t tiles, cudaStream_t s) {
    constexpr int BM = GemmShape<BMT>::ROWS;
    static bool attr = false;
    if (!attr) {
        JCHECK(cudaFuncSetAttribute(k_gemm<AL, BL, BMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        attr = true;
    }
    g_launch_count++;
    constexpr int NT = GemmShape<BMT>::THREADS;
    if (g.coltab) k_gemm<AL, BL, BMT><<<dim3((unsigned)(g.mt * (128 / BM)), (unsigned)g.ncoltab), NT, GEMM_SMEM, s>>>(g);
    else k_gemm<AL, BL, BMT><<<(unsigned)(tiles * (128 / BM)), NT, GEMM_SMEM, s>>>(g);
}

template <int AL, int BL>
static void launch_gemm_l(const GemmDesc &g, cudaStream_t s) {
    const int64_t tiles = g.tri_out ? (int64_t)g.mt * (g.mt + 1) / 2 : (int64_t)g.mt * g.nt;
    if (tiles <= 0) return;
    // a CTA owns BM full-width rows of C; when C aliases the B operand (left-side base cases) the whole 128 x 128
    // tile must stay with one CTA
    const bool alias_b = (const double *)g.C == g.B;
    // A/B measured on B200 (config 5, one final pass): 8 warps 8.06 s, 16 warps 8.31 s -> 8 warps (2 x 4, warp tile 64 x 32)
    // is the default; JAICOV_GEMM_WARPS=16 selects the 4 x 4 shape
    static const bool warps16 = [] { const char *e = getenv("JAICOV_GEMM_WARPS"); return e && atoi(e) == 16; }();
    if (alias_b || tiles >= 148) {
        if (warps16) launch_gemm_t<AL, BL, 128>(g, tiles, s);
        else launch_gemm_t<AL, BL, 129>(g, tiles, s);
    }
    else if (tiles >= 74) launch_gemm_t<AL, BL, 64>(g, tiles, s);
    else launch_gemm_t<AL, BL, 32>(g, tiles, s);
}

void launch_gemm(const GemmDesc &g, cudaStream_t s) {
    if (g.al == 0 && g.bl == 0) launch_gemm_l<0, 0>(g, s);
    else if (g.al == 0 && g.bl == 1) launch_gemm_l<0, 1>(g, s);
    else if (g.al == 1 && g.bl == 1) launch_gemm_l<1, 1>(g, s);
    else launch_gemm_l<1, 0>(g, s);
}

// ---- diagonal block: factor + invert ----------------------------------------------------------------------------
// 512 threads: 4 lanes per matrix row split every dot product (k = q mod 4) and combine with two shuffles, so a
// column step costs <= 32 (2 LDS + FMA) per lane instead of 128.  The shared-memory pitch is == 4 (mod 16) doubles:
// the 16 lanes of a half-warp (4 rows x 4 k-phases) hit 16 distinct 8-byte banks.  The pivot dot product is
// recomputed by every row (no broadcast), which leaves ONE barrier per column in the factorisation.
constexpr int DP = 132;
constexpr int DTHREADS = 512;

__global__ void __launch_bounds__(DTHREADS, 1) k_potrf_diag(double *__restrict__ A, int64_t ld, double *__restrict__ dinv,
                                                             int row0, int *__restrict__ info) {
    extern __shared__ double a[];   // [128][DP]
    __shared__ double s_diag[128], s_rdiag[128];
    const int tid = threadIdx.x;
    const int i = tid >> 2, q = tid & 3;
    // load the lower triangle, coalesced: 4 rows of 128 per pass
    {
        const int c = tid & 127, r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < 128; r += 4) a[r * DP + c] = (c <= r) ? A[(int64_t)r * ld + c] : 0.0;
    }
    __syncthreads();
    // left-looking Cholesky
    const double *ri = a + i * DP;
    for (int j = 0; j < 128; j++) {
        const double *rj = a + j * DP;
        double s = 0.0, p = 0.0;
        if (i >= j) {
            // 4 independent accumulator pairs, 4 k-values in flight: the LDS latency is paid once per 16 columns
            double s1 = 0.0, s2 = 0.0, s3 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            int k = q;
            for (; k + 12 < j; k += 16) {
                const double x0 = rj[k], x1 = rj[k + 4], x2 = rj[k + 8], x3 = rj[k + 12];
                const double y0 = ri[k], y1 = ri[k + 4], y2 = ri[k + 8], y3 = ri[k + 12];
                s += y0 * x0;  p += x0 * x0;
                s1 += y1 * x1; p1 += x1 * x1;
                s2 += y2 * x2; p2 += x2 * x2;
                s3 += y3 * x3; p3 += x3 * x3;
            }
            for (; k < j; k += 4) {
                const double x = rj[k];
                s += ri[k] * x;
                p += x * x;
            }
            s += (s1 + s2) + s3;
            p += (p1 + p2) + p3;
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        // the factored diagonal goes to s_diag so that a[j][j] keeps its input value: nothing read in this step is
        // overwritten in this step (a[i][j] is read and written by the same thread), hence a single barrier
        // one reciprocal square root per column instead of sqrt + divide on the critical path
        if (i >= j && q == 0) {
            const double piv = rj[j] - p;
            const double rinv = rsqrt(piv);
            if (i == j) {
                if (!(piv > 0.0)) atomicCAS(info, 0, row0 + j + 1);
                s_diag[j] = piv * rinv;
                s_rdiag[j] = rinv;
            } else {
                a[i * DP + j] = (ri[j] - s) * rinv;
            }
        }
        __syncthreads();
    }
    if (tid < 128) a[tid * DP + tid] = s_diag[tid];
    __syncthreads();
    // write the factor back (lower triangle only)
    {
        const int c = tid & 127, r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < 128; r += 4)
            if (c <= r) A[(int64_t)r * ld + c] = a[r * DP + c];
    }
    // in-place inversion of the lower-triangular factor, last column first (dtrti2, lower):
    //   inv[j][j] = 1/L[j][j];  inv[i][j] = -(sum_{k=j+1..i} inv[i][k] L[k][j]) * inv[j][j]
    // reciprocal diagonal, one Newton step on the factorisation's rsqrt: 1/L_jj to full precision, computed for all
    // columns at once instead of one division per column step
    if (tid < 128) {
        const double dgg = s_diag[tid];
        double r = s_rdiag[tid];
        r = r + r * (1.0 - dgg * r);
        s_rdiag[tid] = r;
    }
    __syncthreads();
    for (int j = 127; j >= 0; j--) {
        const double ajj = s_rdiag[j];
        double s = 0.0;
        if (i > j) {
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = j + 1 + q;
            for (; k + 12 <= i; k += 16) {
                const double x0 = a[k * DP + j], x1 = a[(k + 4) * DP + j], x2 = a[(k + 8) * DP + j], x3 = a[(k + 12) * DP + j];
                const double y0 = ri[k], y1 = ri[k + 4], y2 = ri[k + 8], y3 = ri[k + 12];
                s += y0 * x0; s1 += y1 * x1; s2 += y2 * x2; s3 += y3 * x3;
            }
            for (; k <= i; k += 4) s += ri[k] * a[k * DP + j];
            s += (s1 + s2) + s3;
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        __syncthreads();
        if (q == 0) {
            if (i > j) a[i * DP + j] = -s * ajj;
            else if (i == j) a[j * DP + j] = ajj;
        }
        __syncthreads();
    }
    {
        const int c = tid & 127, r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < 128; r += 4) dinv[r * 128 + c] = (c <= r) ? a[r * DP + c] : 0.0;
    }
}

void launch_potrf_diag(double *A, int64_t ld, double *dinv, int row0, int *info, cudaStream_t s) {
    static bool attr = false;
    const size_t smem = (size_t)128 * DP * sizeof(double);
    if (!attr) {
        JCHECK(cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    g_launch_count++;
    k_potrf_diag<<<1, DTHREADS, smem, s>>>(A, ld, dinv, row0, info);
}

// ---- 2-D copy ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_copy2d(double *__restrict__ dst, int64_t ldd, const double *__restrict__ src,
                                                int64_t lds, int64_t rows, int64_t cols2) {
    // cols2 = cols / 2 (everything is 16-byte aligned: cols and leading dimensions are even)
    const int64_t total = rows * cols2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols2, c = i - r * cols2;
        reinterpret_cast<double2 *>(dst + r * ldd)[c] = reinterpret_cast<const double2 *>(src + r * lds)[c];
    }
}

void launch_copy2d(double *dst, int64_t ldd, const double *src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t s) {
    const int64_t total = rows * (cols / 2);
    if (total <= 0) return;
    const int64_t blocks = (total + 255) / 256;
    g_launch_count++;
    k_copy2d<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, s>>>(dst, ldd, src, lds, rows, cols / 2);
}

}  // namespace jaicov

namespace jaicov {

// ---- skinny triangular solves: NR = 8 right-hand sides stored as rows R[8][np] ------------------------------------
// One launch per 128-block step, right-looking, every launch spread over the rows (forward) or columns (backward)
// still to be updated; the 128 x 128 diagonal solve is the multiplication by Dinv and is recomputed by every CTA
// (128 KB from L2) instead of being a separate launch.  HBM/L2-bound: L is read exactly once per sweep.
constexpr int SR = 8;          // right-hand sides
constexpr int SCHUNK = 256;    // rows (forward) / columns (backward) per CTA

// forward step j: y_j = Dinv_j b_j;  b[i] -= L[i, J] y_j for all rows i below block j
// (b lives in R and is updated in place below block j; y_j goes to the separate buffer Y so that no CTA can see a
// half-updated block j)
__global__ void __launch_bounds__(256) k_solve_fwd_step(const double *__restrict__ L, int64_t ld, const double *__restrict__ dinv,
                                                        double *__restrict__ R, double *__restrict__ Y, int64_t np, int j) {
    __shared__ double sb[SR][128];
    __shared__ double sy[SR][128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t J = (int64_t)j * 128;
    for (int i = tid; i < SR * 128; i += 256) sb[i >> 7][i & 127] = R[(int64_t)(i >> 7) * np + J + (i & 127)];
    __syncthreads();
    const double *D = dinv + J * 128;
    // y[t] = sum_k Dinv[t][k] b[k]: one warp per row, lanes over k; 4 rows (16 loads) in flight per lane
    for (int t0 = 4 * warp; t0 < 128; t0 += 32) {
        double dv[4][4];
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int kk = 0; kk < 4; kk++) dv[q][kk] = D[(t0 + q) * 128 + lane + 32 * kk];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double acc[SR];
#pragma unroll
            for (int r = 0; r < SR; r++) acc[r] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
#pragma unroll
                for (int r = 0; r < SR; r++) acc[r] += dv[q][kk] * sb[r][lane + 32 * kk];
#pragma unroll
            for (int r = 0; r < SR; r++) {
                double v = acc[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sy[r][t0 + q] = v;
            }
        }
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = tid; i < SR * 128; i += 256) Y[(int64_t)(i >> 7) * np + J + (i & 127)] = sy[i >> 7][i & 127];
    // rows below: each warp takes 4 rows per iteration so that 16 independent loads are in flight per lane
    const int64_t row0 = J + 128 + (int64_t)blockIdx.x * SCHUNK;
    const int64_t row1 = min(np, row0 + SCHUNK);
    for (int64_t i0 = row0 + 4 * warp; i0 < row1; i0 += 32) {
        double lv[4][4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int64_t i = min(i0 + q, row1 - 1);
            const double *Li = L + i * ld + J;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) lv[q][kk] = Li[lane + 32 * kk];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double acc[SR];
#pragma unroll
            for (int r = 0; r < SR; r++) acc[r] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
#pragma unroll
                for (int r = 0; r < SR; r++) acc[r] += lv[q][kk] * sy[r][lane + 32 * kk];
            }
#pragma unroll
            for (int r = 0; r < SR; r++) {
                double v = acc[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[r] = v;
            }
            if (lane < SR && i0 + q < row1) {
                double v = acc[0];
#pragma unroll
                for (int r = 1; r < SR; r++) v = (lane == r) ? acc[r] : v;
                R[(int64_t)lane * np + i0 + q] -= v;
            }
        }
    }
}

// backward step j: x_j = Dinv_j' y_j;  y[c] -= sum_{i in J} L[i][c] x_j[i] for all columns c left of block j
// (y lives in Y and is updated in place left of block j; x_j goes to R)
__global__ void __launch_bounds__(256) k_solve_bwd_step(const double *__restrict__ L, int64_t ld, const double *__restrict__ dinv,
                                                        double *__restrict__ R, double *__restrict__ Y, int64_t np, int j) {
    __shared__ double sy[SR][128];
    __shared__ double sx[SR][128];
    const int tid = threadIdx.x;
    const int64_t J = (int64_t)j * 128;
    for (int i = tid; i < SR * 128; i += 256) sy[i >> 7][i & 127] = Y[(int64_t)(i >> 7) * np + J + (i & 127)];
    __syncthreads();
    const double *D = dinv + J * 128;
    {   // x[t] = sum_k Dinv[k][t] y[k]  (coalesced over t; Dinv[k][t] = 0 for k < t, so the loop is branch-free and
        // unrolled: 16 independent loads in flight).  Two threads per t split k into halves.
        const int t = tid & 127, half = tid >> 7;
        double acc[SR];
#pragma unroll
        for (int r = 0; r < SR; r++) acc[r] = 0.0;
#pragma unroll 16
        for (int kk = 0; kk < 64; kk++) {
            const int k = half * 64 + kk;
            const double dv = D[k * 128 + t];
#pragma unroll
            for (int r = 0; r < SR; r++) acc[r] += dv * sy[r][k];
        }
        if (half == 1) {
#pragma unroll
            for (int r = 0; r < SR; r++) sx[r][t] = acc[r];
        }
        __syncthreads();
        if (half == 0) {
#pragma unroll
            for (int r = 0; r < SR; r++) sx[r][t] += acc[r];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = tid; i < SR * 128; i += 256) R[(int64_t)(i >> 7) * np + J + (i & 127)] = sx[i >> 7][i & 127];
    if (j == 0) return;
    const int64_t c0 = (int64_t)blockIdx.x * SCHUNK;
    for (int64_t c = c0 + tid; c < min(J, c0 + SCHUNK); c += 256) {
        double acc[SR];
#pragma unroll
        for (int r = 0; r < SR; r++) acc[r] = 0.0;
        const double *Lc = L + J * ld + c;
#pragma unroll 16
        for (int i = 0; i < 128; i++) {
            const double lv = Lc[(int64_t)i * ld];
#pragma unroll
            for (int r = 0; r < SR; r++) acc[r] += lv * sx[r][i];
        }
#pragma unroll
        for (int r = 0; r < SR; r++) Y[(int64_t)r * np + c] -= acc[r];
    }
}

// R[8][np] <- (L L')^-1 R' (rows are right-hand sides), L = lower factor in M, Dinv = inverted diagonal blocks
// Y: scratch of the same size as R
void launch_solve_rows8(const double *L, int64_t ld, const double *dinv, double *R, double *Y, int64_t np, cudaStream_t s) {
    const int nb = (int)(np / 128);
    for (int j = 0; j < nb; j++) {
        const int64_t below = np - (int64_t)(j + 1) * 128;
        const int grid = (int)std::max<int64_t>(1, (below + SCHUNK - 1) / SCHUNK);
        g_launch_count++;
        k_solve_fwd_step<<<grid, 256, 0, s>>>(L, ld, dinv, R, Y, np, j);
    }
    for (int j = nb - 1; j >= 0; j--) {
        const int64_t left = (int64_t)j * 128;
        const int grid = (int)std::max<int64_t>(1, (left + SCHUNK - 1) / SCHUNK);
        g_launch_count++;
        k_solve_bwd_step<<<grid, 256, 0, s>>>(L, ld, dinv, R, Y, np, j);
    }
}

}  // namespace jaicov
