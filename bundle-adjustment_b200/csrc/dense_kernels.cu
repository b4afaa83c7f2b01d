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
constexpr int PITCH_K = 20;   // [128][16] slab stored with pitch 20 doubles
constexpr int PITCH_M = 132;  // [16][128] slab stored with pitch 132 doubles
constexpr int SLAB = 128 * PITCH_K;  // 2560 doubles >= 16 * 132

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
// CTA shapes.  BMT = 129: the 128-row tile with 8 warps (2 x 4, warp tile 64 x 32); BMT = 128: the same tile
// with 16 warps (4 x 4, warp tile 32 x 32), tried because "wait" (fixed DMMA issue latency) was the top stall reason with two
// warps per scheduler, but measured slower (more LDS per DMMA); 64- and 32-row tiles use 8 warps.
// BMT = 65: 64-row tile, 4 warps (2 x 2, warp tile 32 x 64), 3 stages, TWO CTAs per SM: independent CTAs hide each other's
// barrier and prologue/epilogue bubbles (the shape class cuBLAS's own d884 kernel uses, 64 x 128 x 16 x 3).  Default for
// every launch of >= 148 tiles: 96 % of the measured cuBLAS DGEMM rate over a whole config-5 factor + inverse (8 warps: 88 %).
template <int BMT> struct GemmShape {
    static constexpr int ROWS = (BMT == 129) ? 128 : (BMT == 65 ? 64 : BMT);
    static constexpr int WARPS_M = (BMT == 128) ? 4 : ((BMT == 129 || BMT == 64 || BMT == 65) ? 2 : 1);
    static constexpr int WARPS_N = (BMT == 128) ? 4 : ((BMT == 129 || BMT == 64) ? 4 : (BMT == 65 ? 2 : 8));
    static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
    static constexpr int STAGES = (BMT == 65) ? 3 : GSTAGES;
    static constexpr int CTAS_PER_SM = (BMT == 65) ? 2 : 1;
    static constexpr int A_SLAB = (BMT == 65) ? 64 * PITCH_K : SLAB;      // doubles per A slab
    static constexpr size_t SMEM = (size_t)STAGES * (A_SLAB + SLAB) * sizeof(double);
};

template <int AL, int BL, int BMT>
__global__ void __launch_bounds__(GemmShape<BMT>::THREADS, GemmShape<BMT>::CTAS_PER_SM) k_gemm(GemmDesc g) {
    constexpr int BM = GemmShape<BMT>::ROWS;
    constexpr int NSTAGES = GemmShape<BMT>::STAGES;
    constexpr int ASLAB = GemmShape<BMT>::A_SLAB;
    constexpr int STAGE = ASLAB + SLAB;       // doubles per pipeline stage: A slab then B slab
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
        if (!g.coltab_full && it < jt) return;    // above the diagonal of this column tile
    } else {
        const int l = blockIdx.x / SPLIT;
        if (g.tri_out) {
            tri_tile_decode(l, g.mt, g.tile_band, it, jt);
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
    for (int s = 0; s < NSTAGES - 1; s++) {
        if (s < nk) {
            load_slab<AL, BM, NTHREADS>(gsm + (size_t)s * STAGE, Ag + s * a_step, g.lda, tid);
            load_slab<BL, 128, NTHREADS>(gsm + (size_t)s * STAGE + ASLAB, Bg + s * b_step, g.ldb, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<NSTAGES - 2>();
        __syncthreads();
        {   // prefetch slab kt + STAGES - 1 into the buffer consumed at iteration kt - 1
            const int kn = kt + NSTAGES - 1;
            if (kn < nk) {
                const int sb = kn % NSTAGES;
                load_slab<AL, BM, NTHREADS>(gsm + (size_t)sb * STAGE, Ag + kn * a_step, g.lda, tid);
                load_slab<BL, 128, NTHREADS>(gsm + (size_t)sb * STAGE + ASLAB, Bg + kn * b_step, g.ldb, tid);
            }
            cp_async_commit();
        }
        const double *As = gsm + (size_t)(kt % NSTAGES) * STAGE;
        const double *Bs = As + ASLAB;
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
    const int64_t ctile = (g.coltab && g.c_local) ? (int64_t)blockIdx.y : (int64_t)jt;
    double *Cg = g.C + (mrow0 + wm * WT_M + grp) * g.ldc + ctile * GB + wn * WT_N + 2 * tig;
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
    static PerDeviceOnce once;
    once.run([] { JCHECK(cudaFuncSetAttribute(k_gemm<AL, BL, BMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmShape<BMT>::SMEM)); });
    g_launch_count++;
    constexpr int NT = GemmShape<BMT>::THREADS;
    constexpr size_t SM = GemmShape<BMT>::SMEM;
    if (g.coltab) k_gemm<AL, BL, BMT><<<dim3((unsigned)(g.mt * (128 / BM)), (unsigned)g.ncoltab), NT, SM, s>>>(g);
    else k_gemm<AL, BL, BMT><<<(unsigned)(tiles * (128 / BM)), NT, SM, s>>>(g);
}

template <int AL, int BL>
static void launch_gemm_l(const GemmDesc &g, cudaStream_t s) {
    const int64_t tiles = g.tri_out ? (int64_t)g.mt * (g.mt + 1) / 2 : (int64_t)g.mt * g.nt;
    if (tiles <= 0) return;
    // a CTA owns BM full-width rows of C; when C aliases the B operand (left-side base cases) the whole 128 x 128
    // tile must stay with one CTA
    const bool alias_b = (const double *)g.C == g.B;
    // A/B measured on B200 (config 5, one final pass; profiles/r01_gemm_shape_ab.txt): 4 warps x 2 CTAs per SM 7.39 s,
    // 8 warps 8.05 s, 16 warps 8.31 s -> the 64 x 128 / 4-warp shape is the default for launches that fill the machine
    // twice over; JAICOV_GEMM_WARPS=8 or 16 select the one-CTA-per-SM shapes
    static const int shape = [] { const char *e = getenv("JAICOV_GEMM_WARPS"); return e ? atoi(e) : 4; }();
    if (alias_b) launch_gemm_t<AL, BL, 129>(g, tiles, s);
    else if (tiles >= 148) {
        if (shape == 16) launch_gemm_t<AL, BL, 128>(g, tiles, s);
        else if (shape == 8) launch_gemm_t<AL, BL, 129>(g, tiles, s);
        else launch_gemm_t<AL, BL, 65>(g, tiles, s);      // 64 x 128 tiles, 2 CTAs per SM
    }
    else if (tiles >= 74) launch_gemm_t<AL, BL, 64>(g, tiles, s);
    else launch_gemm_t<AL, BL, 32>(g, tiles, s);
}

bool launch_gemm_ozaki(const GemmDesc &g, cudaStream_t s);   // ozaki.cu: int8 digit products on tcgen05, takes the big launches (>= 148 tiles, K >= 1024) unless switched off

void launch_gemm(const GemmDesc &g_in, cudaStream_t s) {
    // switch (default off, measured: no gain on the DMMA-bound launches): band-swizzled tile order of the lower-triangular launches, DESIGN.md section 9.4
    static const int band = [] { const char *e = getenv("JAICOV_TILE_BAND"); return e ? atoi(e) : 0; }();
    GemmDesc g = g_in;
    if (g.tri_out && band > 0 && g.tile_band == 0) g.tile_band = band;
    if (launch_gemm_ozaki(g, s)) return;
    if (g.al == 0 && g.bl == 0) launch_gemm_l<0, 0>(g, s);
    else if (g.al == 0 && g.bl == 1) launch_gemm_l<0, 1>(g, s);
    else if (g.al == 1 && g.bl == 1) launch_gemm_l<1, 1>(g, s);
    else launch_gemm_l<1, 0>(g, s);
}

// ---- diagonal block: factor + invert ----------------------------------------------------------------------------
// The 128 x 128 diagonal block is on the critical path of every 128-column step of the factorisation, so this kernel
// is built for latency, not throughput (one CTA, 256 threads, everything in shared memory / registers):
//   phase A, four 32-wide panels:  (1) warp 0 factors the 32 x 32 diagonal sub-block IN REGISTERS (lane = row,
//            register = column; pivots and column entries travel by warp shuffles, no barrier inside) and inverts it
//            the same way (lane = column of the inverse);  (2) all warps: panel below = B * inv(L_kk)' ;
//            (3) all warps: trailing update  A22 -= P P'.
//   phase B: the inverse of the whole 128 x 128 factor from the four 32 x 32 inverses, block column by block column,
//            in place:  X_IJ = -X_II * sum_{K=J..I-1} L_IK X_KJ.
// Replaces a column-by-column version (128 + 128 barrier-separated steps, 205 us measured); see profiles/.
constexpr int DP = 132;        // pitch of the 128 x 128 block in shared memory (== 4 mod 16)
constexpr int SP = 33;         // pitch of the 32 x 32 scratch blocks
constexpr int DTHREADS = 256;
constexpr size_t DIAG_SMEM = ((size_t)128 * DP + 4 * 32 * SP + 96 * SP + 32 * SP) * sizeof(double);

__global__ void __launch_bounds__(DTHREADS, 1) k_potrf_diag(double *__restrict__ A, int64_t ld, double *__restrict__ dinv,
                                                             int row0, int *__restrict__ info) {
    extern __shared__ double dsm[];
    double *a = dsm;                       // [128][DP]   the block: A -> L -> (phase B) inverse, lower triangle
    double *sinv = a + 128 * DP;           // [4][32][SP] inverses of the 32 x 32 diagonal sub-blocks of L
    double *pan = sinv + 4 * 32 * SP;      // [96][SP]    panel below the current diagonal sub-block
    double *tmp = pan + 96 * SP;           // [32][SP]    phase B scratch; phase A: column broadcast buffer
    __shared__ double s_rdiag[128];        // 1 / L_ii
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    // load the lower triangle (coalesced rows), zero above the diagonal
    {
        const int c = tid & 127, r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < 128; r += 2) a[r * DP + c] = (c <= r) ? A[(int64_t)r * ld + c] : 0.0;
    }
    __syncthreads();

    // ---------------------------------------------------------------- phase A: Cholesky, four 32-wide panels
    for (int kb = 0; kb < 4; kb++) {
        const int base = 32 * kb;
        const int nr = 96 - base;          // rows below the diagonal sub-block
        if (warp == 0) {
            // 32 x 32 diagonal sub-block in registers: lane = row, d[c] = column.  The freshly scaled column travels
            // to the other lanes through a 32-double shared buffer (one LDS.128 per two rows instead of four SHFL).
            double d[32];
#pragma unroll
            for (int c = 0; c < 32; c++) d[c] = a[(base + lane) * DP + base + c];
            double *col = tmp;             // [32]
#pragma unroll
            for (int jj = 0; jj < 32; jj++) {
                const double pj = __shfl_sync(FULL, d[jj], jj);
                const double r = rsqrt(pj);
                const double dj = (lane == jj) ? pj * r : d[jj] * r;
                d[jj] = dj;
                if (lane == jj) {
                    if (!(pj > 0.0)) atomicCAS(info, 0, row0 + base + jj + 1);
                    s_rdiag[base + jj] = r + r * (1.0 - dj * r);    // 1 / L_jj, one Newton step on the rsqrt
                }
                if (jj < 31) {
                    col[lane] = dj;
                    __syncwarp();
#pragma unroll
                    for (int c = jj + 1; c < 32; c++) d[c] -= dj * col[c];   // meaningful for lanes >= c
                    __syncwarp();
                }
            }
#pragma unroll
            for (int c = 0; c < 32; c++)
                if (c <= lane) a[(base + lane) * DP + base + c] = d[c];
        }
        __syncthreads();
        if (nr > 0) {
            // panel below: row r solves x L_kk' = b by forward substitution (every thread its own row, L_kk broadcast)
            if (tid < nr) {
                double *brow = a + (base + 32 + tid) * DP + base;
                const double *Lk = a + base * DP + base;
                double x[32];
#pragma unroll
                for (int c = 0; c < 32; c++) x[c] = brow[c];
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    double s0 = x[c], s1 = 0.0;
#pragma unroll
                    for (int k = 0; k + 1 < c; k += 2) {
                        s0 -= x[k] * Lk[c * DP + k];
                        s1 -= x[k + 1] * Lk[c * DP + k + 1];
                    }
                    if (c & 1) s0 -= x[c - 1] * Lk[c * DP + c - 1];
                    x[c] = (s0 + s1) * s_rdiag[base + c];
                }
#pragma unroll
                for (int c = 0; c < 32; c++) { brow[c] = x[c]; pan[tid * SP + c] = x[c]; }
            }
            __syncthreads();
            // trailing update (lower part): A22 -= P P'
            for (int idx = tid; idx < nr * nr; idx += DTHREADS) {
                const int r = idx / nr, c = idx - r * nr;
                if (c > r) continue;
                const double *pr = pan + r * SP, *pc = pan + c * SP;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    s0 += pr[k] * pc[k];
                    s1 += pr[k + 1] * pc[k + 1];
                    s2 += pr[k + 2] * pc[k + 2];
                    s3 += pr[k + 3] * pc[k + 3];
                }
                a[(base + 32 + r) * DP + base + 32 + c] -= (s0 + s1) + (s2 + s3);
            }
            __syncthreads();
        }
    }
    // write the factor back (lower triangle only)
    {
        const int c = tid & 127, r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < 128; r += 2)
            if (c <= r) A[(int64_t)r * ld + c] = a[r * DP + c];
    }
    // ---------------------------------------------------------------- the four 32 x 32 inverses, one warp each
    if (warp < 4) {
        // lane c owns column c of X = L_kk^-1:  X[i][c] = (delta_ic - sum_{k<i} L[i][k] X[k][c]) / L[i][i]
        const int base = 32 * warp;
        const double *Lk = a + base * DP + base;
        double xc[32];
#pragma unroll
        for (int i2 = 0; i2 < 32; i2++) {
            double s0 = (lane == i2) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k + 1 < i2; k += 2) {
                s0 -= Lk[i2 * DP + k] * xc[k];          // xc[k] = X[k][lane] is zero for k < lane
                s1 -= Lk[i2 * DP + k + 1] * xc[k + 1];
            }
            if (i2 & 1) s0 -= Lk[i2 * DP + i2 - 1] * xc[i2 - 1];
            xc[i2] = (lane <= i2) ? (s0 + s1) * s_rdiag[base + i2] : 0.0;
        }
#pragma unroll
        for (int i2 = 0; i2 < 32; i2++) sinv[(warp * 32 + i2) * SP + lane] = xc[i2];
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase B
    // block columns J ascending, block rows I ascending, in place: X_IJ overwrites L_IJ once nothing needs L_IJ any more
    for (int J = 0; J < 3; J++) {
        for (int I = J + 1; I < 4; I++) {
            // T = sum_{K=J}^{I-1} L_IK X_KJ   (X_JJ = sinv[J]; X_KJ for K > J already sits in a[K][J])
            for (int idx = tid; idx < 32 * 32; idx += DTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                double s0 = 0.0, s1 = 0.0;
                for (int K = J; K < I; K++) {
                    const double *lrow = a + (32 * I + r) * DP + 32 * K;
                    const double *xcol = (K == J) ? sinv + (J * 32) * SP + c : a + (32 * K) * DP + 32 * J + c;
                    const int xs = (K == J) ? SP : DP;
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        s0 += lrow[k] * xcol[k * xs];
                        s1 += lrow[k + 1] * xcol[(k + 1) * xs];
                    }
                }
                tmp[r * SP + c] = s0 + s1;
            }
            __syncthreads();
            // X_IJ = -X_II T
            for (int idx = tid; idx < 32 * 32; idx += DTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                const double *xr = sinv + (I * 32 + r) * SP;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    s0 += xr[k] * tmp[k * SP + c];
                    s1 += xr[k + 1] * tmp[(k + 1) * SP + c];
                }
                a[(32 * I + r) * DP + 32 * J + c] = -(s0 + s1);
            }
            __syncthreads();
        }
    }
    // Dinv: off-diagonal blocks from a, diagonal blocks from sinv, zeros above the diagonal
    {
        const int c = tid & 127, r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < 128; r += 2) {
            double v = 0.0;
            if (c <= r) v = ((c >> 5) == (r >> 5)) ? sinv[((r >> 5) * 32 + (r & 31)) * SP + (c & 31)] : a[r * DP + c];
            dinv[r * 128 + c] = v;
        }
    }
}

void launch_potrf_diag(double *A, int64_t ld, double *dinv, int row0, int *info, cudaStream_t s) {
    static PerDeviceOnce once;
    once.run([] { JCHECK(cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM)); });
    g_launch_count++;
    k_potrf_diag<<<1, DTHREADS, DIAG_SMEM, s>>>(A, ld, dinv, row0, info);
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
// (128 KB from L2) instead of being a separate launch.  Every product here is (rows x 128) * (128 x 8): exactly one
// DMMA n-tile wide, so the A fragments are loaded straight from global memory (8 rows x 32-byte sectors per load, all
// sectors fully used) and the 8 right-hand sides sit in shared memory as the B fragment; no shuffle reductions.
// HBM/L2-bound: L is read exactly once per sweep.
constexpr int SR = 8;          // right-hand sides
constexpr int SCHUNK = 256;    // rows (forward) / columns (backward) per CTA: 8 warps x 32
constexpr int SPITCH = 132;    // pitch of the right-hand-side tiles in shared memory (== 4 mod 16: conflict-free B fragments)

// acc[mt] += A_tile(mt) * B for NMT 8-row tiles; TRANS = false: A[m][k] at Ap[m * lda + k]; true: A[m][k] at Ap[k * lda + m]
template <int NMT, bool TRANS>
__device__ __forceinline__ void skinny_mma(const double *__restrict__ Ap, int64_t lda, const double *sB, double (&acc)[NMT][2],
                                           int lane) {
    const int grp = lane >> 2, tig = lane & 3;
#pragma unroll 8
    for (int ks = 0; ks < 32; ks++) {
        const double b = sB[grp * SPITCH + 4 * ks + tig];
        double a[NMT];
#pragma unroll
        for (int mt = 0; mt < NMT; mt++)
            a[mt] = TRANS ? Ap[(int64_t)(4 * ks + tig) * lda + 8 * mt + grp] : Ap[(int64_t)(8 * mt + grp) * lda + 4 * ks + tig];
#pragma unroll
        for (int mt = 0; mt < NMT; mt++) dmma884(acc[mt][0], acc[mt][1], a[mt], b);
    }
}

// forward step j: y_j = Dinv_j b_j;  b[i] -= L[i, J] y_j for all rows i below block j
// (b lives in R and is updated in place below block j; y_j goes to the separate buffer Y so that no CTA can see a
// half-updated block j)
__global__ void __launch_bounds__(256) k_solve_fwd_step(const double *__restrict__ L, int64_t ld, const double *__restrict__ dinv,
                                                        double *__restrict__ R, double *__restrict__ Y, int64_t np, int j) {
    __shared__ double sb[SR * SPITCH];
    __shared__ double sy[SR * SPITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    const int64_t J = (int64_t)j * 128;
    for (int i = tid; i < SR * 128; i += 256) sb[(i >> 7) * SPITCH + (i & 127)] = R[(int64_t)(i >> 7) * np + J + (i & 127)];
    __syncthreads();
    {   // y = Dinv_j b: warp w computes rows 16 w .. 16 w + 15
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        skinny_mma<2, false>(dinv + J * 128 + (int64_t)(16 * warp) * 128, 128, sb, acc, lane);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            sy[(2 * tig) * SPITCH + 16 * warp + 8 * mt + grp] = acc[mt][0];
            sy[(2 * tig + 1) * SPITCH + 16 * warp + 8 * mt + grp] = acc[mt][1];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = tid; i < SR * 128; i += 256) Y[(int64_t)(i >> 7) * np + J + (i & 127)] = sy[(i >> 7) * SPITCH + (i & 127)];
    // rows below: warp w takes 32 rows of this CTA's chunk
    const int64_t row0 = J + 128 + (int64_t)blockIdx.x * SCHUNK + 32 * warp;
    if (row0 >= np) return;
    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    skinny_mma<4, false>(L + row0 * ld + J, ld, sy, acc, lane);
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        const int64_t row = row0 + 8 * mt + grp;
        R[(int64_t)(2 * tig) * np + row] -= acc[mt][0];
        R[(int64_t)(2 * tig + 1) * np + row] -= acc[mt][1];
    }
}

// backward step j: x_j = Dinv_j' y_j;  y[c] -= sum_{i in J} L[i][c] x_j[i] for all columns c left of block j
// (y lives in Y and is updated in place left of block j; x_j goes to R)
__global__ void __launch_bounds__(256) k_solve_bwd_step(const double *__restrict__ L, int64_t ld, const double *__restrict__ dinv,
                                                        double *__restrict__ R, double *__restrict__ Y, int64_t np, int j) {
    __shared__ double sy[SR * SPITCH];
    __shared__ double sx[SR * SPITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    const int64_t J = (int64_t)j * 128;
    for (int i = tid; i < SR * 128; i += 256) sy[(i >> 7) * SPITCH + (i & 127)] = Y[(int64_t)(i >> 7) * np + J + (i & 127)];
    __syncthreads();
    {   // x = Dinv_j' y: A[m][k] = Dinv[k][m]
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        skinny_mma<2, true>(dinv + J * 128 + 16 * warp, 128, sy, acc, lane);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            sx[(2 * tig) * SPITCH + 16 * warp + 8 * mt + grp] = acc[mt][0];
            sx[(2 * tig + 1) * SPITCH + 16 * warp + 8 * mt + grp] = acc[mt][1];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = tid; i < SR * 128; i += 256) R[(int64_t)(i >> 7) * np + J + (i & 127)] = sx[(i >> 7) * SPITCH + (i & 127)];
    if (j == 0) return;
    // columns left of block j: A[m = column][k = row of block J] = L[J + k][c]
    const int64_t c0 = (int64_t)blockIdx.x * SCHUNK + 32 * warp;
    if (c0 >= J) return;
    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    skinny_mma<4, true>(L + J * ld + c0, ld, sx, acc, lane);
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        const int64_t c = c0 + 8 * mt + grp;
        Y[(int64_t)(2 * tig) * np + c] -= acc[mt][0];
        Y[(int64_t)(2 * tig + 1) * np + c] -= acc[mt][1];
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

// ---- the same skinny solves, one 128-block at a time and COLUMN-oriented on the way back (owner-only storage) ------------------
// With the factor streamed panel by panel only the block COLUMN j of L is at hand when block j is processed (a virtual panel base:
// L(r, c) = L[r * ld + c] for c inside the panel).  Forward is right-looking already (k_solve_fwd_step reads column block j).  The
// backward step of launch_solve_rows8 reads the block ROW j of L, which spans every earlier panel; here it is left-looking instead:
//   x_j = Dinv_j' (y_j - sum_{r below block j} L[r, J]' x[r])
// -- partial sums per 256-row chunk (one CTA each), then one CTA adds them in chunk order (bitwise reproducible: every rank solves
// the same right-hand sides and must take the same decisions from them) and applies Dinv_j'.
__global__ void __launch_bounds__(256) k_solve_bwd_col_partial(const double *__restrict__ L, int64_t ld, const double *__restrict__ R,
                                                               int64_t np, int j, double *__restrict__ partial) {
    __shared__ double sx[SR * SPITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    const int64_t J = (int64_t)j * 128;
    const int64_t row0 = J + 128 + (int64_t)blockIdx.x * SCHUNK;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int slab = 0; slab < SCHUNK / 128; slab++) {
        const int64_t r0 = row0 + (int64_t)slab * 128;
        if (r0 >= np) break;                                   // block-uniform
        __syncthreads();
        for (int i = tid; i < SR * 128; i += 256) sx[(i >> 7) * SPITCH + (i & 127)] = R[(int64_t)(i >> 7) * np + r0 + (i & 127)];
        __syncthreads();
        // A[m = column of block j][k = row r0 + k] = L[r0 + k][J + m]; warp w: columns 16 w .. 16 w + 15
        skinny_mma<2, true>(L + r0 * ld + J + 16 * warp, ld, sx, acc, lane);
    }
    double *out = partial + (size_t)blockIdx.x * SR * 128;
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        out[(2 * tig) * 128 + 16 * warp + 8 * mt + grp] = acc[mt][0];
        out[(2 * tig + 1) * 128 + 16 * warp + 8 * mt + grp] = acc[mt][1];
    }
}

__global__ void __launch_bounds__(256) k_solve_bwd_col_finish(const double *__restrict__ dinv, double *__restrict__ R, const double *__restrict__ Y,
                                                              int64_t np, int j, const double *__restrict__ partial, int nchunk) {
    __shared__ double sy[SR * SPITCH];
    __shared__ double sxx[SR * SPITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    const int64_t J = (int64_t)j * 128;
    for (int i = tid; i < SR * 128; i += 256) {
        double s = 0.0;
        for (int c = 0; c < nchunk; c++) s += partial[(size_t)c * SR * 128 + i];      // fixed order
        sy[(i >> 7) * SPITCH + (i & 127)] = Y[(int64_t)(i >> 7) * np + J + (i & 127)] - s;
    }
    __syncthreads();
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    skinny_mma<2, true>(dinv + J * 128 + 16 * warp, 128, sy, acc, lane);      // x = Dinv_j' t
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        sxx[(2 * tig) * SPITCH + 16 * warp + 8 * mt + grp] = acc[mt][0];
        sxx[(2 * tig + 1) * SPITCH + 16 * warp + 8 * mt + grp] = acc[mt][1];
    }
    __syncthreads();
    for (int i = tid; i < SR * 128; i += 256) R[(int64_t)(i >> 7) * np + J + (i & 127)] = sxx[(i >> 7) * SPITCH + (i & 127)];
}

void launch_solve_fwd_block(const double *L, int64_t ld, const double *dinv, double *R, double *Y, int64_t np, int j, cudaStream_t s) {
    const int64_t below = np - (int64_t)(j + 1) * 128;
    const int grid = (int)std::max<int64_t>(1, (below + SCHUNK - 1) / SCHUNK);
    g_launch_count++;
    k_solve_fwd_step<<<grid, 256, 0, s>>>(L, ld, dinv, R, Y, np, j);
}

// partial: scratch of at least (np / 256 + 1) * 8 * 128 doubles
void launch_solve_bwd_block_col(const double *L, int64_t ld, const double *dinv, double *R, const double *Y, int64_t np, int j, double *partial,
                                cudaStream_t s) {
    const int64_t below = np - (int64_t)(j + 1) * 128;
    const int nchunk = (int)((below + SCHUNK - 1) / SCHUNK);
    if (nchunk > 0) {
        g_launch_count++;
        k_solve_bwd_col_partial<<<nchunk, 256, 0, s>>>(L, ld, R, np, j, partial);
    }
    g_launch_count++;
    k_solve_bwd_col_finish<<<1, 256, 0, s>>>(dinv, R, Y, np, j, partial, nchunk);
}

}  // namespace jaicov
