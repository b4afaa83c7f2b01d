// ozaki.cu -- the big tile products of the path (DEFAULT since round 2, 8 digits; JAICOV_GEMM_OZAKI=0 or jaicov_set_gemm_digits(0)
// switch back to the FP64 DMMA tiles of dense_kernels.cu): FP64 GEMM tiles built from int8 digit products on
// the 5th-generation tensor cores (tcgen05.mma kind::i8, operands staged by TMA, s32 accumulators in TMEM).
//
// Why: every O(n^3) flop of the path (Cholesky, inverse, the structured route's products) runs on k_gemm at 94 % DMMA pipe
// utilisation -- 96 % of the measured cuBLAS DGEMM rate.  tcgen05 has no f64 kind, but B200's INT8
// tensor rate is two orders of magnitude above its FP64 tensor rate, so an FP64 product can be assembled from exact
// integer digit products ("Ozaki scheme I"):
//   * every operand row is scaled by a power of two into (-1, 1) and cut into S signed digits |q| <= 64 (6 + 7 (S-1) bits);
//   * digit products q_i[k] * q_j[k] are summed over k EXACTLY in s32 (|sum| <= pairs * K * 4096 < 2^31), one TMEM
//     accumulator per digit-sum group g = i + j; pairs with i + j > S + 1 are below the target accuracy and dropped;
//   * the epilogue combines the groups in FP64 (Horner in 2^-7) and applies the row / column exponents, alpha and beta.
// With S = 8 (36 digit products) the results are FP64-equivalent: tests/test_ozaki_emulation.py runs the product's own
// Cholesky + inverse schedule with this arithmetic emulated bit for bit on the host (tests/emul/host_backend.cpp) and gets
// the cofactor matrix of real bundle networks as close to a long-double reference as the FP64 schedule itself.
//
// STATUS (round 2): validated and measured on B200 -- tools/ozaki_gpu_check.py complete (every operand layout /
// triangular hint / symmetric output with NaN-poisoned operands, SPD solve + inverse: errors exactly the host emulation's), the whole
// parity suite incl. the full-size configs 3 / 4 and the multi-GPU cases with EVERY launch forced onto this path
// (JAICOV_OZAKI_MIN_TILES=1 JAICOV_OZAKI_MIN_K=128), 29 / 36 / 42 / 45 TFLOP/s FP64-equivalent at K = 512 / 1024 / 2048 / 4096 against 33-35
// for the DMMA tiles (hence the thresholds below: launches of >= 148 tiles with K >= 1024), config 5 dense 7.40 s -> 5.39 s per final
// pass.  ncu: tcgen05 pipe 40 % active, bound by L2 -> shared-memory operand traffic (4.8 TB/s) of the 128 x 64 tile; the two-CTA
// cluster variant with TMA multicast of op(A) measures the same on whole passes and stays off (JAICOV_OZAKI_CLUSTER).
//
// Shapes: CTA tile 128 x 64 (S accumulators of 64 TMEM columns = all 512 columns for S = 8), k-block 64 bytes (SWIZZLE_64B
// rows), two smem stages of S * (128 + 64) * 64 bytes (96 KB each for S = 8); warp 0 = TMA producer, warp 1 = MMA issuer and
// TMEM owner, warps 2..5 = epilogue (one TMEM lane quarter each).  One CTA per SM; the grid is the tile list of the launch in
// the same order as k_gemm (longest contractions first, lower tiles only for symmetric outputs).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "dense_driver.hpp"

namespace jaicov {

// ---- digit extraction ----------------------------------------------------------------------------------------------------
// valid k-range of the rows of row tile t (rows 128 t .. 128 t + 127) of an operand; everything outside becomes zero digits
enum OzRange : int { OZ_FULL = 0, OZ_FROM_TILE = 1 /* k >= 128 t */, OZ_UPTO_TILE = 2 /* k < 128 (t + 1) */ };

struct OzSplitArgs {
    const double *X;
    int64_t ld;
    int layout;          // 0: X(row, k) at X[row * ld + k];  1: at X[k * ld + row]
    int64_t rows, K;     // multiples of 128
    int64_t row_min;     // rows below this one are not wanted by any tile of the launch: neither read nor split
    int range;           // OzRange
    int S;
    int8_t *Q;           // [S][rows][K]
    int32_t *e;          // [rows]: |x| < 2^e over the row's valid range (0 for an all-zero row)
    unsigned long long *amax;   // [rows]: bit pattern of max |x| (non-negative doubles order like unsigned integers)
    const int32_t *fk;   // [K] or null: contraction-index balancing, the operand is read as x * 2^(fsign * fk[k]) (see k_oz_fk)
    int fsign;
    unsigned long long *cmax;   // [K]: bit pattern of the column maxima (k_oz_colmax)
};

__device__ __forceinline__ bool oz_valid(const OzSplitArgs &a, int64_t row, int64_t k) {
    const int64_t t = row >> 7;
    if (a.range == OZ_FROM_TILE) return k >= (t << 7);
    if (a.range == OZ_UPTO_TILE) return k < ((t + 1) << 7);
    return true;
}

// thread <-> (row, 16 consecutive k).  layout 0: a warp covers one row and 512 consecutive k (contiguous 4 KB read);
// layout 1: a warp covers 32 consecutive rows of one 16-wide k chunk (every k is one coalesced 256-byte read).
__device__ __forceinline__ bool oz_map(const OzSplitArgs &a, int64_t &row, int64_t &k0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.layout == 0) {
        row = (int64_t)blockIdx.y * 4 + warp;
        k0 = ((int64_t)blockIdx.x * 32 + lane) * 16;
    } else {
        row = (int64_t)blockIdx.y * 128 + threadIdx.x;
        k0 = (int64_t)blockIdx.x * 16;
    }
    return row >= a.row_min && row < a.rows && k0 < a.K;
}

__device__ __forceinline__ void oz_load16(const OzSplitArgs &a, int64_t row, int64_t k0, double (&x)[16]) {
    if (a.layout == 0) {
        const double2 *p = reinterpret_cast<const double2 *>(a.X + row * a.ld + k0);
#pragma unroll
        for (int i = 0; i < 8; i++) { const double2 v = p[i]; x[2 * i] = v.x; x[2 * i + 1] = v.y; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = a.X[(k0 + i) * a.ld + row];
    }
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (!oz_valid(a, row, k0 + i)) x[i] = 0.0;
    if (a.fk) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] *= __longlong_as_double((long long)(1023 + a.fsign * a.fk[k0 + i]) << 52);   // exact, |f| <= 1023
    }
}

// Contraction-index balancing (tests/emul/host_backend.cpp: gemm_ozaki; tests/ozaki_study.py --structured): column maxima of both
// operands, then f_k = floor((exponent of max |B[., k]| - exponent of max |A[., k]|) / 2).  op(A)[., k] * 2^f_k and op(B)[., k] * 2^-f_k
// leave every product unchanged but even out the magnitudes inside the operand rows, which the per-row digit grid resolves.
__global__ void __launch_bounds__(128) k_oz_colmax(OzSplitArgs a) {
    double m = 0.0;
    int64_t k;
    if (a.layout == 0) {          // thread <-> one k, a slab of 256 rows (every row access is one coalesced 1 KB read of the block)
        k = (int64_t)blockIdx.x * 128 + threadIdx.x;
        const int64_t r0 = max(a.row_min, (int64_t)blockIdx.y * 256), r1 = min(a.rows, (int64_t)blockIdx.y * 256 + 256);
        if (k < a.K)
            for (int64_t r = r0; r < r1; r++)
                if (oz_valid(a, r, k)) m = fmax(m, fabs(a.X[r * a.ld + k]));
        if (k < a.K && m > 0.0) atomicMax(a.cmax + k, (unsigned long long)__double_as_longlong(m));
    } else {                      // warp <-> one k, lanes stride over a slab of 4096 rows (contiguous in memory)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        k = (int64_t)blockIdx.x * 4 + warp;
        const int64_t r0 = max(a.row_min, (int64_t)blockIdx.y * 4096), r1 = min(a.rows, (int64_t)blockIdx.y * 4096 + 4096);
        if (k < a.K)
            for (int64_t r = r0 + lane; r < r1; r += 32)
                if (oz_valid(a, r, k)) m = fmax(m, fabs(a.X[k * a.ld + r]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && k < a.K && m > 0.0) atomicMax(a.cmax + k, (unsigned long long)__double_as_longlong(m));
    }
}

__global__ void __launch_bounds__(256) k_oz_fk(const unsigned long long *ca, const unsigned long long *cb, int64_t K, int32_t *fk) {
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const int ea = (int)((ca[k] >> 52) & 0x7ff), eb = (int)((cb[k] >> 52) & 0x7ff);
    const int f = (ea == 0 || eb == 0 || ea == 2047 || eb == 2047) ? 0 : ((eb - ea) >> 1);   // zero / subnormal / non-finite column: leave it alone
    fk[k] = max(-1000, min(1000, f));                                                      // 2^(+-f) stays a normal number
}

static void launch_colmax(const OzSplitArgs &a, cudaStream_t s) {
    JCHECK(cudaMemsetAsync(a.cmax, 0, (size_t)a.K * sizeof(unsigned long long), s));
    const dim3 grid = a.layout == 0 ? dim3((unsigned)((a.K + 127) / 128), (unsigned)((a.rows + 255) / 256))
                                    : dim3((unsigned)((a.K + 3) / 4), (unsigned)((a.rows + 4095) / 4096));
    g_launch_count++;
    k_oz_colmax<<<grid, 128, 0, s>>>(a);
}

__global__ void __launch_bounds__(128) k_oz_rowmax(OzSplitArgs a) {
    int64_t row, k0;
    const bool on = oz_map(a, row, k0);
    double m = 0.0;
    if (on) {
        double x[16];
        oz_load16(a, row, k0, x);
#pragma unroll
        for (int i = 0; i < 16; i++) m = fmax(m, fabs(x[i]));
    }
    if (a.layout == 0) {   // the warp shares one row
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0 && on && m > 0.0) atomicMax(a.amax + row, (unsigned long long)__double_as_longlong(m));
    } else if (on && m > 0.0) {
        atomicMax(a.amax + row, (unsigned long long)__double_as_longlong(m));
    }
}

__global__ void __launch_bounds__(128) k_oz_digits(OzSplitArgs a) {
    int64_t row, k0;
    if (!oz_map(a, row, k0)) return;
    const unsigned long long bits = a.amax[row];
    const int ef = (int)((bits >> 52) & 0x7ff);               // biased exponent of the row maximum
    const bool zero = ef < 5 || ef > 2046;                    // all-zero row, or too small to scale / not finite: zero digits
    if (k0 == 0) a.e[row] = zero ? 0 : ef - 1022;             // |x| < 2^(ef - 1022)
    double t[16];
    if (!zero) {
        oz_load16(a, row, k0, t);
        const double sc = __longlong_as_double((long long)(2051 - ef) << 52);   // 2^(6 - e): |t| < 64
#pragma unroll
        for (int i = 0; i < 16; i++) t[i] *= sc;
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) t[i] = 0.0;
    }
    for (int j = 0; j < a.S; j++) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int q = __double2int_rn(t[i]);              // |q| <= 64
            t[i] = (t[i] - (double)q) * 128.0;                // exact: |t - q| <= 1/2
            w[i >> 2] |= (uint32_t)(q & 0xff) << (8 * (i & 3));
        }
        *reinterpret_cast<uint4 *>(a.Q + ((size_t)j * a.rows + row) * a.K + k0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

static void launch_split(const OzSplitArgs &a, cudaStream_t s) {
    JCHECK(cudaMemsetAsync(a.amax, 0, (size_t)a.rows * sizeof(unsigned long long), s));
    const dim3 grid = a.layout == 0 ? dim3((unsigned)((a.K + 511) / 512), (unsigned)(a.rows / 4))
                                    : dim3((unsigned)(a.K / 16), (unsigned)(a.rows / 128));
    g_launch_count += 2;
    k_oz_rowmax<<<grid, 128, 0, s>>>(a);
    k_oz_digits<<<grid, 128, 0, s>>>(a);
}

// ---- the tile kernel -----------------------------------------------------------------------------------------------------
// digits used when neither jaicov_set_gemm_digits nor JAICOV_GEMM_OZAKI says otherwise (0 = FP64 DMMA tiles everywhere).
// 8 since round 2: parity-green on every GPU test with ALL launches on this path (1 and 2 GPUs) and 1.39x faster than the DMMA
// tiles on the dense route at config 5 (profiles/r02_ozaki_*)
constexpr int kOzakiDefaultDigits = 8;
constexpr int OZ_BM = 128, OZ_BN = 64, OZ_BK = 64, OZ_STAGES = 2, OZ_THREADS = 192, OZ_TMEM_COLS = 512;

template <int S> struct OzCfg {
    static constexpr int A_STAGE = S * OZ_BM * OZ_BK;   // bytes: [S][128 rows][64 B]
    static constexpr int B_STAGE = S * OZ_BN * OZ_BK;   //        [S][ 64 rows][64 B]
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int SMEM = OZ_STAGES * STAGE + 1024 /* alignment slack */ + 128 /* barriers, TMEM slot */;
    static_assert(S >= 2 && S * OZ_BN <= OZ_TMEM_COLS, "one 64-column accumulator per digit-sum group");
    static_assert(SMEM <= 227 * 1024, "two stages must fit the 227 KB of one CTA");
};

struct OzGemmArgs {
    const int32_t *ea, *eb;   // row exponents of op(A) / op(B)
    double *C;
    int64_t ldc, K;
    double alpha, beta;
    int mt, nt, tri_out, tile_band, kmode;
    // launches of the multi-GPU schedule (dense_driver.hpp): column table (trapezoid update of many panels at once) and the
    // per-column-tile masks of the column-panel inverse
    const int32_t *coltab;    // 2-D launch: blockIdx.y picks the global column tile coltab[y] / 128
    int coltab_full, c_local;
    const int32_t *ktab;      // K_COL_BEG / K_ROW_MASK: first global row of interest of every column tile
    int64_t koff, roff;
};

__device__ __forceinline__ uint32_t oz_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void oz_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_tma_load_3d(const CUtensorMap *map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// the same copy delivered to every CTA of the cluster named in `mask` (same smem offset, same barrier offset in each)
__device__ __forceinline__ void oz_tma_load_3d_mc(const CUtensorMap *map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void oz_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major operand rows of 64 bytes, SWIZZLE_64B: 8-row groups are 512 bytes apart (cute::UMMA::SmemDescriptor, sm_100 version 1)
__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);       // start address >> 4                      bits [0, 14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused, swizzled)  bits [16, 30)
    d |= (uint64_t)(512 >> 4) << 32;                // stride byte offset: 8 rows x 64 B       bits [32, 46)
    d |= (uint64_t)1 << 46;                         // descriptor version of sm_100            bits [46, 48)
    d |= (uint64_t)4 << 61;                         // layout type SWIZZLE_64B                 bits [61, 64)
    return d;
}
// D[tmem] (+)= A[smem] * B[smem]', s8 x s8 -> s32, issued by one thread for the whole CTA
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)   // no lane disabled
        : "memory");
}
__device__ __forceinline__ void oz_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void oz_commit_mc(uint32_t bar, uint16_t mask) {   // arrives on the barrier at this offset in every CTA of `mask`
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void oz_tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    // load and wait in ONE statement: the registers are only defined once the wait has retired, and nothing can be scheduled between
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}

// CL = 1: every CTA loads its own operands.  CL = 2 (JAICOV_OZAKI_CLUSTER=2): the two 64-column halves of one 128 x 128 tile
// form a thread-block cluster and share op(A): each CTA fetches 64 of the 128 rows of every digit plane and TMA multicasts
// them into both CTAs' shared memory, which cuts the L2 -> SM traffic per stage from 96 KB to 64 KB (S = 8).  A stage may
// then only be refilled once BOTH issuers have released it, so the "empty" barriers count two multicast commits.
template <int S, int CL>
__global__ void __launch_bounds__(OZ_THREADS, 1)
k_gemm_oz(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, OzGemmArgs p) {
    using Cfg = OzCfg<S>;
    extern __shared__ uint8_t oz_raw[];
    const uint32_t raw = oz_smem_u32(oz_raw);
    uint8_t *sm = oz_raw + (((raw + 1023u) & ~1023u) - raw);           // swizzle atoms need their natural alignment
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + OZ_STAGES * Cfg::STAGE);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    const uint32_t full0 = oz_smem_u32(bars), empty0 = oz_smem_u32(bars + OZ_STAGES), accum = oz_smem_u32(bars + 2 * OZ_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile decode: the same order as k_gemm, every 128 x 128 tile as two 64-column halves ---------------------------
    // (tiles that are not wanted leave here, before any barrier or TMEM allocation exists; both CTAs of a cluster share the tile)
    const int l = blockIdx.x >> 1, half = blockIdx.x & 1;     // CL = 2: half == rank of the CTA in its cluster
    int it, jt;
    if (p.coltab) {
        it = l;
        jt = p.coltab[blockIdx.y] / 128;
        if (!p.coltab_full && it < jt) return;                // above the diagonal of this column tile
    } else if (p.tri_out) {
        tri_tile_decode(l, p.mt, p.tile_band, it, jt);
    } else {
        // bands of tile_band tile rows, column by column inside a band: the CTAs resident together (one per SM, ~74 tiles) then
        // cover a roughly square patch of C and share ~band + 74 / band operand strips instead of 1 + 64 -- the row-by-row order
        // re-read op(B) from DRAM once per tile row (ncu: 32 GB of DRAM reads for 1.07 GB of operands at 8192^3, tensor pipe 40 %
        // active; profiles/r02_ncu_full_k_gemm_oz_summary.txt)
        rect_tile_decode(l, p.mt, p.nt, p.tile_band, it, jt);
        if (p.kmode == K_A_LOWER) it = p.mt - 1 - it;
    }
    int64_t kbeg = 0, kend = p.K;
    if (p.kmode == K_B_LOWER) kbeg = (int64_t)jt * 128;
    else if (p.kmode == K_A_LOWER) kend = min(p.K, (int64_t)(it + 1) * 128);
    else if (p.kmode == K_MAX_IJ) kbeg = (int64_t)max(it, jt) * 128;
    else if (p.kmode == K_COL_BEG) {
        kbeg = max((int64_t)0, (int64_t)p.ktab[jt] - p.koff);
        if (kbeg >= p.K) return;                              // this column tile is still zero in the whole contraction range
    } else if (p.kmode == K_ROW_MASK) {
        if (p.roff + (int64_t)it * 128 < (int64_t)p.ktab[jt]) return;   // above the column's diagonal: not wanted
    }
    const int nkb = (int)((kend - kbeg) / OZ_BK);
    const int m0 = it * 128, n0 = jt * 128 + half * OZ_BN;
    const int64_t c0 = ((p.coltab && p.c_local) ? (int64_t)blockIdx.y : (int64_t)jt) * 128 + half * OZ_BN;   // first column of C

    // ---- one-time setup --------------------------------------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
        for (int i = 0; i < OZ_STAGES; i++) {
            oz_mbar_init(full0 + 8 * i, 1);
            oz_mbar_init(empty0 + 8 * i, CL);
        }
        oz_mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tmem_slot)), "r"(OZ_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) oz_cluster_sync();                 // the peer's barriers exist before anything is multicast at them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp == 0) {
        // ===== TMA producer: one elected lane, S digit planes of A and of B per stage as two 3-D boxes =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int st = kb % OZ_STAGES;
                const uint32_t ph = (uint32_t)(kb / OZ_STAGES) & 1u;
                oz_mbar_wait(empty0 + 8 * st, ph ^ 1u);
                oz_mbar_expect_tx(full0 + 8 * st, (uint32_t)Cfg::STAGE);
                const uint32_t dst = oz_smem_u32(sm + st * Cfg::STAGE);
                const int k0 = (int)(kbeg + (int64_t)kb * OZ_BK);
                if (CL == 1) {
                    oz_tma_load_3d(&mapA, full0 + 8 * st, dst, k0, m0, 0);
                } else {
                    // this CTA's 64 rows of every digit plane, to both CTAs (mapA's box is 64 bytes x 64 rows x 1 plane here)
                    for (int pl = 0; pl < S; pl++)
                        oz_tma_load_3d_mc(&mapA, full0 + 8 * st, dst + pl * (OZ_BM * OZ_BK) + half * (64 * OZ_BK), k0, m0 + 64 * half, pl,
                                          (uint16_t)3);
                }
                oz_tma_load_3d(&mapB, full0 + 8 * st, dst + Cfg::A_STAGE, k0, n0, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: digit pair (i, j) accumulates into the TMEM accumulator of its group i + j =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = signed 8 bit, both K-major, N = 64, M = 128
            constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
            for (int kb = 0; kb < nkb; kb++) {
                const int st = kb % OZ_STAGES;
                const uint32_t ph = (uint32_t)(kb / OZ_STAGES) & 1u;
                oz_mbar_wait(full0 + 8 * st, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_base = oz_smem_u32(sm + st * Cfg::STAGE), b_base = a_base + Cfg::A_STAGE;
#pragma unroll 1
                for (int i = 0; i < S; i++)
#pragma unroll 1
                    for (int j = 0; i + j < S; j++)
#pragma unroll
                        for (int kk = 0; kk < OZ_BK / 32; kk++) {
                            const uint64_t ad = oz_smem_desc(a_base + i * (OZ_BM * OZ_BK) + kk * 32);
                            const uint64_t bd = oz_smem_desc(b_base + j * (OZ_BN * OZ_BK) + kk * 32);
                            oz_mma_i8(tmem + (uint32_t)((i + j) * OZ_BN), ad, bd, idesc, (kb > 0 || kk > 0 || i > 0) ? 1u : 0u);
                        }
                if (CL == 1) oz_commit(empty0 + 8 * st);        // frees the stage once these MMAs have read it
                else oz_commit_mc(empty0 + 8 * st, (uint16_t)3);   // ... in both CTAs: the peer writes half of our op(A)
            }
            oz_commit(accum);                      // all groups complete
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes 32 (warp % 4) .. +31 = rows of the tile =====
        const int q = warp & 3;
        oz_mbar_wait(accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t m = (int64_t)m0 + 32 * q + lane;
        const int ea = p.ea[m];
        double *crow = p.C + m * p.ldc + c0;
        const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
        for (int c8 = 0; c8 < OZ_BN / 8; c8++) {
            double r[8];
#pragma unroll
            for (int c = 0; c < 8; c++) r[c] = 0.0;
#pragma unroll 1
            for (int gi = S - 1; gi >= 0; gi--) {          // Horner in 2^-7, smallest contributions first
                uint32_t v[8];
                oz_tmem_ld8(trow + (uint32_t)(gi * OZ_BN + c8 * 8), v);
#pragma unroll
                for (int c = 0; c < 8; c++) r[c] = r[c] * 0.0078125 + (double)(int)v[c];
            }
            double out[8];
#pragma unroll
            for (int c = 0; c < 8; c++) out[c] = p.alpha * ldexp(r[c], ea + p.eb[n0 + c8 * 8 + c] - 12);   // digit weights 2^-(7 g - 2)
            double2 *dst = reinterpret_cast<double2 *>(crow + c8 * 8);
            if (p.beta != 0.0) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const double2 old = dst[c];
                    out[2 * c] += p.beta * old.x;
                    out[2 * c + 1] += p.beta * old.y;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) dst[c] = make_double2(out[2 * c], out[2 * c + 1]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (CL > 1) oz_cluster_sync();                 // no CTA leaves while its peer may still signal its barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(OZ_TMEM_COLS) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// digit planes Q[S][rows][K] (int8, K contiguous) as a 3-D tensor; one box = box_planes planes x box_rows rows x 64 bytes
bool make_map(CUtensorMap *map, const int8_t *Q, int64_t rows, int64_t K, int S, int box_rows, int box_planes) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t gstride[2] = {(cuuint64_t)K, (cuuint64_t)rows * (cuuint64_t)K};   // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t *>(Q), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct OzScratch {
    int8_t *q[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    int32_t *e[2] = {nullptr, nullptr};
    unsigned long long *amax[2] = {nullptr, nullptr};
    size_t rows_cap[2] = {0, 0};
    unsigned long long *cmax[2] = {nullptr, nullptr};
    int32_t *fk = nullptr;
    size_t k_cap = 0;
    bool ensure_k(size_t K) {
        if (K <= k_cap) return true;
        for (int w = 0; w < 2; w++) { if (cmax[w]) cudaFree(cmax[w]); cmax[w] = nullptr; }
        if (fk) cudaFree(fk);
        fk = nullptr; k_cap = 0;
        if (cudaMalloc(&cmax[0], K * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&cmax[1], K * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&fk, K * sizeof(int32_t)) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        k_cap = K;
        return true;
    }
    bool ensure(int w, size_t bytes, size_t rows) {
        if (bytes > cap[w]) {
            if (q[w]) cudaFree(q[w]);
            q[w] = nullptr; cap[w] = 0;
            if (cudaMalloc(&q[w], bytes) != cudaSuccess) { cudaGetLastError(); return false; }
            cap[w] = bytes;
        }
        if (rows > rows_cap[w]) {
            if (e[w]) cudaFree(e[w]);
            if (amax[w]) cudaFree(amax[w]);
            e[w] = nullptr; amax[w] = nullptr; rows_cap[w] = 0;
            if (cudaMalloc(&e[w], rows * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&amax[w], rows * sizeof(unsigned long long)) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            rows_cap[w] = rows;
        }
        return true;
    }
    size_t release() {
        size_t b = cap[0] + cap[1];
        for (int w = 0; w < 2; w++) {
            if (q[w]) cudaFree(q[w]);
            if (e[w]) cudaFree(e[w]);
            if (amax[w]) cudaFree(amax[w]);
            q[w] = nullptr; e[w] = nullptr; amax[w] = nullptr; cap[w] = rows_cap[w] = 0;
            if (cmax[w]) cudaFree(cmax[w]);
            cmax[w] = nullptr;
        }
        if (fk) cudaFree(fk);
        fk = nullptr; k_cap = 0;
        return b;
    }
};
// one scratch set per device ordinal (a handle works on one device; the single-process multi-GPU handle runs one host thread per
// device).  Launches of ONE stream reuse it safely (stream order); two adjustments running concurrently on the same device would
// share it -- the library documents one adjustment at a time per device.
OzScratch g_oz_dev[64];
std::atomic<int> g_oz_digits{-1};     // -1: not decided yet (JAICOV_GEMM_OZAKI, else the built-in default)

template <int S, int CL>
void launch_tiles(const CUtensorMap &ma, const CUtensorMap &mb, const OzGemmArgs &a, int64_t tiles, int grid_y, cudaStream_t s) {
    static PerDeviceOnce once;
    once.run([] { JCHECK(cudaFuncSetAttribute(k_gemm_oz<S, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzCfg<S>::SMEM)); });
    g_launch_count++;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles * 2), (unsigned)grid_y);   // x: row tile (or tile list) x two 64-column halves; y: column-table slot
    cfg.blockDim = dim3(OZ_THREADS);
    cfg.dynamicSmemBytes = OzCfg<S>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    JCHECK(cudaLaunchKernelEx(&cfg, k_gemm_oz<S, CL>, ma, mb, a));
}

template <int CL>
void launch_tiles_digits(int digits, const CUtensorMap &ma, const CUtensorMap &mb, const OzGemmArgs &a, int64_t tiles, int grid_y, cudaStream_t s) {
    switch (digits) {
        case 4: launch_tiles<4, CL>(ma, mb, a, tiles, grid_y, s); break;
        case 5: launch_tiles<5, CL>(ma, mb, a, tiles, grid_y, s); break;
        case 6: launch_tiles<6, CL>(ma, mb, a, tiles, grid_y, s); break;
        case 7: launch_tiles<7, CL>(ma, mb, a, tiles, grid_y, s); break;
        default: launch_tiles<8, CL>(ma, mb, a, tiles, grid_y, s); break;
    }
}

}  // namespace

size_t ozaki_release_scratch() {
    size_t b = 0;
    int cur = 0, n = 0;
    cudaGetDevice(&cur);
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int d = 0; d < n && d < 64; d++)
        if (g_oz_dev[d].cap[0] || g_oz_dev[d].cap[1] || g_oz_dev[d].k_cap || g_oz_dev[d].rows_cap[0]) { cudaSetDevice(d); b += g_oz_dev[d].release(); }
    cudaSetDevice(cur);
    return b;
}

// digits of the int8-digit tile product (0 = FP64 DMMA tiles for every launch); returns the previous setting
int ozaki_set_digits(int digits) {
    if (digits < 0) return g_oz_digits.load();
    return g_oz_digits.exchange((digits >= 4 && digits <= 8) ? digits : 0);
}

int ozaki_digits() {
    int d = g_oz_digits.load();
    if (d < 0) {
        const char *e = getenv("JAICOV_GEMM_OZAKI");
        d = e ? atoi(e) : kOzakiDefaultDigits;
        if (d < 4 || d > 8) d = 0;
        g_oz_digits.store(d);
    }
    return d;
}

// Takes the launch if the digit path is on (default) and the launch is one of the big plain / triangular-operand products;
// returns false (nothing launched) otherwise, and the caller runs k_gemm.
bool launch_gemm_ozaki(const GemmDesc &g, cudaStream_t s) {
    const int digits = ozaki_digits();
    if (digits < 4 || digits > 8) return false;
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (dev_ < 0 || dev_ >= 64) return false;
    OzScratch &g_oz = g_oz_dev[dev_];
    static const int64_t min_tiles = [] { const char *e = getenv("JAICOV_OZAKI_MIN_TILES"); return e ? (int64_t)atoll(e) : (int64_t)148; }();
    // column-table launches: only the trapezoid update of the distributed Cholesky (both operands the same panel, C addressed by
    // global tiles, or by compact own tiles with owner-only storage); the structured route's column tables keep the FP64 kernel
    const bool trapezoid = g.coltab && !g.coltab_full && g.A == g.B && g.lda == g.ldb && g.al == g.bl && g.kmode == K_FULL;
    if ((g.coltab && !trapezoid) || g.kmode > K_ROW_MASK) return false;
    const int64_t tiles = g.coltab ? (int64_t)g.mt * g.ncoltab : (g.tri_out ? (int64_t)g.mt * (g.mt + 1) / 2 : (int64_t)g.mt * g.nt);
    // short contractions do not pay for the digit pre-pass (six small launches and 8 + S bytes per operand element)
    static const int64_t min_k = [] { const char *e = getenv("JAICOV_OZAKI_MIN_K"); return e ? (int64_t)atoll(e) : (int64_t)1024; }();
    if (tiles < min_tiles || g.K < std::max<int64_t>(128, min_k)) return false;
    if ((double)g.K * digits * 4096.0 >= 2147483648.0) return false;     // a digit-sum group must fit its s32 accumulator
    const int64_t Mr = (int64_t)g.mt * 128, Nr = g.coltab ? (int64_t)g.mt * 128 : (int64_t)g.nt * 128;   // column table: op(B) rows by global tile
    const int ra = g.kmode == K_MAX_IJ ? OZ_FROM_TILE : (g.kmode == K_A_LOWER ? OZ_UPTO_TILE : OZ_FULL);
    const int rb = (g.kmode == K_MAX_IJ || g.kmode == K_B_LOWER) ? OZ_FROM_TILE : OZ_FULL;
    const bool shared = g.A == g.B && g.lda == g.ldb && g.al == g.bl && ra == rb && Mr == Nr;
    if (!g_oz.ensure(0, (size_t)digits * Mr * g.K, (size_t)Mr)) return false;
    if (!shared && !g_oz.ensure(1, (size_t)digits * Nr * g.K, (size_t)Nr)) return false;
    CUtensorMap ma, mb;
    const int wb = shared ? 0 : 1;
    static const int cluster = [] { const char *e = getenv("JAICOV_OZAKI_CLUSTER"); return (e && atoi(e) == 2) ? 2 : 1; }();
    const bool mapped = cluster == 2 ? make_map(&ma, g_oz.q[0], Mr, g.K, digits, 64, 1)            // half of op(A)'s rows, one plane per copy
                                     : make_map(&ma, g_oz.q[0], Mr, g.K, digits, OZ_BM, digits);
    if (!mapped || !make_map(&mb, g_oz.q[wb], Nr, g.K, digits, OZ_BN, digits)) return false;
    // contraction-index balancing is on unless JAICOV_OZAKI_KSCALE=0; it has nothing to do when both operands are one array
    static const bool kscale = [] { const char *e = getenv("JAICOV_OZAKI_KSCALE"); return !(e && atoi(e) == 0); }();
    const bool balance = kscale && !shared;
    if (balance && !g_oz.ensure_k((size_t)g.K)) return false;
    const int64_t rmin = g.coltab ? (g.row_min / 128) * 128 : 0;
    OzSplitArgs sa{g.A, g.lda, g.al, Mr, g.K, rmin, ra, digits, g_oz.q[0], g_oz.e[0], g_oz.amax[0], nullptr, 1, g_oz.cmax[0]};
    OzSplitArgs sb{g.B, g.ldb, g.bl, Nr, g.K, rmin, rb, digits, g_oz.q[1], g_oz.e[1], g_oz.amax[1], nullptr, -1, g_oz.cmax[1]};
    if (balance) {
        launch_colmax(sa, s);
        launch_colmax(sb, s);
        g_launch_count++;
        k_oz_fk<<<(unsigned)((g.K + 255) / 256), 256, 0, s>>>(g_oz.cmax[0], g_oz.cmax[1], g.K, g_oz.fk);
        sa.fk = sb.fk = g_oz.fk;
    }
    launch_split(sa, s);
    if (!shared) launch_split(sb, s);
    static const int oz_band = [] { const char *e = getenv("JAICOV_OZAKI_BAND"); return e ? atoi(e) : 8; }();
    OzGemmArgs a{g_oz.e[0], g_oz.e[wb], g.C, g.ldc, g.K, g.alpha, g.beta, g.mt, g.nt, g.tri_out, g.tile_band > 0 ? g.tile_band : oz_band, g.kmode,
                 g.coltab, g.coltab_full, g.c_local, g.ktab, g.koff, g.roff};
    const int64_t grid_tiles = g.coltab ? (int64_t)g.mt : tiles;     // column table: x runs over the row tiles, y over the table
    const int grid_y = g.coltab ? g.ncoltab : 1;
    if (cluster == 2) launch_tiles_digits<2>(digits, ma, mb, a, grid_tiles, grid_y, s);
    else launch_tiles_digits<1>(digits, ma, mb, a, grid_tiles, grid_y, s);
    return true;
}

}  // namespace jaicov
