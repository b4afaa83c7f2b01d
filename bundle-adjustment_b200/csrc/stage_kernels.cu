// stage_kernels.cu -- the O(n^2) and O(n) stages around the dense factorisation (K3, K4, K7, K9 of SURVEY.md 2.1):
// Jacobi preconditioner (BA:824-828, NES:82-91), datum condition rows (BA:493-635), the SPD reformulation of the
// bordered system and its rank-d correction, Qxx un-scaling (BA:273), parameter update + max|dx| (BA:450-462),
// directly observed parameter groups (PDF:447-473, DOPG:67-91) and packing into MTJ layout.  All HBM-bound.
#include "common.h"

namespace jaicov {

__device__ __forceinline__ bool col_active_s(int32_t c) { return c >= 0 && c != JAICOV_COL_FIXED; }
__device__ __forceinline__ int64_t lower_idx_s(int64_t a, int64_t b, int64_t ld) { return a >= b ? a * ld + b : b * ld + a; }

// V = diag(N)^-1/2 (1 where diag <= EPS), BA:824-828; padding rows get 1
// (owner-only storage: a rank sees only its own diagonal entries; the others get 0 here and the ranks' vectors are summed --
//  api.cu all-reduces V right after this launch)
__global__ void k_precond_diag(SysView M, int u, int64_t np, double *__restrict__ V) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= np) return;
    const double *q = sys_at(M, e, e);
    double v = q ? 1.0 : 0.0;
    if (e < u && q) {
        const double x = *q;
        v = x > kEps ? 1.0 / sqrt(x) : 1.0;
    }
    V[e] = v;
}

void launch_precond_diag(const SysView &M, int u, int64_t np, double *V, cudaStream_t s) {
    g_launch_count++;
    k_precond_diag<<<(unsigned)((np + 255) / 256), 256, 0, s>>>(M, u, np, V);
}

// Datum rows (addDatumConditionRows, BA:493-635): Bt[a][e], a = condition, e = internal unknown index; Bt must be
// zero on entry.  free_mask bit k: 0 tx,1 ty,2 tz,3 rx,4 ry,5 rz,6 scale.
__global__ void __launch_bounds__(256) k_datum_rows(const double *__restrict__ xyz, const int32_t *__restrict__ pt_col,
                                                    const int32_t *__restrict__ datum_pts, int nDatum, int free_mask,
                                                    int d, int64_t np, double *__restrict__ Bt) {
    __shared__ double red[256][3];
    __shared__ double s_c[3];
    __shared__ double s_norm[7];
    const int tid = threadIdx.x;
    double sx = 0, sy = 0, sz = 0;
    for (int i = tid; i < nDatum; i += 256) {
        const int p = datum_pts[i];
        sx += xyz[3 * (int64_t)p]; sy += xyz[3 * (int64_t)p + 1]; sz += xyz[3 * (int64_t)p + 2];
    }
    red[tid][0] = sx; red[tid][1] = sy; red[tid][2] = sz;
    __syncthreads();
    if (tid < 3) {
        double s = 0.0;
        for (int i = 0; i < 256; i++) s += red[i][tid];
        s_c[tid] = s / (double)nDatum;
    }
    __syncthreads();
    const double x0 = s_c[0], y0 = s_c[1], z0 = s_c[2];
    int rowOf[7], k = 0;
#pragma unroll
    for (int b = 0; b < 7; b++) rowOf[b] = (free_mask >> b) & 1 ? k++ : -1;
    double nr[3] = {0, 0, 0}, ns = 0;  // rotation rows and scale row norms (translation rows: count)
    for (int i = tid; i < nDatum; i += 256) {
        const int p = datum_pts[i];
        const double x = xyz[3 * (int64_t)p] - x0, y = xyz[3 * (int64_t)p + 1] - y0, z = xyz[3 * (int64_t)p + 2] - z0;
        nr[0] += z * z + y * y; nr[1] += z * z + x * x; nr[2] += x * x + y * y; ns += x * x + y * y + z * z;
    }
    __syncthreads();
    red[tid][0] = nr[0]; red[tid][1] = nr[1]; red[tid][2] = nr[2];
    __syncthreads();
    if (tid < 3) {
        double s = 0.0;
        for (int i = 0; i < 256; i++) s += red[i][tid];
        s_norm[3 + tid] = s;
        s_norm[tid] = (double)nDatum;
    }
    __syncthreads();
    red[tid][0] = ns;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < 256; i++) s += red[i][0];
        s_norm[6] = s;
    }
    __syncthreads();
    for (int i = tid; i < nDatum; i += 256) {
        const int p = datum_pts[i];
        const int64_t cX = pt_col[3 * (int64_t)p] - d, cY = pt_col[3 * (int64_t)p + 1] - d, cZ = pt_col[3 * (int64_t)p + 2] - d;
        const double x = xyz[3 * (int64_t)p] - x0, y = xyz[3 * (int64_t)p + 1] - y0, z = xyz[3 * (int64_t)p + 2] - z0;
        if (rowOf[0] >= 0) Bt[rowOf[0] * np + cX] = 1.0 / sqrt(s_norm[0]);
        if (rowOf[1] >= 0) Bt[rowOf[1] * np + cY] = 1.0 / sqrt(s_norm[1]);
        if (rowOf[2] >= 0) Bt[rowOf[2] * np + cZ] = 1.0 / sqrt(s_norm[2]);
        if (rowOf[3] >= 0) { Bt[rowOf[3] * np + cY] = z / sqrt(s_norm[3]); Bt[rowOf[3] * np + cZ] = -y / sqrt(s_norm[3]); }
        if (rowOf[4] >= 0) { Bt[rowOf[4] * np + cX] = -z / sqrt(s_norm[4]); Bt[rowOf[4] * np + cZ] = x / sqrt(s_norm[4]); }
        if (rowOf[5] >= 0) { Bt[rowOf[5] * np + cX] = y / sqrt(s_norm[5]); Bt[rowOf[5] * np + cY] = -x / sqrt(s_norm[5]); }
        if (rowOf[6] >= 0) {
            Bt[rowOf[6] * np + cX] = x / sqrt(s_norm[6]); Bt[rowOf[6] * np + cY] = y / sqrt(s_norm[6]);
            Bt[rowOf[6] * np + cZ] = z / sqrt(s_norm[6]);
        }
    }
}

void launch_datum_rows(const double *xyz, const int32_t *pt_col, const int32_t *datum_pts, int nDatum, int free_mask, int d,
                       int64_t np, double *Bt, cudaStream_t s) {
    g_launch_count++;
    k_datum_rows<<<1, 256, 0, s>>>(xyz, pt_col, datum_pts, nDatum, free_mask, d, np, Bt);
}

// K4: M <- V M V + (B V)'(B V) on the lower triangle (NES:82-91 plus the SPD reformulation), identity on the padding
__global__ void __launch_bounds__(256) k_scale_system(double *__restrict__ M, int64_t ld, int u, const double *__restrict__ V,
                                                      const double *__restrict__ Bt, int d, int64_t np, int64_t row0) {
    const int64_t r = row0 + blockIdx.x;
    double *row = M + r * ld;
    if (r >= u) {
        for (int64_t c = threadIdx.x; c <= r; c += blockDim.x) row[c] = (c == r) ? 1.0 : 0.0;
        return;
    }
    const double vr = V[r];
    double br[kMaxDatum];
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) br[a] = a < d ? Bt[a * np + r] * vr : 0.0;
    for (int64_t c = threadIdx.x; c <= r; c += blockDim.x) {
        const double vc = V[c];
        double x = (vr * row[c]) * vc;
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < d) x += br[a] * (Bt[a * np + c] * vc);
        row[c] = x;
    }
}

// rows [row0, np) only (row0 > 0: the structured route scales its point blocks on the fly and never reads the rest of them)
void launch_scale_system(double *M, int64_t ld, int u, const double *V, const double *Bt, int d, int64_t np, int64_t row0, cudaStream_t s) {
    if (row0 >= np) return;
    g_launch_count++;
    k_scale_system<<<(unsigned)(np - row0), 256, 0, s>>>(M, ld, u, V, Bt, d, np, row0);
}

// the same on owner-only storage: row r of the own tiles (own_cols[lt] = first global column of local tile lt), columns <= r
__global__ void __launch_bounds__(256) k_scale_system_own(double *__restrict__ Mo, int64_t ldo, const int32_t *__restrict__ own_cols, int n_own,
                                                          int u, const double *__restrict__ V, const double *__restrict__ Bt, int d, int64_t np) {
    const int64_t r = blockIdx.x;
    double *row = Mo + r * ldo;
    const double vr = r < u ? V[r] : 0.0;
    double br[kMaxDatum];
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) br[a] = (a < d && r < u) ? Bt[a * np + r] * vr : 0.0;
    for (int64_t lc = threadIdx.x; lc < (int64_t)n_own * 128; lc += blockDim.x) {
        const int64_t c = own_cols[lc >> 7] + (lc & 127);
        if (c > r) continue;
        if (r >= u) { row[lc] = (c == r) ? 1.0 : 0.0; continue; }
        const double vc = V[c];
        double x = (vr * row[lc]) * vc;
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < d) x += br[a] * (Bt[a * np + c] * vc);
        row[lc] = x;
    }
}

void launch_scale_system_own(double *Mo, int64_t ldo, const int32_t *own_cols, int n_own, int u, const double *V, const double *Bt, int d,
                             int64_t np, cudaStream_t s) {
    if (n_own == 0) return;
    g_launch_count++;
    k_scale_system_own<<<(unsigned)np, 256, 0, s>>>(Mo, ldo, own_cols, n_own, u, V, Bt, d, np);
}

// right-hand-side block (rows): Rt[0] = V n, Rt[1+a] = B[a] V, everything else zero (Rt is kRhsRows x np, zeroed by caller)
__global__ void k_build_rhs(double *__restrict__ Rt, double *__restrict__ Btv, int64_t np, int u, const double *__restrict__ V,
                            const double *__restrict__ rhs, const double *__restrict__ Bt, int d, int simulation) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= u) return;
    const double v = V[e];
    Rt[e] = simulation ? 0.0 : v * rhs[e];
    for (int a = 0; a < d; a++) {
        const double b = Bt[a * np + e] * v;
        Rt[(1 + a) * np + e] = b;
        Btv[a * np + e] = b;
    }
}

void launch_build_rhs(double *Rt, double *Btv, int64_t np, int u, const double *V, const double *rhs, const double *Bt, int d,
                      int simulation, cudaStream_t s) {
    if (u == 0) return;
    g_launch_count++;
    k_build_rhs<<<(unsigned)((u + 255) / 256), 256, 0, s>>>(Rt, Btv, np, u, V, rhs, Bt, d, simulation);
}

// Datum correction of the solution (DESIGN.md "datum identity"):
//   z = Xt[0] = M^-1 n~, G = Xt[1..d] = (M^-1 B~')', S = B~ G', t = B~ z, c = S^-1 t,
//   y = z - G' c, dx = V y, lambda = c, H = S^-1 G, Tq[a][e] = V[e] H[a][e], Q11 = I - S^-1
// small[]: [0,49) Sinv, [49,98) Q11, [98] singular flag
__global__ void __launch_bounds__(256) k_datum_solve(const double *__restrict__ Xt, const double *__restrict__ Btv, int d,
                                                     int64_t np, int u, const double *__restrict__ V,
                                                     double *__restrict__ dxref, double *__restrict__ H,
                                                     double *__restrict__ Tq, double *__restrict__ small) {
    __shared__ double red[8][kMaxDatum * (kMaxDatum + 1)];
    __shared__ double s_S[kMaxDatum][kMaxDatum], s_Sinv[kMaxDatum][kMaxDatum], s_t[kMaxDatum], s_c[kMaxDatum];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (d > 0) {
        double acc[kMaxDatum][kMaxDatum + 1];
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
#pragma unroll
            for (int b = 0; b <= kMaxDatum; b++) acc[a][b] = 0.0;
        for (int64_t e = tid; e < u; e += 256) {
            double bt[kMaxDatum], x[kMaxDatum + 1];
#pragma unroll
            for (int a = 0; a < kMaxDatum; a++) bt[a] = a < d ? Btv[a * np + e] : 0.0;
#pragma unroll
            for (int b = 0; b <= kMaxDatum; b++) x[b] = b <= d ? Xt[b * np + e] : 0.0;
#pragma unroll
            for (int a = 0; a < kMaxDatum; a++)
#pragma unroll
                for (int b = 0; b <= kMaxDatum; b++) acc[a][b] += bt[a] * x[b];
        }
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
#pragma unroll
            for (int b = 0; b <= kMaxDatum; b++) {
                double v = acc[a][b];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp][a * (kMaxDatum + 1) + b] = v;
            }
        __syncthreads();
        if (tid < kMaxDatum * (kMaxDatum + 1)) {
            double s = 0.0;
            for (int w = 0; w < 8; w++) s += red[w][tid];
            const int a = tid / (kMaxDatum + 1), b = tid % (kMaxDatum + 1);
            if (b == 0) s_t[a] = s; else s_S[a][b - 1] = s;
        }
        __syncthreads();
        if (tid == 0) {
            // Gauss-Jordan with partial pivoting on the d x d matrix S
            double A[kMaxDatum][2 * kMaxDatum];
            for (int i = 0; i < d; i++)
                for (int j = 0; j < d; j++) { A[i][j] = 0.5 * (s_S[i][j] + s_S[j][i]); A[i][d + j] = (i == j) ? 1.0 : 0.0; }
            bool sing = false;
            for (int c = 0; c < d; c++) {
                int pr = c;
                for (int i = c + 1; i < d; i++) if (fabs(A[i][c]) > fabs(A[pr][c])) pr = i;
                if (!(fabs(A[pr][c]) > 0.0)) { sing = true; break; }
                if (pr != c) for (int j = 0; j < 2 * d; j++) { const double t = A[c][j]; A[c][j] = A[pr][j]; A[pr][j] = t; }
                const double pv = 1.0 / A[c][c];
                for (int j = 0; j < 2 * d; j++) A[c][j] *= pv;
                for (int i = 0; i < d; i++) {
                    if (i == c) continue;
                    const double f = A[i][c];
                    for (int j = 0; j < 2 * d; j++) A[i][j] -= f * A[c][j];
                }
            }
            for (int i = 0; i < d; i++) {
                double cs = 0.0;
                for (int j = 0; j < d; j++) { s_Sinv[i][j] = A[i][d + j]; cs += A[i][d + j] * s_t[j]; }
                s_c[i] = cs;
            }
            for (int i = 0; i < d; i++)
                for (int j = 0; j < d; j++) {
                    small[i * kMaxDatum + j] = s_Sinv[i][j];
                    small[49 + i * kMaxDatum + j] = ((i == j) ? 1.0 : 0.0) - s_Sinv[i][j];
                }
            small[98] = sing ? 1.0 : 0.0;
            for (int i = 0; i < d; i++) dxref[i] = s_c[i];
        }
        __syncthreads();
    }
    for (int64_t e = tid; e < u; e += 256) {
        double y = Xt[e];
        const double v = V[e];
        if (d > 0) {
            double g[kMaxDatum];
#pragma unroll
            for (int b = 0; b < kMaxDatum; b++) g[b] = b < d ? Xt[(1 + b) * np + e] : 0.0;
#pragma unroll
            for (int b = 0; b < kMaxDatum; b++) if (b < d) y -= g[b] * s_c[b];
            for (int a = 0; a < d; a++) {
                double h = 0.0;
#pragma unroll
                for (int b = 0; b < kMaxDatum; b++) if (b < d) h += s_Sinv[a][b] * g[b];
                H[a * np + e] = h;
                Tq[a * np + e] = v * h;
            }
        }
        dxref[d + e] = v * y;
    }
}

void launch_datum_solve(const double *Xt, const double *Btv, int d, int64_t np, int u, const double *V, double *dxref, double *H,
                        double *Tq, double *small, cudaStream_t s) {
    g_launch_count++;
    k_datum_solve<<<1, 256, 0, s>>>(Xt, Btv, d, np, u, V, dxref, H, Tq, small);
}

// K7: Qxx(lower) = V (M^-1 - H' G) V, BA:273 + rank-d datum correction
__global__ void __launch_bounds__(256) k_qxx_epilogue(double *__restrict__ M, int64_t ld, int u, const double *__restrict__ V,
                                                      const double *__restrict__ H, const double *__restrict__ G, int d,
                                                      int64_t np) {
    const int64_t r = blockIdx.x;
    if (r >= u) return;
    double *row = M + r * ld;
    const double vr = V[r];
    double hr[kMaxDatum];
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) hr[a] = a < d ? H[a * np + r] : 0.0;
    for (int64_t c = threadIdx.x; c <= r; c += blockDim.x) {
        double x = row[c];
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < d) x -= hr[a] * G[a * np + c];
        row[c] = (V[c] * x) * vr;
    }
}

void launch_qxx_epilogue(double *M, int64_t ld, int u, const double *V, const double *H, const double *G, int d, int64_t np,
                         cudaStream_t s) {
    if (u == 0) return;
    g_launch_count++;
    k_qxx_epilogue<<<(unsigned)u, 256, 0, s>>>(M, ld, u, V, H, G, d, np);
}

// K9: x += dx, max|dx| over unknown columns (BA:450-462).  out[0] = max|dx| (as double bits, non-negative => monotone),
// out[1] = NaN/Inf flag
__global__ void __launch_bounds__(256) k_update(double *__restrict__ val, const int32_t *__restrict__ col, int64_t n,
                                                const double *__restrict__ dxref, int apply,
                                                unsigned long long *__restrict__ out) {
    double mx = 0.0;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = col[i];
        if (!col_active_s(c)) continue;
        const double dv = dxref[c];
        if (!(fabs(dv) <= 1.7976931348623157e308)) bad = true;
        mx = fmax(mx, fabs(dv));
        if (apply) val[i] += dv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, (unsigned long long)__double_as_longlong(mx));
        if (anybad) atomicMax(out + 1, 1ull);
    }
}

void launch_update(double *val, const int32_t *col, int64_t n, const double *dxref, int apply, unsigned long long *out,
                   cudaStream_t s) {
    if (n == 0) return;
    const int64_t blocks = (n + 255) / 256;
    g_launch_count++;
    k_update<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, s>>>(val, col, n, dxref, apply, out);
}

// Pack reference columns [c0, c1) of the (u+d) symmetric matrix into MTJ packed-upper order.
// lower = u x u lower triangle (row-major, ld); border[a][e] (d x np), q11 (7x7 stride kMaxDatum)
__global__ void __launch_bounds__(256) k_pack_columns(const double *__restrict__ lower, int64_t ld, const double *__restrict__ border,
                                                      int64_t np, const double *__restrict__ q11, int d, int64_t c0,
                                                      int64_t c1, double *__restrict__ out) {
    const int64_t c = c0 + blockIdx.x;
    if (c >= c1) return;
    const int64_t off = c * (c + 1) / 2 - c0 * (c0 + 1) / 2;
    double *o = out + off;
    if (c < d) {
        for (int r = threadIdx.x; r <= c; r += blockDim.x) o[r] = q11 ? q11[r * kMaxDatum + c] : 0.0;
        return;
    }
    const int64_t e = c - d;
    for (int a = threadIdx.x; a < d; a += blockDim.x) o[a] = border[a * np + e];
    const double *row = lower + e * ld;
    for (int64_t k = threadIdx.x; k <= e; k += blockDim.x) o[d + k] = row[k];
}

void launch_pack_columns(const double *lower, int64_t ld, const double *border, int64_t np, const double *q11, int d, int64_t c0,
                         int64_t c1, double *out, cudaStream_t s) {
    if (c1 <= c0) return;
    g_launch_count++;
    k_pack_columns<<<(unsigned)(c1 - c0), 256, 0, s>>>(lower, ld, border, np, q11, d, c0, c1, out);
}

// rectangular block of the full symmetric (u+d) matrix
__global__ void __launch_bounds__(256) k_get_block(const double *__restrict__ lower, int64_t ld, const double *__restrict__ border,
                                                   int64_t np, const double *__restrict__ q11, int d, int r0, int r1, int c0,
                                                   int c1, double *__restrict__ out) {
    const int64_t total = (int64_t)(r1 - r0) * (c1 - c0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = r0 + (int)(i / (c1 - c0)), c = c0 + (int)(i % (c1 - c0));
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        double v;
        if (hi < d) v = q11 ? q11[lo * kMaxDatum + hi] : 0.0;
        else if (lo < d) v = border[(int64_t)lo * np + (hi - d)];
        else v = lower[(int64_t)(hi - d) * ld + (lo - d)];
        out[i] = v;
    }
}

void launch_get_block(const double *lower, int64_t ld, const double *border, int64_t np, const double *q11, int d, int r0, int r1,
                      int c0, int c1, double *out, cudaStream_t s) {
    const int64_t total = (int64_t)(r1 - r0) * (c1 - c0);
    if (total <= 0) return;
    const int64_t blocks = (total + 255) / 256;
    g_launch_count++;
    k_get_block<<<(unsigned)(blocks < 4736 ? blocks : 4736), 256, 0, s>>>(lower, ld, border, np, q11, d, r0, r1, c0, c1, out);
}

// Levenberg-Marquardt: N_cc += lambda * N_cc for every unknown column (BA:814-822)
__global__ void k_damp_diag(SysView M, int u, double lambda) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= u) return;
    double *q = sys_at(M, e, e);
    if (q) { const double v = *q; *q = v + lambda * v; }
}
void launch_damp_diag(const SysView &M, int u, double lambda, cudaStream_t s) {
    if (u) { g_launch_count++; k_damp_diag<<<(u + 255) / 256, 256, 0, s>>>(M, u, lambda); }
}
__global__ void k_scale_vector(double *__restrict__ x, int64_t n, double alpha) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= alpha;
}
void launch_scale_vector(double *x, int64_t n, double alpha, cudaStream_t s) {
    if (n) { g_launch_count++; k_scale_vector<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n, alpha); }
}

// ---- column-tile form of the inverse (multi-GPU: every rank holds np x 128*ntc) -----------------------------------
__global__ void k_identity_columns(double *__restrict__ X, int64_t ldx, const int32_t *__restrict__ ktab, int ntc) {
    const int jl = blockIdx.x, i = threadIdx.x;
    if (jl < ntc) X[(int64_t)(ktab[jl] + i) * ldx + jl * 128 + i] = 1.0;
}

void launch_identity_columns(double *X, int64_t ldx, int64_t np, const int32_t *ktab, int ntc, cudaStream_t s) {
    if (ntc) { g_launch_count++; k_identity_columns<<<ntc, 128, 0, s>>>(X, ldx, ktab, ntc); }
}

// K7 on column tiles: Qxx[r][c] = V[r] V[c] (Minv[r][c] - sum_a H[a][r] G[a][c]) for r >= first row of the tile
__global__ void __launch_bounds__(256) k_qxx_epilogue_cols(double *__restrict__ X, int64_t ldx, int ntc, const int32_t *__restrict__ ktab,
                                                           int u, const double *__restrict__ V, const double *__restrict__ H,
                                                           const double *__restrict__ G, int d, int64_t np) {
    const int64_t r = blockIdx.x;
    if (r >= u) return;
    const double vr = V[r];
    double hr[kMaxDatum];
#pragma unroll
    for (int a = 0; a < kMaxDatum; a++) hr[a] = a < d ? H[a * np + r] : 0.0;
    double *row = X + r * ldx;
    for (int64_t cl = threadIdx.x; cl < (int64_t)ntc * 128; cl += blockDim.x) {
        const int64_t c = ktab[cl >> 7] + (cl & 127);
        if (c > r || c >= u) continue;
        double x = row[cl];
#pragma unroll
        for (int a = 0; a < kMaxDatum; a++)
            if (a < d) x -= hr[a] * G[a * np + c];
        row[cl] = (V[c] * x) * vr;
    }
}

void launch_qxx_epilogue_cols(double *X, int64_t ldx, int ntc, const int32_t *ktab, int u, const double *V, const double *H,
                              const double *G, int d, int64_t np, cudaStream_t s) {
    if (u == 0 || ntc == 0) return;
    g_launch_count++;
    k_qxx_epilogue_cols<<<(unsigned)u, 256, 0, s>>>(X, ldx, ntc, ktab, u, V, H, G, d, np);
}

// block of the full symmetric matrix from the column-tile form; entries owned by other ranks are 0 (the sum over
// ranks is the block); the replicated border is contributed by rank 0 only
__global__ void __launch_bounds__(256) k_get_block_dist(const double *__restrict__ X, int64_t ldx, const int32_t *__restrict__ col_local,
                                                        const double *__restrict__ border, int64_t np, const double *__restrict__ q11,
                                                        int d, int rank, int r0, int r1, int c0, int c1, double *__restrict__ out) {
    const int64_t total = (int64_t)(r1 - r0) * (c1 - c0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = r0 + (int)(i / (c1 - c0)), c = c0 + (int)(i % (c1 - c0));
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        double v = 0.0;
        if (hi < d) v = rank == 0 ? q11[lo * kMaxDatum + hi] : 0.0;
        else if (lo < d) v = rank == 0 ? border[(int64_t)lo * np + (hi - d)] : 0.0;
        else {
            const int e = lo - d;
            const int jl = col_local[e >> 7];
            if (jl >= 0) v = X[(int64_t)(hi - d) * ldx + jl * 128 + (e & 127)];
        }
        out[i] = v;
    }
}

void launch_get_block_dist(const double *X, int64_t ldx, const int32_t *col_local, const double *border, int64_t np,
                           const double *q11, int d, int rank, int r0, int r1, int c0, int c1, double *out, cudaStream_t s) {
    const int64_t total = (int64_t)(r1 - r0) * (c1 - c0);
    if (total <= 0) return;
    const int64_t blocks = (total + 255) / 256;
    g_launch_count++;
    k_get_block_dist<<<(unsigned)(blocks < 4736 ? blocks : 4736), 256, 0, s>>>(X, ldx, col_local, border, np, q11, d, rank, r0, r1, c0, c1, out);
}

// Gather of an arbitrary sub-matrix of Qxx (what util/io/writer/MatlabResultWriter.java:209-223 and
// DefaultResultWriter.java:126-155 loop over with cofactor.get(row, column)): out[i][j] = scale * Qxx[row_idx[i]][col_idx[j]].
// Single GPU: lower != nullptr; distributed: X/col_local (entries owned elsewhere are 0, the sum over ranks is the matrix).
__global__ void __launch_bounds__(256) k_get_submatrix(const double *__restrict__ lower, int64_t ld, const double *__restrict__ X, int64_t ldx,
                                                       const int32_t *__restrict__ col_local, const double *__restrict__ border, int64_t np,
                                                       const double *__restrict__ q11, int d, int rank, const int32_t *__restrict__ row_idx,
                                                       int n_rows, const int32_t *__restrict__ col_idx, int n_cols, double scale,
                                                       double *__restrict__ out) {
    const int64_t total = (int64_t)n_rows * n_cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = row_idx[i / n_cols], c = col_idx[i % n_cols];
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        double v = 0.0;
        if (hi < d) v = rank == 0 ? q11[lo * kMaxDatum + hi] : 0.0;
        else if (lo < d) v = rank == 0 ? border[(int64_t)lo * np + (hi - d)] : 0.0;
        else if (lower) v = lower[(int64_t)(hi - d) * ld + (lo - d)];
        else {
            const int e = lo - d;
            const int jl = col_local[e >> 7];
            if (jl >= 0) v = X[(int64_t)(hi - d) * ldx + jl * 128 + (e & 127)];
        }
        out[i] = scale * v;
    }
}

void launch_get_submatrix(const double *lower, int64_t ld, const double *X, int64_t ldx, const int32_t *col_local, const double *border,
                          int64_t np, const double *q11, int d, int rank, const int32_t *row_idx, int n_rows, const int32_t *col_idx,
                          int n_cols, double scale, double *out, cudaStream_t s) {
    const int64_t total = (int64_t)n_rows * n_cols;
    if (total <= 0) return;
    const int64_t blocks = (total + 255) / 256;
    g_launch_count++;
    k_get_submatrix<<<(unsigned)(blocks < 4736 ? blocks : 4736), 256, 0, s>>>(lower, ld, X, ldx, col_local, border, np, q11, d, rank, row_idx,
                                                                            n_rows, col_idx, n_cols, scale, out);
}

// ---- directly observed parameter groups (PDF:447-473) -------------------------------------------------------------
// w_i = obs_i - value(target_i)
__global__ void k_group_w(int r, const double *const *__restrict__ tptr, const double *__restrict__ obs, double *__restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r) w[i] = obs[i] - *tptr[i];
}

// diagonal weights: N[c,c] += sigma0^2/var, n[c] += P w
__global__ void k_group_stack_diag(int r, const int32_t *__restrict__ col, const double *__restrict__ var, double sigma2,
                                   const double *__restrict__ w, int d, SysView M, double *__restrict__ rhs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r || !col_active_s(col[i])) return;
    const double P = sigma2 / var[i];
    const int64_t e = col[i] - d;
    sys_add(M, e, e, P);
    rhs[e] += P * w[i];
}

// full weights (Pw = symmetric r x r, row-major, leading dimension ldp): n[c_i] += sum_j P_ij w_j; N[c_i,c_j] += P_ij
__global__ void __launch_bounds__(256) k_group_stack_full(int r, const int32_t *__restrict__ col, const double *__restrict__ Pw,
                                                          int64_t ldp, const double *__restrict__ w, int d, SysView M,
                                                          double *__restrict__ rhs) {
    __shared__ double red[8];
    const int i = blockIdx.x;
    if (!col_active_s(col[i])) return;
    const int64_t ei = col[i] - d;
    const double *Pi = Pw + (int64_t)i * ldp;
    double s = 0.0;
    for (int j = threadIdx.x; j < r; j += blockDim.x) {
        const double p = Pi[j];
        s += p * w[j];
        if (j <= i && col_active_s(col[j])) sys_add(M, ei, col[j] - d, p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += red[k];
        rhs[ei] += t;
    }
}

// omega contribution: v = w - dx[col]; out += v' P v
__global__ void __launch_bounds__(256) k_group_omega(int r, const int32_t *__restrict__ col, const double *__restrict__ var,
                                                     const double *__restrict__ Pw, int64_t ldp, double sigma2,
                                                     const double *__restrict__ w, const double *__restrict__ dxref,
                                                     double *__restrict__ out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < r; i += blockDim.x) {
        const double vi = w[i] - (col_active_s(col[i]) ? dxref[col[i]] : 0.0);
        if (Pw == nullptr) {
            s += vi * vi * (sigma2 / var[i]);
        } else {
            double pv = 0.0;
            for (int j = 0; j < r; j++) {
                const double vj = w[j] - (col_active_s(col[j]) ? dxref[col[j]] : 0.0);
                pv += Pw[(int64_t)i * ldp + j] * vj;
            }
            s += vi * pv;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += red[k];
        out[0] = t;
    }
}

void launch_group_w(int r, const double *const *tptr, const double *obs, double *w, cudaStream_t s) {
    if (r) { g_launch_count++; k_group_w<<<(r + 255) / 256, 256, 0, s>>>(r, tptr, obs, w); }
}
void launch_group_stack(int r, const int32_t *col, const double *var, const double *Pw, int64_t ldp, double sigma2, const double *w,
                        int d, const SysView &M, double *rhs, cudaStream_t s) {
    if (!r) return;
    if (Pw == nullptr) { g_launch_count++; k_group_stack_diag<<<(r + 255) / 256, 256, 0, s>>>(r, col, var, sigma2, w, d, M, rhs); }
    else { g_launch_count++; k_group_stack_full<<<r, 256, 0, s>>>(r, col, Pw, ldp, w, d, M, rhs); }
}
void launch_group_omega(int r, const int32_t *col, const double *var, const double *Pw, int64_t ldp, double sigma2, const double *w,
                        const double *dxref, double *out, cudaStream_t s) {
    g_launch_count++;
    k_group_omega<<<1, 256, 0, s>>>(r, col, var, Pw, ldp, sigma2, w, dxref, out);
}

// dense symmetric helpers for the group weight matrix: expand packed-upper Sigma/sigma0^2 into the lower triangle of
// a padded square (identity padding), and mirror a lower triangle into a full symmetric matrix
__global__ void k_unpack_scaled(const double *__restrict__ ap, int r, double scale, double *__restrict__ M, int64_t ld, int64_t np) {
    const int64_t row = blockIdx.x;
    for (int64_t c = threadIdx.x; c <= row; c += blockDim.x) {
        double v;
        if (row < r) v = ap[c + row * (row + 1) / 2] * scale;
        else v = (c == row) ? 1.0 : 0.0;
        M[row * ld + c] = v;
    }
}
__global__ void k_symmetrize(double *__restrict__ M, int64_t ld, int r) {
    const int64_t row = blockIdx.x;
    for (int64_t c = threadIdx.x; c < row; c += blockDim.x) M[c * ld + row] = M[row * ld + c];
}
void launch_unpack_scaled(const double *ap, int r, double scale, double *M, int64_t ld, int64_t np, cudaStream_t s) {
    g_launch_count++;
    k_unpack_scaled<<<(unsigned)np, 256, 0, s>>>(ap, r, scale, M, ld, np);
}
void launch_symmetrize(double *M, int64_t ld, int r, cudaStream_t s) {
    if (r) { g_launch_count++; k_symmetrize<<<r, 256, 0, s>>>(M, ld, r); }
}

// matrix-free product with a directly observed group's block of N (verification entry point jaicov_normal_product):
// Y_v[c_i] += sum_j P_ij x_v[c_j], rhs[c_i] += sum_j P_ij w_j, wpw += w'Pw.  One CTA per row i; plain += is safe because the
// targets of one group are distinct unknowns and the groups run one after the other on the stream.
__global__ void __launch_bounds__(256) k_group_product(int r, const int32_t *__restrict__ col, const double *__restrict__ var,
                                                       const double *__restrict__ Pw, int64_t ldp, double sigma2,
                                                       const double *__restrict__ w, int nv, int64_t n, const double *__restrict__ X,
                                                       double *__restrict__ Y, double *__restrict__ rhs, double *__restrict__ wpw) {
    __shared__ double red[8];
    const int i = blockIdx.x;
    const bool act = col_active_s(col[i]);
    auto block_sum = [&](double s) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += red[k];
        return t;
    };
    for (int v = -1; v < nv; v++) {          // v = -1: the right-hand side and w'Pw
        if (v < 0 && !rhs) continue;
        const double *x = v < 0 ? nullptr : X + (int64_t)v * n;
        double s = 0.0;
        if (Pw == nullptr) {
            if (threadIdx.x == 0) s = (sigma2 / var[i]) * (v < 0 ? w[i] : (act ? x[col[i]] : 0.0));
        } else {
            const double *Pi = Pw + (int64_t)i * ldp;
            for (int j = threadIdx.x; j < r; j += blockDim.x)
                s += Pi[j] * (v < 0 ? w[j] : (col_active_s(col[j]) ? x[col[j]] : 0.0));
        }
        const double t = block_sum(s);
        if (threadIdx.x == 0) {
            if (v < 0) {
                if (act) rhs[col[i]] += t;
                if (wpw) atomicAdd(wpw, w[i] * t);
            } else if (act) {
                Y[(int64_t)v * n + col[i]] += t;
            }
        }
    }
}

void launch_group_product(int r, const int32_t *col, const double *var, const double *Pw, int64_t ldp, double sigma2, const double *w,
                          int nv, int64_t n, const double *X, double *Y, double *rhs, double *wpw, cudaStream_t s) {
    if (!r) return;
    g_launch_count++;
    k_group_product<<<r, 256, 0, s>>>(r, col, var, Pw, ldp, sigma2, w, nv, n, X, Y, rhs, wpw);
}

}  // namespace jaicov
