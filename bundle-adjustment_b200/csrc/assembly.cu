// assembly.cu -- K1/K2/K8 of SURVEY.md section 2.1: residual + Jacobian evaluation, normal-equation assembly
// N = A'PA, n = A'Pw and Omega = v'Pv, for sm_100a.
//
// Reference semantics: PartialDerivativeFactory.getPartialDerivativeImageCoordinate (PDF:285-445) followed by
// stackNormalEquationSystem (PDF:475-505) for every image point, getPartialDerivativeScaleBar (PDF:210-283),
// BundleAdjustment.getOmega (BA:472-491).
//
// B200 design (nothing here mirrors the Java control flow):
//  * one thread per image observation evaluates the model (model.cuh) with the camera staged in shared memory;
//  * a warp's 32 observations are staged as 64 weighted rows R*[A_cam | w] (P = R'R) in a shared-memory tile and
//    contracted with FP64 tensor-core tiles (mma.sync m8n8k4.f64 -> SASS DMMA): the per-image camera block
//    [EO | IO | coefficients | w]' P [EO | IO | coefficients | w] is a 64 x NC Gram per warp batch;
//  * ownership instead of atomics: the EO x point 6x3 block of an observation has exactly one producer and is
//    stored straight into N; image-level sums go through per-work-item partials that are reduced in a fixed order
//    (bitwise reproducible); point-level sums (point x point, point x IO, n_point) are produced by a second sweep
//    with one warp per object point over that point's observations (CSC), again as DMMA tiles;
//  * the packed per-point / per-camera partial buffers are exactly what an image-sharded multi-GPU run all-reduces.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "model.cuh"
#include "sweep_common.cuh"

namespace jaicov {

// ---- pose table ---------------------------------------------------------------------------------------------------
__global__ void k_pose(DevProblem P) {
    int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= P.nImg) return;
    ImgPose q = make_pose(P.eo_val + 6 * (int64_t)img);
    double *p = P.pose + (int64_t)img * kPoseStride;
    p[0] = q.r11; p[1] = q.r12; p[2] = q.r13; p[3] = q.r21; p[4] = q.r22; p[5] = q.r23;
    p[6] = q.r31; p[7] = q.r32; p[8] = q.r33; p[9] = q.sinK; p[10] = q.cosK; p[11] = q.X0; p[12] = q.Y0; p[13] = q.Z0;
    p[14] = 0.0; p[15] = 0.0;
}

void launch_pose(const DevProblem &P, cudaStream_t s) {
    if (P.nImg == 0) return;
    g_launch_count++;
    k_pose<<<(P.nImg + 127) / 128, 128, 0, s>>>(P);
}

// ---- K1 materialised (parity / profiling) ---------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_eval_k1(DevProblem P, int ns_max, double *__restrict__ a, double *__restrict__ w,
                                                  double *__restrict__ p) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.m) return;
    const int img = P.img_of_obs[j];
    const int cam = P.cam_of_img[img];
    const int pt = P.obj_idx[j];
    const ImgPose q = load_pose(P.pose, img);
    const CamView cv = view_global(P, cam);
    double *a0 = a + j * 2 * (int64_t)ns_max, *a1 = a0 + ns_max;
    const int32_t *ccol = P.coef_col + P.coef_ptr[cam];
    BaseRows r;
    eval_observation(q, cv, P.xyz[3 * (int64_t)pt], P.xyz[3 * (int64_t)pt + 1], P.xyz[3 * (int64_t)pt + 2], P.xy[2 * j],
                     P.xy[2 * j + 1], r, [&](int k, double v0, double v1) {
                         const bool act = col_active(ccol[k]);
                         a0[12 + k] = act ? v0 : 0.0;
                         a1[12 + k] = act ? v1 : 0.0;
                     });
    for (int s = 12 + cv.ncoef; s < ns_max; s++) { a0[s] = 0.0; a1[s] = 0.0; }
    const int32_t *pc = P.pt_col + 3 * (int64_t)pt, *ic = P.io_col + 3 * cam, *ec = P.eo_col + 6 * (int64_t)img;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const bool ap = col_active(pc[i]), ae = col_active(ec[i]), ar = col_active(ec[3 + i]);
        a0[i] = ap ? r.ax[i] : 0.0;  a1[i] = ap ? r.ay[i] : 0.0;
        a0[6 + i] = ae ? -r.ax[i] : 0.0;  a1[6 + i] = ae ? -r.ay[i] : 0.0;
        a0[9 + i] = ar ? r.ax[4 + i] : 0.0;  a1[9 + i] = ar ? r.ay[4 + i] : 0.0;
    }
    a0[3] = col_active(ic[0]) ? 1.0 : 0.0;  a1[3] = 0.0;
    a0[4] = 0.0;  a1[4] = col_active(ic[1]) ? 1.0 : 0.0;
    a0[5] = col_active(ic[2]) ? r.ax[3] : 0.0;  a1[5] = col_active(ic[2]) ? r.ay[3] : 0.0;
    w[2 * j] = r.w0;  w[2 * j + 1] = r.w1;
    double p00, p01, p11;
    point_weight(P.sigma2, P.var[2 * j], P.var[2 * j + 1], P.rho[j], p00, p01, p11);
    p[3 * j] = p00;  p[3 * j + 1] = p01;  p[3 * j + 2] = p11;
}

void launch_eval_k1(const DevProblem &P, int ns_max, double *a, double *w, double *p, cudaStream_t s) {
    if (P.m == 0) return;
    g_launch_count++;
    k_eval_k1<<<(unsigned)((P.m + 127) / 128), 128, 0, s>>>(P, ns_max, a, w, p);
}

// ---- by-image sweep ---------------------------------------------------------------------------------------------
// tile columns: 0..5 EO (X0,Y0,Z0,omega,phi,kappa), 6..8 IO (x0,y0,c), 9..9+ncoef-1 coefficients, 9+ncoef = w.
constexpr int kImgWarps = 4;

template <int NT, int MINB, bool STD>
__global__ void __launch_bounds__(kImgWarps * 32, MINB) k_by_image(DevProblem P, const WorkItem *__restrict__ work,
                                                                   double *__restrict__ partial, SysView M) {
    constexpr int LDT = 68;  // tile is stored column-major [NC][68]: == 4 (mod 16) -> conflict-free fragment reads,
                             // and lanes write consecutive rows of one column -> conflict-free stores
    constexpr int NC = 8 * NT;
    extern __shared__ double smem[];
    __shared__ CamSmem cs;
    __shared__ int32_t s_eocol[6];
    __shared__ ImgPose s_pose;                 // per-image constants stay in shared memory: 28 registers less per thread
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WorkItem wi = work[blockIdx.x];
    const int img = wi.img, cam = P.cam_of_img[img];
    load_camera(P, cam, cs, tid, blockDim.x);
    if (tid < 6) s_eocol[tid] = P.eo_col[6 * (int64_t)img + tid];
    if (tid < 14) reinterpret_cast<double *>(&s_pose)[tid] = P.pose[(int64_t)img * kPoseStride + tid];
    double *tile = smem + (size_t)warp * NC * LDT;
    for (int i = lane; i < NC * LDT; i += 32) tile[i] = 0.0;
    __syncthreads();
    const CamView cv = view_of(P, cs);
    const ImgPose &q = s_pose;
    const int wcol = 9 + cs.ncoef;
    const int d = P.d;

    double acc[NT][NT][2];
#pragma unroll
    for (int i = 0; i < NT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // lane l owns tile rows l and l+32; element (row, col) lives at tile[col * LDT + row]
#define ROW0(c) tile[(c) * LDT + lane]
#define ROW1(c) tile[(c) * LDT + lane + 32]
    for (int64_t base = wi.begin + warp * 32; base < wi.end; base += kImgWarps * 32) {
        const int64_t j = base + lane;
        if (j < wi.end) {
            const int pt = P.obj_idx[j];
            // P = R'R, R = [r00 r01; 0 r11]: constant over the passes, precomputed per observation (k_obs_weights)
            const double r00 = P.rw[3 * j], r01 = P.rw[3 * j + 1], r11 = P.rw[3 * j + 2];
            const double2 xyo = reinterpret_cast<const double2 *>(P.xy)[j];
            BaseRows r;
            eval_observation_t<true, STD>(q, cv, P.xyz[3 * (int64_t)pt], P.xyz[3 * (int64_t)pt + 1], P.xyz[3 * (int64_t)pt + 2],
                                          xyo.x, xyo.y, r, [&](int k, double v0, double v1) {
                                              ROW0(9 + k) = r00 * v0 + r01 * v1;
                                              ROW1(9 + k) = r11 * v1;
                                          });
            // weighted base rows
            double tpx[3], tpy[3], tex[6], tey[6];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                tpx[i] = r00 * r.ax[i] + r01 * r.ay[i];
                tpy[i] = r11 * r.ay[i];
                tex[i] = -tpx[i];  tey[i] = -tpy[i];                         // X0,Y0,Z0 = -(X,Y,Z)
                tex[3 + i] = r00 * r.ax[4 + i] + r01 * r.ay[4 + i];          // omega,phi,kappa
                tey[3 + i] = r11 * r.ay[4 + i];
            }
#pragma unroll
            for (int e = 0; e < 6; e++) { ROW0(e) = tex[e]; ROW1(e) = tey[e]; }
            ROW0(6) = r00;  ROW1(6) = 0.0;                                   // x0: (1,0)
            ROW0(7) = r01;  ROW1(7) = r11;                                   // y0: (0,1)
            ROW0(8) = r00 * r.ax[3] + r01 * r.ay[3];  ROW1(8) = r11 * r.ay[3];  // c
            ROW0(wcol) = r00 * r.w0 + r01 * r.w1;  ROW1(wcol) = r11 * r.w1;
            // EO x point block: single producer -> plain stores
            const int32_t *pc = P.pt_col + 3 * (int64_t)pt;
#pragma unroll
            for (int e = 0; e < 6; e++) {
                const int32_t ce = s_eocol[e];
                if (!col_active(ce)) continue;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int32_t cp = pc[c];
                    if (!col_active(cp)) continue;
                    sys_set(M, ce - d, cp - d, tex[e] * tpx[c] + tey[e] * tpy[c]);
                }
            }
        } else {
            for (int i = 0; i <= wcol; i++) { ROW0(i) = 0.0; ROW1(i) = 0.0; }
        }
        __syncwarp();
        // Gram of the 64 x NC tile on FP64 tensor cores
#pragma unroll 4
        for (int ks = 0; ks < 16; ks++) {
            double f[NT];
            const double *src = tile + (lane >> 2) * LDT + 4 * ks + (lane & 3);
#pragma unroll
            for (int b = 0; b < NT; b++) f[b] = src[8 * b * LDT];
#pragma unroll
            for (int i = 0; i < NT; i++)
#pragma unroll
                for (int jb = 0; jb < NT; jb++) dmma884(acc[i][jb][0], acc[i][jb][1], f[i], f[jb]);
        }
        __syncwarp();
    }
    // cross-warp reduction in warp order (deterministic)
    __syncthreads();
#undef ROW0
#undef ROW1
    double *red = smem + (size_t)warp * NC * LDT;  // each warp reuses its own tile: NC*NC <= NC*LDT
#pragma unroll
    for (int i = 0; i < NT; i++)
#pragma unroll
        for (int jb = 0; jb < NT; jb++) {
            const int rr = 8 * i + (lane >> 2), cc = 8 * jb + 2 * (lane & 3);
            red[rr * NC + cc] = acc[i][jb][0];
            red[rr * NC + cc + 1] = acc[i][jb][1];
        }
    __syncthreads();
    double *out = partial + (size_t)blockIdx.x * NC * NC;
    for (int i = tid; i < NC * NC; i += blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int wp = 0; wp < kImgWarps; wp++) s += smem[(size_t)wp * NC * LDT + i];
        out[i] = s;
    }
}

// per image: sum work-item partials (fixed order), scatter EO blocks, hand the camera block to cam_partial
__global__ void __launch_bounds__(256) k_image_finalize(DevProblem P, AssemblyScratch S, SysView M, double *__restrict__ rhs) {
    extern __shared__ double G[];  // NC*NC
    const int img = P.img0 + blockIdx.x, tid = threadIdx.x;
    const int NC = 8 * S.ntImg;
    const int w0 = S.img_work_ptr[blockIdx.x], w1 = S.img_work_ptr[blockIdx.x + 1];   // work items of local image #blockIdx.x
    for (int i = tid; i < NC * NC; i += blockDim.x) {
        double s = 0.0;
        for (int w = w0; w < w1; w++) s += S.img_partial[(size_t)w * NC * NC + i];
        G[i] = s;
    }
    __syncthreads();
    const int cam = P.cam_of_img[img];
    const int ncoef = P.coef_ptr[cam + 1] - P.coef_ptr[cam];
    const int kc = 3 + ncoef, wcol = 9 + ncoef;
    const int32_t *ec = P.eo_col + 6 * (int64_t)img;
    const int32_t *cc = P.campos_col + P.cam_kbase[cam];
    const int d = P.d;
    // EO x EO (lower incl. diagonal), EO x camera parameters, rhs of EO
    for (int i = tid; i < 6 * (6 + kc + 1); i += blockDim.x) {
        const int e = i / (6 + kc + 1), c = i % (6 + kc + 1);
        const int32_t ce = ec[e];
        if (!col_active(ce)) continue;
        if (c < 6) {
            if (c > e) continue;
            const int32_t c2 = ec[c];
            if (!col_active(c2)) continue;
            sys_add(M, ce - d, c2 - d, G[e * NC + c]);
        } else if (c < 6 + kc) {
            const int32_t c2 = cc[c - 6];
            if (!col_active(c2)) continue;
            sys_add(M, ce - d, c2 - d, G[e * NC + c]);
        } else {
            rhs[ce - d] += G[e * NC + wcol];
        }
    }
    // camera block (kc x kc) and its rhs column -> cam_partial[img][kcMax][kcMax+1]
    double *cp = S.cam_partial + (size_t)img * S.kcMax * (S.kcMax + 1);
    for (int i = tid; i < kc * (kc + 1); i += blockDim.x) {
        const int a = i / (kc + 1), b = i % (kc + 1);
        cp[a * (S.kcMax + 1) + b] = (b < kc) ? G[(6 + a) * NC + 6 + b] : G[(6 + a) * NC + wcol];
    }
}

// per camera: sum the camera blocks of this rank's images in image order (-> cam_sum, all-reduced across ranks)
__global__ void __launch_bounds__(256) k_camera_sum(DevProblem P, AssemblyScratch S) {
    const int cam = blockIdx.x, tid = threadIdx.x;
    const int kc = 3 + P.coef_ptr[cam + 1] - P.coef_ptr[cam];
    double *out = S.cam_sum + (size_t)cam * S.kcMax * (S.kcMax + 1);
    for (int i = tid; i < S.kcMax * (S.kcMax + 1); i += blockDim.x) {
        const int a = i / (S.kcMax + 1), b = i % (S.kcMax + 1);
        double s = 0.0;
        if (a < kc && (b < kc || b == S.kcMax)) {
            // column kcMax of cam_sum holds the rhs (cam_partial keeps it at column kc)
            const int bsrc = (b == S.kcMax) ? kc : b;
            for (int img = P.img0; img < P.img1; img++)
                if (P.cam_of_img[img] == cam) s += S.cam_partial[(size_t)img * S.kcMax * (S.kcMax + 1) + a * (S.kcMax + 1) + bsrc];
        }
        out[i] = s;
    }
}

// scatter the camera blocks into N / n
__global__ void __launch_bounds__(256) k_camera_scatter(DevProblem P, AssemblyScratch S, SysView M, double *__restrict__ rhs) {
    const int cam = blockIdx.x, tid = threadIdx.x;
    const int kc = 3 + P.coef_ptr[cam + 1] - P.coef_ptr[cam];
    const int32_t *cc = P.campos_col + P.cam_kbase[cam];
    const double *in = S.cam_sum + (size_t)cam * S.kcMax * (S.kcMax + 1);
    const int d = P.d;
    for (int i = tid; i < kc * (kc + 1); i += blockDim.x) {
        const int a = i / (kc + 1), b = i % (kc + 1);
        if (b < kc && b > a) continue;
        const int32_t ca = cc[a];
        if (!col_active(ca)) continue;
        if (b < kc && !col_active(cc[b])) continue;
        if (b < kc) sys_add(M, ca - d, cc[b] - d, in[a * (S.kcMax + 1) + b]);
        else rhs[ca - d] += in[a * (S.kcMax + 1) + S.kcMax];
    }
}

// ---- by-point sweep -----------------------------------------------------------------------------------------------
// tile columns: 0..2 point X,Y,Z; 3..3+kraw-1 raw parameters (x0,y0,c,coefs per camera) of the cameras of ONE camera group
// [cam0, cam1); 3+kraw = w.  One warp per object point; output rows 0..2 of the skinny Gram: [PP | P x camera | n_P].
// Camera groups: the Gram row is at most 72 columns wide, so a network whose cameras have more than 68 raw parameters in
// total is swept once per group of consecutive cameras (observations of other cameras contribute nothing to a group's
// sweep; every observation belongs to exactly one group, so the point block and n_P add up over the groups).  One group --
// one sweep -- is the common case.  A group of a single camera stages it in shared memory.
constexpr int kPtWarps = 4;

template <int NT, bool ONE_CAM, int MINB, bool STD>
__global__ void __launch_bounds__(kPtWarps * 32, MINB) k_by_point(DevProblem P, int cam0, int cam1, int kraw, double *__restrict__ pt_partial) {
    constexpr int LDT = 68;
    constexpr int NC = 8 * NT;
    extern __shared__ double smem[];
    __shared__ CamSmem cs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (ONE_CAM) {
        load_camera(P, cam0, cs, tid, blockDim.x);
        __syncthreads();
    }
    const int pt = blockIdx.x * kPtWarps + warp;
    if (pt >= P.nPt) return;  // warp-uniform
    double *tile = smem + (size_t)warp * NC * LDT;
#define ROW0(c) tile[(c) * LDT + lane]
#define ROW1(c) tile[(c) * LDT + lane + 32]
    const int wcol = 3 + kraw;
    const int kbase0 = P.cam_kbase[cam0];
    const double X = P.xyz[3 * (int64_t)pt], Y = P.xyz[3 * (int64_t)pt + 1], Z = P.xyz[3 * (int64_t)pt + 2];
    double acc[NT][2];
#pragma unroll
    for (int i = 0; i < NT; i++) acc[i][0] = acc[i][1] = 0.0;
    const int64_t o0 = P.pt_obs_ptr[pt], o1 = P.pt_obs_ptr[pt + 1];
    for (int64_t base = o0; base < o1; base += 32) {
        for (int i = 0; i <= wcol; i++) { ROW0(i) = 0.0; ROW1(i) = 0.0; }
        const int64_t oi = base + lane;
        if (oi < o1) {
            const int64_t j = P.pt_obs[oi];
            const int img = P.img_of_obs[j], cam = P.cam_of_img[img];
            if (cam >= cam0 && cam < cam1) {
                const ImgPose q = load_pose(P.pose, img);
                const CamView cv = ONE_CAM ? view_of(P, cs) : view_global(P, cam);
                const int kb = 3 + P.cam_kbase[cam] - kbase0;
                const double r00 = P.rw[3 * j], r01 = P.rw[3 * j + 1], r11 = P.rw[3 * j + 2];
                const double2 xyo = reinterpret_cast<const double2 *>(P.xy)[j];
                BaseRows r;
                eval_observation_t<false, STD>(q, cv, X, Y, Z, xyo.x, xyo.y, r, [&](int k, double v0, double v1) {
                    ROW0(kb + 3 + k) = r00 * v0 + r01 * v1;
                    ROW1(kb + 3 + k) = r11 * v1;
                });
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    ROW0(i) = r00 * r.ax[i] + r01 * r.ay[i];
                    ROW1(i) = r11 * r.ay[i];
                }
                ROW0(kb) = r00;  ROW1(kb) = 0.0;
                ROW0(kb + 1) = r01;  ROW1(kb + 1) = r11;
                ROW0(kb + 2) = r00 * r.ax[3] + r01 * r.ay[3];  ROW1(kb + 2) = r11 * r.ay[3];
                ROW0(wcol) = r00 * r.w0 + r01 * r.w1;  ROW1(wcol) = r11 * r.w1;
            }
        }
        __syncwarp();
#pragma unroll 4
        for (int ks = 0; ks < 16; ks++) {
            const double *src = tile + (lane >> 2) * LDT + 4 * ks + (lane & 3);
            const double fa = src[0];
#pragma unroll
            for (int b = 0; b < NT; b++) dmma884(acc[b][0], acc[b][1], fa, src[8 * b * LDT]);
        }
        __syncwarp();
    }
#undef ROW0
#undef ROW1
    if ((lane >> 2) < 3) {
        double *out = pt_partial + ((size_t)pt * 3 + (lane >> 2)) * NC;
#pragma unroll
        for (int b = 0; b < NT; b++) {
            out[8 * b + 2 * (lane & 3)] = acc[b][0];
            out[8 * b + 2 * (lane & 3) + 1] = acc[b][1];
        }
    }
}

// scatter the per-point partials of one camera group into N / n (single owner per entry)
__global__ void __launch_bounds__(256) k_point_scatter(DevProblem P, AssemblyScratch S, int NC, int kbase0, int kraw, SysView M,
                                                        double *__restrict__ rhs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)P.nPt * 3 * NC) return;
    const int pt = (int)(i / (3 * NC));
    const int c = (int)((i / NC) % 3), j = (int)(i % NC);
    const int32_t cp = P.pt_col[3 * (int64_t)pt + c];
    if (!col_active(cp)) return;
    const double v = S.pt_partial[i];
    const int d = P.d;
    if (j < 3) {
        if (j > c) return;
        const int32_t c2 = P.pt_col[3 * (int64_t)pt + j];
        if (!col_active(c2)) return;
        sys_add(M, cp - d, c2 - d, v);
    } else if (j < 3 + kraw) {
        const int32_t c2 = P.campos_col[kbase0 + j - 3];
        if (!col_active(c2)) return;
        sys_add(M, cp - d, c2 - d, v);
    } else if (j == 3 + kraw) {
        rhs[cp - d] += v;
    }
}

// resident CTAs per SM the observation sweeps are compiled for (register cap 128 at 4 CTAs of 128 threads): JAICOV_SWEEP_MINB
static int sweep_min_blocks() {
    static const int v = [] { const char *e = getenv("JAICOV_SWEEP_MINB"); return e ? atoi(e) : 4; }();
    return v;
}

template <int NT>
static void run_by_image(const DevProblem &P, const AssemblyScratch &S, const SysView &M, cudaStream_t s) {
    const size_t smem = (size_t)kImgWarps * 8 * NT * 68 * sizeof(double);
    static PerDeviceOnce once3, once4, once4s;
    g_launch_count++;
    if (sweep_min_blocks() >= 4 && NT <= 4 && S.std_eval) {
        once4s.run([&] { JCHECK(cudaFuncSetAttribute(k_by_image<NT, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
        k_by_image<NT, 4, true><<<S.nWork, kImgWarps * 32, smem, s>>>(P, S.work, S.img_partial, M);
    } else if (sweep_min_blocks() >= 4 && NT <= 4) {
        once4.run([&] { JCHECK(cudaFuncSetAttribute(k_by_image<NT, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
        k_by_image<NT, 4, false><<<S.nWork, kImgWarps * 32, smem, s>>>(P, S.work, S.img_partial, M);
    } else {
        once3.run([&] { JCHECK(cudaFuncSetAttribute(k_by_image<NT, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
        k_by_image<NT, 1, false><<<S.nWork, kImgWarps * 32, smem, s>>>(P, S.work, S.img_partial, M);
    }
}

template <int NT, bool ONE_CAM>
static void run_by_point(const DevProblem &P, const AssemblyScratch &S, const PtGroup &g, cudaStream_t s) {
    const size_t smem = (size_t)kPtWarps * 8 * NT * 68 * sizeof(double);
    static PerDeviceOnce once3, once4, once4s;
    g_launch_count++;
    if (sweep_min_blocks() >= 4 && NT <= 4 && ONE_CAM && S.std_eval) {
        once4s.run([&] { JCHECK(cudaFuncSetAttribute(k_by_point<NT, ONE_CAM, 4, ONE_CAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
        k_by_point<NT, ONE_CAM, 4, ONE_CAM><<<(P.nPt + kPtWarps - 1) / kPtWarps, kPtWarps * 32, smem, s>>>(P, g.cam0, g.cam1, g.kraw, S.pt_partial);
    } else if (sweep_min_blocks() >= 4 && NT <= 4) {
        once4.run([&] { JCHECK(cudaFuncSetAttribute(k_by_point<NT, ONE_CAM, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
        k_by_point<NT, ONE_CAM, 4, false><<<(P.nPt + kPtWarps - 1) / kPtWarps, kPtWarps * 32, smem, s>>>(P, g.cam0, g.cam1, g.kraw, S.pt_partial);
    } else {
        once3.run([&] { JCHECK(cudaFuncSetAttribute(k_by_point<NT, ONE_CAM, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
        k_by_point<NT, ONE_CAM, 1, false><<<(P.nPt + kPtWarps - 1) / kPtWarps, kPtWarps * 32, smem, s>>>(P, g.cam0, g.cam1, g.kraw, S.pt_partial);
    }
}

#define JAICOV_DISPATCH_NT(nt, CALL)                 \
    switch (nt) {                                    \
        case 1: CALL(1); break;                      \
        case 2: CALL(2); break;                      \
        case 3: CALL(3); break;                      \
        case 4: CALL(4); break;                      \
        case 5: CALL(5); break;                      \
        case 6: CALL(6); break;                      \
        case 7: CALL(7); break;                      \
        case 8: CALL(8); break;                      \
        case 9: CALL(9); break;                      \
        default: throw CudaError{cudaErrorInvalidValue, "too many camera parameters for one Gram tile row", __FILE__, __LINE__}; \
    }

// image points, first half: sweeps over this rank's images (unique EO blocks straight into N, per-camera sums).
// M and rhs must be zero on entry.
void launch_assemble_images(const DevProblem &P, const AssemblyScratch &S, const SysView &M, double *rhs, cudaStream_t s) {
    if (P.m == 0) return;
    if (S.nWork > 0) {
#define CALL_IMG(NT) run_by_image<NT>(P, S, M, s)
        JAICOV_DISPATCH_NT(S.ntImg, CALL_IMG)
#undef CALL_IMG
        const int NC = 8 * S.ntImg;
        g_launch_count++;
        k_image_finalize<<<P.img1 - P.img0, 256, NC * NC * sizeof(double), s>>>(P, S, M, rhs);
    }
    g_launch_count++;
    k_camera_sum<<<P.nCam, 256, 0, s>>>(P, S);
}

// per-point partial sums of one camera group over this rank's observations (-> S.pt_partial, all-reduced across ranks)
void launch_by_point(const DevProblem &P, const AssemblyScratch &S, const PtGroup &g, cudaStream_t s) {
    if (P.m == 0) return;
    if (g.cam1 - g.cam0 == 1) {
#define CALL_PT(NT) run_by_point<NT, true>(P, S, g, s)
        JAICOV_DISPATCH_NT(g.nt, CALL_PT)
#undef CALL_PT
    } else {
#define CALL_PT(NT) run_by_point<NT, false>(P, S, g, s)
        JAICOV_DISPATCH_NT(g.nt, CALL_PT)
#undef CALL_PT
    }
}

void launch_camera_scatter(const DevProblem &P, const AssemblyScratch &S, const SysView &M, double *rhs, cudaStream_t s) {
    if (P.m == 0) return;
    g_launch_count++;
    k_camera_scatter<<<P.nCam, 256, 0, s>>>(P, S, M, rhs);
}

void launch_point_scatter(const DevProblem &P, const AssemblyScratch &S, const PtGroup &g, int kbase0, const SysView &M, double *rhs,
                          cudaStream_t s) {
    if (P.m == 0) return;
    const int NC = 8 * g.nt;
    const int64_t tot = (int64_t)P.nPt * 3 * NC;
    g_launch_count++;
    k_point_scatter<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(P, S, NC, kbase0, g.kraw, M, rhs);
}

// R with P = R'R of every image point (PDF:296-319): constant over the passes of one adjustment
__global__ void __launch_bounds__(256) k_obs_weights(DevProblem P, double *__restrict__ rw) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.m) return;
    double p00, p01, p11;
    point_weight(P.sigma2, P.var[2 * j], P.var[2 * j + 1], P.rho[j], p00, p01, p11);
    const double r00 = sqrt(p00), r01 = p01 / r00;
    rw[3 * j] = r00;  rw[3 * j + 1] = r01;  rw[3 * j + 2] = sqrt(p11 - r01 * r01);
}

void launch_obs_weights(const DevProblem &P, double *rw, cudaStream_t s) {
    if (P.m == 0) return;
    g_launch_count++;
    k_obs_weights<<<(unsigned)((P.m + 255) / 256), 256, 0, s>>>(P, rw);
}

// ---- scale bars (PDF:210-283): a handful of observations, one thread, reference order ------------------------
__global__ void k_scale_bars(DevProblem P, SysView M, double *__restrict__ rhs) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int d = P.d;
    for (int b = 0; b < P.nBar; b++) {
        const double *A = P.xyz + 3 * (int64_t)P.bar_a[b], *B = P.xyz + 3 * (int64_t)P.bar_b[b];
        const double dX = B[0] - A[0], dY = B[1] - A[1], dZ = B[2] - A[2];
        const double len = sqrt(dX * dX + dY * dY + dZ * dZ);
        const double a[6] = {-dX / len, -dY / len, -dZ / len, dX / len, dY / len, dZ / len};
        int32_t col[6];
        for (int i = 0; i < 3; i++) { col[i] = P.pt_col[3 * (int64_t)P.bar_a[b] + i]; col[3 + i] = P.pt_col[3 * (int64_t)P.bar_b[b] + i]; }
        const double Pw = P.sigma2 / P.bar_var[b];
        const double w = P.bar_len[b] - len;
        for (int i = 0; i < 6; i++) {
            if (!col_active(col[i])) continue;
            rhs[col[i] - d] += a[i] * Pw * w;
            for (int j = 0; j <= i; j++) {
                if (!col_active(col[j])) continue;
                sys_add(M, col[i] - d, col[j] - d, a[i] * Pw * a[j]);
            }
        }
    }
}

void launch_scale_bars(const DevProblem &P, const SysView &M, double *rhs, cudaStream_t s) {
    if (P.nBar == 0) return;
    g_launch_count++;
    k_scale_bars<<<1, 32, 0, s>>>(P, M, rhs);
}

__global__ void k_omega_bars(DevProblem P, const double *__restrict__ dxref, double *__restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double om = 0.0;
    for (int b = 0; b < P.nBar; b++) {
        const double *A = P.xyz + 3 * (int64_t)P.bar_a[b], *B = P.xyz + 3 * (int64_t)P.bar_b[b];
        const double dX = B[0] - A[0], dY = B[1] - A[1], dZ = B[2] - A[2];
        const double len = sqrt(dX * dX + dY * dY + dZ * dZ);
        const double a[6] = {-dX / len, -dY / len, -dZ / len, dX / len, dY / len, dZ / len};
        double v = P.bar_len[b] - len;
        for (int i = 0; i < 3; i++) {
            const int32_t ca = P.pt_col[3 * (int64_t)P.bar_a[b] + i], cb = P.pt_col[3 * (int64_t)P.bar_b[b] + i];
            if (col_active(ca)) v -= a[i] * dxref[ca];
            if (col_active(cb)) v -= a[3 + i] * dxref[cb];
        }
        om += v * v * (P.sigma2 / P.bar_var[b]);
    }
    out[0] = om;
}

void launch_omega_bars(const DevProblem &P, const double *dxref, double *omega_out, cudaStream_t s) {
    g_launch_count++;
    k_omega_bars<<<1, 32, 0, s>>>(P, dxref, omega_out);
}

// ---- K8: Omega = sum v'Pv, v = w - A dx (BA:472-491) -----------------------------------------------------------------------
// One CTA per (image, chunk of <= 1024 points) -- the work items of the by-image sweep -- so that everything an observation
// shares with its neighbours sits in shared memory: the camera, and the dx of the image's exterior orientation and of the
// camera parameters (0 where a parameter is fixed: no per-observation column tests).  The dx of the object points is gathered
// once per pass into a per-point array.  v'Pv = |R v|^2 with the precomputed R.  Block partial sums, then one warp adds them
// in a fixed order (bitwise reproducible).
constexpr int kOmegaThreads = 128;

__global__ void __launch_bounds__(256) k_gather_point_dx(DevProblem P, const double *__restrict__ dxref, double *__restrict__ dxp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * (int64_t)P.nPt) return;
    const int32_t c = P.pt_col[i];
    dxp[i] = col_active(c) ? dxref[c] : 0.0;
}

template <int MINB, bool STD>
__global__ void __launch_bounds__(kOmegaThreads, MINB) k_omega(DevProblem P, const WorkItem *__restrict__ work, const double *__restrict__ dxref,
                                                         const double *__restrict__ dxp, double *__restrict__ partial) {
    __shared__ CamSmem cs;
    __shared__ double s_dxe[6], s_dxc[3 + kMaxCoef];
    __shared__ double s_red[kOmegaThreads / 32];
    __shared__ ImgPose s_pose;
    const int tid = threadIdx.x;
    const WorkItem wi = work[blockIdx.x];
    const int img = wi.img, cam = P.cam_of_img[img];
    load_camera(P, cam, cs, tid, blockDim.x);
    if (tid < 14) reinterpret_cast<double *>(&s_pose)[tid] = P.pose[(int64_t)img * kPoseStride + tid];
    if (tid < 6) {
        const int32_t c = P.eo_col[6 * (int64_t)img + tid];
        s_dxe[tid] = col_active(c) ? dxref[c] : 0.0;
    }
    {
        const int nraw = 3 + P.coef_ptr[cam + 1] - P.coef_ptr[cam];
        const int32_t *cc = P.campos_col + P.cam_kbase[cam];
        for (int k = tid; k < nraw; k += blockDim.x) s_dxc[k] = col_active(cc[k]) ? dxref[cc[k]] : 0.0;
    }
    __syncthreads();
    const CamView cv = view_of(P, cs);
    const ImgPose &q = s_pose;
    const double ex = s_dxe[0], ey = s_dxe[1], ez = s_dxe[2], eo = s_dxe[3], ep = s_dxe[4], ek = s_dxe[5];
    const double cx0 = s_dxc[0], cy0 = s_dxc[1], cc0 = s_dxc[2];
    double local = 0.0;
    for (int64_t j = wi.begin + tid; j < wi.end; j += kOmegaThreads) {
        const int pt = P.obj_idx[j];
        const double2 xyo = reinterpret_cast<const double2 *>(P.xy)[j];
        const double r00 = P.rw[3 * j], r01 = P.rw[3 * j + 1], r11 = P.rw[3 * j + 2];
        const double *xp = P.xyz + 3 * (int64_t)pt, *dp = dxp + 3 * (int64_t)pt;
        double s0 = 0.0, s1 = 0.0;  // (A dx) rows
        BaseRows r;
        eval_observation_t<true, STD>(q, cv, xp[0], xp[1], xp[2], xyo.x, xyo.y, r, [&](int k, double v0, double v1) {
            const double dx = s_dxc[3 + k];
            s0 += v0 * dx;  s1 += v1 * dx;
        });
        const double d0 = dp[0] - ex, d1 = dp[1] - ey, d2 = dp[2] - ez;      // X0,Y0,Z0 columns = -(X,Y,Z) columns
        s0 += r.ax[0] * d0 + r.ax[1] * d1 + r.ax[2] * d2 + r.ax[4] * eo + r.ax[5] * ep + r.ax[6] * ek + cx0 + r.ax[3] * cc0;
        s1 += r.ay[0] * d0 + r.ay[1] * d1 + r.ay[2] * d2 + r.ay[4] * eo + r.ay[5] * ep + r.ay[6] * ek + cy0 + r.ay[3] * cc0;
        const double v0 = r.w0 - s0, v1 = r.w1 - s1;
        const double t0 = r00 * v0 + r01 * v1, t1 = r11 * v1;
        local += t0 * t0 + t1 * t1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = local;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < kOmegaThreads / 32; i++) s += s_red[i];
        partial[blockIdx.x] = s;
    }
}

__global__ void k_omega_final(const double *__restrict__ partial, int n, double *__restrict__ out) {
    // one warp, fixed order
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = s;
}

// omega_out[0] = image points part
void launch_omega(const DevProblem &P, const AssemblyScratch &S, const double *dxref, double *omega_out, cudaStream_t s) {
    if (P.obs1 <= P.obs0 || S.nWork == 0) {
        cudaMemsetAsync(omega_out, 0, sizeof(double), s);
        return;
    }
    g_launch_count++;
    k_gather_point_dx<<<(unsigned)((3 * (int64_t)P.nPt + 255) / 256), 256, 0, s>>>(P, dxref, S.dxp);
    g_launch_count++;
    if (sweep_min_blocks() >= 4 && S.std_eval) k_omega<4, true><<<S.nWork, kOmegaThreads, 0, s>>>(P, S.work, dxref, S.dxp, S.omega_partial);
    else if (sweep_min_blocks() >= 4) k_omega<4, false><<<S.nWork, kOmegaThreads, 0, s>>>(P, S.work, dxref, S.dxp, S.omega_partial);
    else k_omega<1, false><<<S.nWork, kOmegaThreads, 0, s>>>(P, S.work, dxref, S.dxp, S.omega_partial);
    g_launch_count++;
    k_omega_final<<<1, 32, 0, s>>>(S.omega_partial, S.nWork, omega_out);
}

// ---- matrix-free product with the bordered normal matrix (verification entry point jaicov_normal_product) -------------
// Y_v += A' P (A x_v) over this rank's image points, straight from the observations: nothing of the assembled N, its
// factor or its inverse is read, so  K Qxx e_c = e_c  and  K [lambda; dx] = [0; n]  can be checked at ANY size (the CPU
// oracle stops at n ~ 2e4) and for any storage layout of the factorisation.  Same model code as the assembly
// (eval_observation); FP64 atomics (order of summation is free here: the result is only ever compared to a tolerance).
// X, Y: [nv][n] in reference column numbering (border first); rhs (may be null): += A' P w; wpw (may be null): += w' P w.
constexpr int kProdThreads = 128;

__global__ void __launch_bounds__(kProdThreads) k_normal_product(DevProblem P, int nv, int64_t n, const double *__restrict__ X,
                                                                 double *__restrict__ Y, double *__restrict__ rhs,
                                                                 double *__restrict__ wpw) {
    __shared__ double s_red[kProdThreads / 32];
    double local = 0.0;
    for (int64_t j = P.obs0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P.obs1; j += (int64_t)gridDim.x * blockDim.x) {
        const int img = P.img_of_obs[j], cam = P.cam_of_img[img], pt = P.obj_idx[j];
        if (P.img_dense && P.img_dense[img]) continue;      // weighted by the image's dense dispersion: dense_sigma.cu
        const ImgPose q = load_pose(P.pose, img);
        const CamView cv = view_global(P, cam);
        const int32_t *ccol = P.coef_col + P.coef_ptr[cam];
        double a0[12 + kMaxCoef], a1[12 + kMaxCoef];
        int32_t col[12 + kMaxCoef];
        BaseRows r;
        eval_observation(q, cv, P.xyz[3 * (int64_t)pt], P.xyz[3 * (int64_t)pt + 1], P.xyz[3 * (int64_t)pt + 2],
                         P.xy[2 * j], P.xy[2 * j + 1], r, [&](int k, double v0, double v1) {
                             a0[12 + k] = v0; a1[12 + k] = v1; col[12 + k] = ccol[k];
                         });
        const int ns = 12 + cv.ncoef;
        const int32_t *pc = P.pt_col + 3 * (int64_t)pt, *ic = P.io_col + 3 * cam, *ec = P.eo_col + 6 * (int64_t)img;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            a0[i] = r.ax[i];  a1[i] = r.ay[i];  col[i] = pc[i];
            a0[6 + i] = -r.ax[i];  a1[6 + i] = -r.ay[i];  col[6 + i] = ec[i];
            a0[9 + i] = r.ax[4 + i];  a1[9 + i] = r.ay[4 + i];  col[9 + i] = ec[3 + i];
            col[3 + i] = ic[i];
        }
        a0[3] = 1.0; a1[3] = 0.0; a0[4] = 0.0; a1[4] = 1.0; a0[5] = r.ax[3]; a1[5] = r.ay[3];
        double p00, p01, p11;
        point_weight(P.sigma2, P.var[2 * j], P.var[2 * j + 1], P.rho[j], p00, p01, p11);
        if (rhs) {
            const double t0 = p00 * r.w0 + p01 * r.w1, t1 = p01 * r.w0 + p11 * r.w1;
            for (int s = 0; s < ns; s++)
                if (col_active(col[s])) atomicAdd(rhs + col[s], a0[s] * t0 + a1[s] * t1);
            local += r.w0 * t0 + r.w1 * t1;
        }
        for (int v = 0; v < nv; v++) {
            const double *x = X + (int64_t)v * n;
            double s0 = 0.0, s1 = 0.0;
            for (int s = 0; s < ns; s++)
                if (col_active(col[s])) { const double xv = x[col[s]]; s0 += a0[s] * xv; s1 += a1[s] * xv; }
            const double t0 = p00 * s0 + p01 * s1, t1 = p01 * s0 + p11 * s1;
            double *y = Y + (int64_t)v * n;
            for (int s = 0; s < ns; s++)
                if (col_active(col[s])) atomicAdd(y + col[s], a0[s] * t0 + a1[s] * t1);
        }
    }
    if (wpw) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int i = 0; i < kProdThreads / 32; i++) s += s_red[i];
            atomicAdd(wpw, s);
        }
    }
}

// scale bars (PDF:210-283) and the datum border rows (BA:493-635): K = [[0, B], [B', N]]
__global__ void k_product_bars_border(DevProblem P, int nv, int64_t n, const double *__restrict__ X, double *__restrict__ Y,
                                      double *__restrict__ rhs, double *__restrict__ wpw, const double *__restrict__ Bt, int64_t ldb,
                                      int with_bars) {
    const int tid = threadIdx.x;
    if (with_bars && tid == 0) {
        double om = 0.0;
        for (int b = 0; b < P.nBar; b++) {
            const double *A = P.xyz + 3 * (int64_t)P.bar_a[b], *B = P.xyz + 3 * (int64_t)P.bar_b[b];
            const double dX = B[0] - A[0], dY = B[1] - A[1], dZ = B[2] - A[2];
            const double len = sqrt(dX * dX + dY * dY + dZ * dZ);
            const double a[6] = {-dX / len, -dY / len, -dZ / len, dX / len, dY / len, dZ / len};
            int32_t col[6];
            for (int i = 0; i < 3; i++) { col[i] = P.pt_col[3 * (int64_t)P.bar_a[b] + i]; col[3 + i] = P.pt_col[3 * (int64_t)P.bar_b[b] + i]; }
            const double Pw = P.sigma2 / P.bar_var[b], w = P.bar_len[b] - len;
            if (rhs) {
                for (int i = 0; i < 6; i++)
                    if (col_active(col[i])) rhs[col[i]] += a[i] * Pw * w;
                om += w * Pw * w;
            }
            for (int v = 0; v < nv; v++) {
                double s = 0.0;
                for (int i = 0; i < 6; i++)
                    if (col_active(col[i])) s += a[i] * X[(int64_t)v * n + col[i]];
                for (int i = 0; i < 6; i++)
                    if (col_active(col[i])) Y[(int64_t)v * n + col[i]] += a[i] * Pw * s;
            }
        }
        if (wpw) wpw[0] += om;
    }
    __syncthreads();
    // border: Y[a] = sum_c B[a][c] x[d + c];  Y[d + c] += sum_a B[a][c] x[a]   (Bt: d rows of internal columns, ldb apart)
    const int d = P.d;
    for (int v = 0; v < nv; v++) {
        const double *x = X + (int64_t)v * n;
        double *y = Y + (int64_t)v * n;
        for (int a = 0; a < d; a++) {
            __shared__ double red[256];
            double s = 0.0;
            for (int c = tid; c < P.u; c += blockDim.x) s += Bt[a * ldb + c] * x[d + c];
            red[tid] = s;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int i = 0; i < (int)blockDim.x; i++) t += red[i];
                y[a] += t;
            }
            __syncthreads();
        }
        for (int c = tid; c < P.u; c += blockDim.x) {
            double s = 0.0;
            for (int a = 0; a < d; a++) s += Bt[a * ldb + c] * x[a];
            y[d + c] += s;
        }
    }
}

void launch_normal_product_points(const DevProblem &P, int nv, int64_t n, const double *X, double *Y, double *rhs, double *wpw,
                                  cudaStream_t s) {
    if (P.obs1 <= P.obs0) return;
    const int64_t nobs = P.obs1 - P.obs0;
    const int blocks = (int)std::min<int64_t>(148 * 8, (nobs + kProdThreads - 1) / kProdThreads);
    g_launch_count++;
    k_normal_product<<<blocks, kProdThreads, 0, s>>>(P, nv, n, X, Y, rhs, wpw);
}

void launch_normal_product_rest(const DevProblem &P, int nv, int64_t n, const double *X, double *Y, double *rhs, double *wpw,
                                const double *Bt, int64_t ldb, cudaStream_t s) {
    g_launch_count++;
    k_product_bars_border<<<1, 256, 0, s>>>(P, nv, n, X, Y, rhs, wpw, Bt, ldb, P.nBar > 0);
}

}  // namespace jaicov
