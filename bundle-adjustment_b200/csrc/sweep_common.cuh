// sweep_common.cuh -- device helpers shared by the observation sweeps (assembly.cu, dense_sigma.cu): camera staging in shared
// memory, pose table access, index helpers.
#pragma once
#include "common.h"
#include "model.cuh"

namespace jaicov {

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ bool col_active(int32_t c) { return c >= 0 && c != JAICOV_COL_FIXED; }

// element (r,c) of the symmetric system stored as lower triangle, row-major, internal indices (col - d)
__device__ __forceinline__ int64_t lower_idx(int64_t a, int64_t b, int64_t ld) {
    return a >= b ? a * ld + b : b * ld + a;
}

struct CamSmem {
    double io[3];
    double r0;
    double val[kMaxCoef], r0pow[kMaxCoef];
    int32_t type[kMaxCoef], order[kMaxCoef], zm[kMaxCoef], zptr[kMaxCoef + 1];
    int32_t ncoef;
    int32_t std5[5];     // hasC, hasB, nB, nA, nD of a canonical coefficient list (model.cuh: STD evaluation); hasC < 0: not canonical
};

__device__ __forceinline__ void load_camera(const DevProblem &P, int cam, CamSmem &s, int tid, int nthreads) {
    const int c0 = P.coef_ptr[cam], c1 = P.coef_ptr[cam + 1];
    if (tid < 3) s.io[tid] = P.io_val[3 * cam + tid];
    if (tid == 3) { s.r0 = P.r0[cam]; s.ncoef = c1 - c0; }
    if (tid >= 4 && tid < 9) s.std5[tid - 4] = P.cam_std[5 * cam + tid - 4];
    for (int k = tid; k < c1 - c0; k += nthreads) {
        s.val[k] = P.coef_val[c0 + k];
        s.r0pow[k] = P.coef_r0pow[c0 + k];
        s.type[k] = P.coef_type[c0 + k];
        s.order[k] = P.coef_order[c0 + k];
        s.zm[k] = P.zern_m[c0 + k];
    }
    for (int k = tid; k <= c1 - c0; k += nthreads) s.zptr[k] = P.zern_ptr[c0 + k];
}

__device__ __forceinline__ CamView view_of(const DevProblem &P, const CamSmem &s) {
    CamView v;
    v.io = s.io; v.r0 = s.r0; v.ncoef = s.ncoef; v.type = s.type; v.order = s.order; v.val = s.val; v.r0pow = s.r0pow;
    v.zern_m = s.zm; v.zern_ptr = s.zptr; v.zern_p = P.zern_p; v.zern_c = P.zern_c;
    v.hasC = s.std5[0]; v.hasB = s.std5[1]; v.nB = s.std5[2]; v.nA = s.std5[3]; v.nD = s.std5[4];
    return v;
}

__device__ __forceinline__ CamView view_global(const DevProblem &P, int cam) {
    const int c0 = P.coef_ptr[cam];
    CamView v;
    v.io = P.io_val + 3 * cam; v.r0 = P.r0[cam]; v.ncoef = P.coef_ptr[cam + 1] - c0;
    v.type = P.coef_type + c0; v.order = P.coef_order + c0; v.val = P.coef_val + c0; v.r0pow = P.coef_r0pow + c0;
    v.zern_m = P.zern_m + c0; v.zern_ptr = P.zern_ptr + c0; v.zern_p = P.zern_p; v.zern_c = P.zern_c;
    return v;
}

__device__ __forceinline__ ImgPose load_pose(const double *pose, int img) {
    const double *p = pose + (int64_t)img * kPoseStride;
    ImgPose q;
    q.r11 = p[0]; q.r12 = p[1]; q.r13 = p[2]; q.r21 = p[3]; q.r22 = p[4]; q.r23 = p[5];
    q.r31 = p[6]; q.r32 = p[7]; q.r33 = p[8]; q.sinK = p[9]; q.cosK = p[10]; q.X0 = p[11]; q.Y0 = p[12]; q.Z0 = p[13];
    return q;
}

}  // namespace jaicov
