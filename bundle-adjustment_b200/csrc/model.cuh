// model.cuh -- functional model of one image observation on the device.
//
// Collinearity equations and their partials (reference: bundle/derivation/PartialDerivativeFactory.java:94-190),
// the chain rule of the distortion models (DistortionModelFactory.java:33-101) and the five model families
// (RadiallySymmetric...:39-90, Tangential...:39-134, AffinityShear...:37-81, RadialDistance...:39-161,
// Zernike...:41-227), written for one thread per observation.
//
// Design notes (not a port of the Java):
//  * the per-image trigonometry is hoisted into an ImgPose table (one sincos per image and pass, not per point);
//  * the reference's dense 2 x n row pair is never formed.  Only 7 "base" partial pairs are tracked in registers
//    (X,Y,Z,c,omega,phi,kappa): the X0,Y0,Z0 entries are the exact negatives of X,Y,Z at every step (the chain
//    rule is linear and dN/dX0 = -dN/dX), and x0,y0 are the constants (1,0),(0,1) no model touches
//    (DistortionModelFactory never updates them);
//  * each coefficient's own column is handed to a caller-supplied sink (shared-memory tile, global buffer or a dot
//    product), so no dynamically indexed register array exists;
//  * integer powers are repeated multiplications; Zernike radial terms come from a host-built table that uses the
//    reference's integer arithmetic (long division p/2, binomials) verbatim.
#pragma once
#include <cstdint>

#include "../../include/jaicov_b200.h"

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace jaicov {

constexpr int kPoseStride = 16;  // doubles per image in the pose table

// r11,r12,r13,r21,r22,r23,r31,r32,r33, sinKappa, cosKappa, X0, Y0, Z0
struct ImgPose {
    double r11, r12, r13, r21, r22, r23, r31, r32, r33, sinK, cosK, X0, Y0, Z0;
};

// view of one camera's parameters (pointers may address shared or global memory)
struct CamView {
    const double *io;         // x0, y0, c
    double r0;
    int ncoef;
    const int32_t *type;      // [ncoef] ParameterType id
    const int32_t *order;     // [ncoef]
    const double *val;        // [ncoef]
    const double *r0pow;      // [ncoef] r0^(2 order): the constant subtracted by the radial / distance polynomials (:58,66)
    // canonical coefficient list (STD evaluation): [Cx Cy]? [Bx By B1..BnB]? [A1..AnA] [D1..DnD], orders consecutive from 1, no Zernike
    int hasC = 0, hasB = 0, nB = 0, nA = 0, nD = 0;
    // Zernike table (global): per coefficient (global index base k0): m, term range; per term: exponent p, coefficient c
    const int32_t *zern_m;    // indexed by local coefficient position
    const int32_t *zern_ptr;  // [ncoef+1] local positions -> term range
    const int32_t *zern_p;
    const double *zern_c;
};

struct BaseRows {
    // compact base slots: 0,1,2 = X,Y,Z; 3 = c; 4,5,6 = omega,phi,kappa
    double ax[7], ay[7];
    double w0, w1;
};

__host__ __device__ __forceinline__ double ipow(double x, int e) {
    // Math.pow(x, e) for integer e (RadiallySymmetricDistortionModelFactory.java:58,66 etc.)
    bool neg = e < 0;
    unsigned n = neg ? (unsigned)(-e) : (unsigned)e;
    double r = 1.0, b = x;
    while (n) {
        if (n & 1u) r *= b;
        b *= b;
        n >>= 1;
    }
    return neg ? 1.0 / r : r;
}

__host__ __device__ __forceinline__ ImgPose make_pose(const double *eo) {
    double so, co, sp, cp, sk, ck;
#ifdef __CUDA_ARCH__
    sincos(eo[3], &so, &co);
    sincos(eo[4], &sp, &cp);
    sincos(eo[5], &sk, &ck);
#else
    so = sin(eo[3]); co = cos(eo[3]); sp = sin(eo[4]); cp = cos(eo[4]); sk = sin(eo[5]); ck = cos(eo[5]);
#endif
    ImgPose q;
    // Rotation, PartialDerivativeFactory.java:125-135
    q.r11 = cp * ck;  q.r12 = -cp * sk;  q.r13 = sp;
    q.r21 = co * sk + so * sp * ck;  q.r22 = co * ck - so * sp * sk;  q.r23 = -so * cp;
    q.r31 = so * sk - co * sp * ck;  q.r32 = so * ck + co * sp * sk;  q.r33 = co * cp;
    q.sinK = sk; q.cosK = ck; q.X0 = eo[0]; q.Y0 = eo[1]; q.Z0 = eo[2];
    return q;
}

// Evaluate one observation.  sink(k, v0, v1) receives the own-column entries of coefficient k (local position in
// the camera's list) exactly once per coefficient.
//
// Arithmetic (round 2; same formulas, fewer FP64 instructions -- the three sweeps that call this are FP64-issue bound, ncu in
// profiles/r02_*): one reciprocal of N instead of 14 divisions by it; every model's 2 x 2 chain-rule factor d(delta)/d(xs, ys) is
// SUMMED first and applied to the base partials once at the end -- DistortionModelFactory.apply adds
// (d delta_x / d xs) px_i + (d delta_x / d ys) py_i with the BASE partials px, py for every model (:33-101), so the sum over models
// factors out exactly; the explicit N-chain of the distance model (:105-159) factors the same way; r0^(2e) comes from a per-
// coefficient table (it is a constant of the camera).  Results differ from the term-by-term order by rounding only (~1e-16
// relative; the per-entry parity test against the oracle holds at 1e-11).  WITH_EO = false skips the omega / phi / kappa partials
// (the by-point sweep does not use them).
template <bool WITH_EO, bool STD, class Sink>
__host__ __device__ __forceinline__ void eval_observation_t(const ImgPose &q, const CamView &cam, double X, double Y, double Z,
                                                            double xobs, double yobs, BaseRows &r, Sink &&sink) {
    constexpr int NB = WITH_EO ? 7 : 4;
    const double x0 = cam.io[0], y0 = cam.io[1], c = cam.io[2];
    const double dX = X - q.X0, dY = Y - q.Y0, dZ = Z - q.Z0;
    const double kx = q.r11 * dX + q.r21 * dY + q.r31 * dZ;
    const double ky = q.r12 * dX + q.r22 * dY + q.r32 * dZ;
    const double N = q.r13 * dX + q.r23 * dY + q.r33 * dZ;
    const double iN = 1.0 / N;
    const double kxN = kx * iN, kyN = ky * iN;
    const double xs = -c * kxN, ys = -c * kyN;

    // base partials, PartialDerivativeFactory.java:157-189
    double px[NB], py[NB];
    px[0] = -(q.r13 * xs + c * q.r11) * iN;
    px[1] = -(q.r23 * xs + c * q.r21) * iN;
    px[2] = -(q.r33 * xs + c * q.r31) * iN;
    px[3] = -kxN;
    py[0] = -(q.r13 * ys + c * q.r12) * iN;
    py[1] = -(q.r23 * ys + c * q.r22) * iN;
    py[2] = -(q.r33 * ys + c * q.r32) * iN;
    py[3] = -kyN;
    if (WITH_EO) {
        const double tO = q.r33 * dY - q.r23 * dZ;
        const double tP = ky * q.sinK - kx * q.cosK;
        px[NB - 3] = (xs * tO + c * (q.r31 * dY - q.r21 * dZ)) * iN;
        px[NB - 2] = (xs * tP + c * N * q.cosK) * iN;
        px[NB - 1] = ys;
        py[NB - 3] = (ys * tO + c * (q.r32 * dY - q.r22 * dZ)) * iN;
        py[NB - 2] = (ys * tP - c * N * q.sinK) * iN;
        py[NB - 1] = -xs;
    }
    double w0 = xobs - (x0 + xs), w1 = yobs - (y0 + ys);
    // accumulated chain-rule factors: [ax; ay] = [[Dxx, Dxy], [Dyx, Dyy]] [px; py] + [gx; gy] dN/dp
    double Dxx = 1.0, Dxy = 0.0, Dyx = 0.0, Dyy = 1.0, gxs = 0.0, gys = 0.0;
    auto chain = [&](double dlx, double dly, double dXxs, double dXys, double dYxs, double dYys) {
        w0 -= dlx;  w1 -= dly;
        Dxx += dXxs;  Dxy += dXys;  Dyx += dYxs;  Dyy += dYys;
    };

    const double r2 = xs * xs + ys * ys;
    const double xxs2 = 2.0 * xs * xs, yys2 = 2.0 * ys * ys, xys2 = 2.0 * xs * ys;
    const double r02 = cam.r0 * cam.r0;
    // r2^(e) for the polynomial models: the orders of consecutive coefficients usually differ by one, so the last power is
    // kept and extended by one multiplication instead of a square-and-multiply loop per coefficient
    int pe = 0;
    double pv = 1.0;
    auto r2pow = [&](int e) {
        if (e == pe + 1) { pv *= r2; pe = e; }
        else if (e != pe) { pv = ipow(r2, e); pe = e; }
        return pv;
    };
    if (STD) {
        // Straight-line evaluation of the canonical coefficient list (what AICONReportFileReader builds, :308): the same formulas and
        // the same order of operations as the interpreter below, without its per-coefficient type dispatch, power look-ups and
        // int -> double conversions.  The loop counts are per camera, i.e. uniform over the CTA.
        int k = 0;
        if (cam.hasC) {  // AffinityShearDistortionModelFactory.java:37-81
            const double cx = cam.val[0], cy = cam.val[1];
            chain(cx * xs + cy * ys, 0.0, cx, cy, 0.0, 0.0);
            sink(0, xs, 0.0);
            sink(1, ys, 0.0);
            k = 2;
        }
        if (cam.hasB) {  // TangentialDistortionModelFactory.java:39-134
            const double bx = cam.val[k], by = cam.val[k + 1];
            double sum = 1.0;
            const double dlx = bx * (r2 + xxs2) + by * xys2;
            const double dly = by * (r2 + yys2) + bx * xys2;
            const double dXxs = 2.0 * (3.0 * bx * xs + by * ys);
            const double dXys = 2.0 * (by * xs + bx * ys);
            const double dYxs = dXys;
            const double dYys = 2.0 * (bx * xs + 3.0 * by * ys);
            chain(dlx, dly, dXxs, dXys, dYxs, dYys);
            double rim1 = 1.0, ef = 1.0;
            for (int i = 0; i < cam.nB; i++) {
                const double bi = cam.val[k + 2 + i];
                const double ri = rim1 * r2;
                const double dT = bi * ri;
                sum += dT;
                const double cT = 2.0 * bi * ef * rim1;
                const double cTx = dlx * cT, cTy = dly * cT;
                chain(dlx * dT, dly * dT, dT * dXxs + xs * cTx, dT * dXys + ys * cTx, dT * dYxs + xs * cTy, dT * dYys + ys * cTy);
                sink(k + 2 + i, dlx * ri, dly * ri);
                rim1 = ri;
                ef += 1.0;
            }
            sink(k, sum * (r2 + xxs2), sum * xys2);
            sink(k + 1, sum * xys2, sum * (r2 + yys2));
            k += 2 + cam.nB;
        }
        {   // RadiallySymmetricDistortionModelFactory.java:39-90
            double rim1 = 1.0, ef = 1.0;
            for (int i = 0; i < cam.nA; i++, k++) {
                const double ai = cam.val[k];
                const double ri = rim1 * r2;
                const double dRi = ri - cam.r0pow[k];
                const double dRad = ai * dRi;
                const double cR = ai * ef * rim1;
                chain(xs * dRad, ys * dRad, xxs2 * cR + dRad, xys2 * cR, xys2 * cR, yys2 * cR + dRad);
                sink(k, xs * dRi, ys * dRi);
                rim1 = ri;
                ef += 1.0;
            }
        }
        {   // RadialDistanceDistortionModelFactory.java:39-161
            double rim1 = 1.0, ef = 1.0;
            for (int i = 0; i < cam.nD; i++, k++) {
                const double di = cam.val[k];
                const double ri = rim1 * r2;
                const double dRi = ri - cam.r0pow[k];
                const double dD = (di * dRi) * iN;
                const double dlx = xs * dD, dly = ys * dD;
                const double cR = (di * ef * rim1) * iN;
                chain(dlx, dly, xxs2 * cR + dD, xys2 * cR, xys2 * cR, yys2 * cR + dD);
                sink(k, (xs * dRi) * iN, (ys * dRi) * iN);
                gxs -= dlx * iN;
                gys -= dly * iN;
                rim1 = ri;
                ef += 1.0;
            }
        }
    } else {
        int k = 0;
        const int nc = cam.ncoef;
        while (k < nc) {
            const int t = cam.type[k];
            if (t == JAICOV_PT_AFFINITY_CX) {  // AffinityShearDistortionModelFactory.java:37-81
                const double cx = cam.val[k], cy = cam.val[k + 1];
                chain(cx * xs + cy * ys, 0.0, cx, cy, 0.0, 0.0);
                sink(k, xs, 0.0);
                sink(k + 1, ys, 0.0);
                k += 2;
            } else if (t == JAICOV_PT_TANGENTIAL_BX) {  // TangentialDistortionModelFactory.java:39-134
                const double bx = cam.val[k], by = cam.val[k + 1];
                double sum = 1.0;
                const double dlx = bx * (r2 + xxs2) + by * xys2;
                const double dly = by * (r2 + yys2) + bx * xys2;
                const double dXxs = 2.0 * (3.0 * bx * xs + by * ys);
                const double dXys = 2.0 * (by * xs + bx * ys);
                const double dYxs = dXys;
                const double dYys = 2.0 * (bx * xs + 3.0 * by * ys);
                chain(dlx, dly, dXxs, dXys, dYxs, dYys);
                int kb = k + 2;
                while (kb < nc && cam.type[kb] == JAICOV_PT_TANGENTIAL_B) {
                    const double bi = cam.val[kb];
                    const int e = cam.order[kb];
                    const double rim1 = r2pow(e - 1);
                    const double ri = rim1 * r2;
                    const double dT = bi * ri;
                    sum += dT;
                    const double cT = 2.0 * bi * e * rim1;
                    const double cTx = dlx * cT, cTy = dly * cT;
                    chain(dlx * dT, dly * dT, dT * dXxs + xs * cTx, dT * dXys + ys * cTx, dT * dYxs + xs * cTy,
                          dT * dYys + ys * cTy);
                    sink(kb, dlx * ri, dly * ri);
                    kb++;
                }
                sink(k, sum * (r2 + xxs2), sum * xys2);
                sink(k + 1, sum * xys2, sum * (r2 + yys2));
                k = kb;
            } else if (t == JAICOV_PT_RADIAL_A) {  // RadiallySymmetricDistortionModelFactory.java:39-90
                const double ai = cam.val[k];
                const int e = cam.order[k];
                const double rim1 = r2pow(e - 1);
                const double dRi = rim1 * r2 - cam.r0pow[k];
                const double dRad = ai * dRi;
                const double cR = ai * e * rim1;
                chain(xs * dRad, ys * dRad, xxs2 * cR + dRad, xys2 * cR, xys2 * cR, yys2 * cR + dRad);
                sink(k, xs * dRi, ys * dRi);
                k++;
            } else if (t == JAICOV_PT_DISTANCE_D) {  // RadialDistanceDistortionModelFactory.java:39-161
                const double di = cam.val[k];
                const int e = cam.order[k];
                const double rim1 = r2pow(e - 1);
                const double dRi = rim1 * r2 - cam.r0pow[k];
                const double dD = (di * dRi) * iN;
                const double dlx = xs * dD, dly = ys * dD;
                const double cR = (di * e * rim1) * iN;
                chain(dlx, dly, xxs2 * cR + dD, xys2 * cR, xys2 * cR, yys2 * cR + dD);
                sink(k, (xs * dRi) * iN, (ys * dRi) * iN);
                // explicit chain through N (:67-77, :105-159): coefficient -delta / N of dN/dp, applied after the loop
                gxs -= dlx * iN;
                gys -= dly * iN;
                k++;
            } else if (t == JAICOV_PT_ZERNIKE_Z) {  // ZernikeDistortionModelFactory.java:41-137 (gradient model)
                const double xxs = xs * xs, yys = ys * ys, xys = xs * ys;
                const double phi = atan2(ys, xs);
                const double rn2 = r2 / r02, c2 = 2.0 / rn2 / r02;
                const double zi = cam.val[k], m = (double)cam.zern_m[k];
                const double sm = sin(m * phi), cm = cos(m * phi);
                double pxZ = 0.0, pyZ = 0.0;
                for (int jt = cam.zern_ptr[k]; jt < cam.zern_ptr[k + 1]; jt++) {
                    const int pji = cam.zern_p[jt];
                    const double pj = (double)pji;
                    const int cei = pji / 2 - 1;
                    const double ce = (double)cei;
                    const double cC = cam.zern_c[jt] / r02 * ipow(rn2, cei);
                    double cX, cY, a, b, c_, d;
                    if (m < 0) {
                        cX = (-pj * xs * sm + m * ys * cm);
                        cY = (-pj * ys * sm - m * xs * cm);
                        a = zi * cC * (ce * xs * c2 * cX - pj * sm + m / r2 * (pj * xys * cm + m * yys * sm));
                        b = zi * cC * (ce * ys * c2 * cX + m * cm - m / r2 * (pj * xxs * cm + m * xys * sm));
                        c_ = zi * cC * (ce * xs * c2 * cY - m * cm + m / r2 * (pj * yys * cm - m * xys * sm));
                        d = zi * cC * (ce * ys * c2 * cY - pj * sm - m / r2 * (pj * xys * cm - m * xxs * sm));
                    } else {
                        cX = (pj * xs * cm + m * ys * sm);
                        cY = (pj * ys * cm - m * xs * sm);
                        a = zi * cC * (ce * xs * c2 * cX + pj * cm + m / r2 * (pj * xys * sm - m * yys * cm));
                        b = zi * cC * (ce * ys * c2 * cX + m * sm - m / r2 * (pj * xxs * sm - m * xys * cm));
                        c_ = zi * cC * (ce * xs * c2 * cY - m * sm + m / r2 * (pj * yys * sm + m * xys * cm));
                        d = zi * cC * (ce * ys * c2 * cY + pj * cm - m / r2 * (pj * xys * sm + m * xxs * cm));
                    }
                    chain(zi * cC * cX, zi * cC * cY, a, b, c_, d);
                    pxZ += cC * cX;
                    pyZ += cC * cY;
                }
                sink(k, pxZ, pyZ);
                k++;
            } else if (t == JAICOV_PT_ZERNIKE_X || t == JAICOV_PT_ZERNIKE_Y) {  // :147-227 scalar models, verbatim
                const double phi = atan2(ys, xs);
                const double rn2 = r2 / r02;
                const double zi = cam.val[k], m = (double)cam.zern_m[k];
                const double sm = sin(m * phi), cm = cos(m * phi);
                double pZ = 0.0;
                for (int jt = cam.zern_ptr[k]; jt < cam.zern_ptr[k + 1]; jt++) {
                    const int pji = cam.zern_p[jt];
                    const double pj = (double)pji;
                    const double cj = cam.zern_c[jt];
                    const double rp = ipow(rn2, pji / 2 - 1);
                    const double cC = cj * ipow(rn2, pji / 2);
                    const double cZ = zi * cj / r02 * rp;
                    double delta, pdx, pdy;
                    if (m < 0) {
                        pdx = cZ * (-pj * xs * sm + m * ys * cm);
                        pdy = cZ * (-pj * ys * sm - m * xs * cm);
                        delta = -zi * cC * sm;
                        pZ += -cC * sm;
                    } else {
                        pdx = cZ * (pj * xs * cm + m * ys * sm);
                        pdy = cZ * (pj * ys * cm - m * xs * sm);
                        delta = zi * cC * cm;
                        pZ += cC * cm;
                    }
                    if (t == JAICOV_PT_ZERNIKE_X) chain(delta, 0.0, pdx, pdy, 0.0, 0.0);
                    else chain(0.0, delta, 0.0, 0.0, pdx, pdy);
                }
                if (t == JAICOV_PT_ZERNIKE_X) sink(k, pZ, 0.0);
                else sink(k, 0.0, pZ);
                k++;
            } else {
                k++;
            }
        }
    }
    // apply the accumulated factors once; dN/d(X,Y,Z) = (r13, r23, r33), dN/dc = 0, dN/d omega = -r33 dY + r23 dZ,
    // dN/d phi = kx cos(kappa) - ky sin(kappa), dN/d kappa = 0   (RadialDistanceDistortionModelFactory.java:67-77)
    const double pN0 = q.r13, pN1 = q.r23, pN2 = q.r33;
    r.ax[0] = Dxx * px[0] + Dxy * py[0] + pN0 * gxs;  r.ay[0] = Dyx * px[0] + Dyy * py[0] + pN0 * gys;
    r.ax[1] = Dxx * px[1] + Dxy * py[1] + pN1 * gxs;  r.ay[1] = Dyx * px[1] + Dyy * py[1] + pN1 * gys;
    r.ax[2] = Dxx * px[2] + Dxy * py[2] + pN2 * gxs;  r.ay[2] = Dyx * px[2] + Dyy * py[2] + pN2 * gys;
    r.ax[3] = Dxx * px[3] + Dxy * py[3];              r.ay[3] = Dyx * px[3] + Dyy * py[3];
    if (WITH_EO) {
        const double pN4 = -q.r33 * dY + q.r23 * dZ, pN5 = kx * q.cosK - ky * q.sinK;
        r.ax[4] = Dxx * px[NB - 3] + Dxy * py[NB - 3] + pN4 * gxs;  r.ay[4] = Dyx * px[NB - 3] + Dyy * py[NB - 3] + pN4 * gys;
        r.ax[5] = Dxx * px[NB - 2] + Dxy * py[NB - 2] + pN5 * gxs;  r.ay[5] = Dyx * px[NB - 2] + Dyy * py[NB - 2] + pN5 * gys;
        r.ax[6] = Dxx * px[NB - 1] + Dxy * py[NB - 1];              r.ay[6] = Dyx * px[NB - 1] + Dyy * py[NB - 1];
    } else {
        r.ax[4] = r.ax[5] = r.ax[6] = 0.0;
        r.ay[4] = r.ay[5] = r.ay[6] = 0.0;
    }
    r.w0 = w0;
    r.w1 = w1;
}

template <class Sink>
__host__ __device__ __forceinline__ void eval_observation(const ImgPose &q, const CamView &cam, double X, double Y, double Z,
                                                          double xobs, double yobs, BaseRows &r, Sink &&sink) {
    eval_observation_t<true, false>(q, cam, X, Y, Z, xobs, yobs, r, sink);
}

// Weight matrix of an image point, PartialDerivativeFactory.java:296-319: returns P00, P01, P11
__host__ __device__ __forceinline__ void point_weight(double sigma2, double varX, double varY, double rho, double &p00,
                                             double &p01, double &p11) {
    if (rho == 0.0) {
        p00 = sigma2 / varX; p11 = sigma2 / varY; p01 = 0.0;
    } else {
        const double invDet = sigma2 / ((1.0 - rho * rho) * varX * varY);
        p00 = invDet * varY; p11 = invDet * varX; p01 = -invDet * rho * sqrt(varX * varY);
    }
}

}  // namespace jaicov
