/*
 * jaicov_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the floating-point part of JAICOV's adjustment hot path,
 * written to follow the reference's evaluation and accumulation ORDER so that it can
 * serve as the parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's library.
 * Pinned (a) at entry level: every Jacobian entry and misclosure of eval_point equals -- bit for bit -- the value the
 * reference's own formula bodies give when executed (tests/golden/make_jacobian_fixture.py ->
 * tests/golden/reference_jacobian.npz, tests/test_reference_formulas.py), for all distortion models incl. Zernike; and
 * (b) end to end against the known answers of the reference's bundled example (tests/test_oracle_golden.py).
 * Compile with -ffp-contract=off (no FMA contraction: the JVM does not fuse).
 *
 * Reference files restated here (paths relative to
 * /root/reference/JAICOV/src/org/applied_geodesy/adjustment/):
 *   PDF  = bundle/derivation/PartialDerivativeFactory.java
 *   DMF  = bundle/derivation/DistortionModelFactory.java
 *   RAD  = bundle/derivation/RadiallySymmetricDistortionModelFactory.java
 *   TAN  = bundle/derivation/TangentialDistortionModelFactory.java
 *   AFF  = bundle/derivation/AffinityShearDistortionModelFactory.java
 *   DIST = bundle/derivation/RadialDistanceDistortionModelFactory.java
 *   ZER  = bundle/derivation/ZernikeDistortionModelFactory.java
 *   ZC   = bundle/parameter/ZernikeCoefficient.java
 *   BA   = bundle/BundleAdjustment.java
 *   NES  = NormalEquationSystem.java
 *   ME   = MathExtension.java
 *
 * The dense factor/solve/invert (LAPACK dspsv/dsptri/dpptrf/dpptri, ME:304-366) lives in
 * binary jars in the reference (mtj-1.0.4, netlib-java core-1.1.2, f2j arpack_combined_all-0.1);
 * the oracle calls the same reference-LAPACK routines out of scipy's OpenBLAS (oracle/oracle.py).
 *
 * Data layout: the "flat problem" of include/jaicov_b200.h (values + column indices after
 * the integer bookkeeping of BA:667-782, which oracle/bookkeeping.py restates).
 * Matrix N is MTJ UpperSymmPackMatrix layout: column-major packed upper, idx = r + c(c+1)/2, r<=c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define COL_FIXED 2147483647 /* Integer.MAX_VALUE, parameter/UnknownParameter.java:27 */
#define ORC_MAXSLOT 256

/* ParameterType ids, parameter/ParameterType.java:27-56 */
enum {
    PT_RADIAL_A = 121, PT_TANG_B = 131, PT_TANG_BX = 132, PT_TANG_BY = 133,
    PT_AFF_CX = 141, PT_AFF_CY = 142, PT_DIST_D = 151,
    PT_ZERN_X = 161, PT_ZERN_Y = 162, PT_ZERN_Z = 163
};

typedef struct {
    /* cameras */
    int32_t nCam;
    const double *io_val;      /* [3*nCam] x0, y0, c  (iterator order InteriorOrientation.java:70-79) */
    const int32_t *io_col;     /* [3*nCam] */
    const double *r0;          /* [nCam] */
    const int32_t *coef_ptr;   /* [nCam+1] */
    const int32_t *coef_type;  /* ParameterType id, evaluation order (camera/Camera.java:50) */
    const int32_t *coef_order;
    const double *coef_val;
    const int32_t *coef_col;
    /* images */
    int32_t nImg;
    const int32_t *cam_of_img;
    const double *eo_val;      /* [6*nImg] X0,Y0,Z0,omega,phi,kappa */
    const int32_t *eo_col;
    const int64_t *pt_ptr;     /* [nImg+1] CSR over image points */
    /* image points */
    int64_t m;
    const int32_t *obj_idx;
    const double *xy;          /* [2m] */
    const double *var;         /* [2m] sigma_x^2, sigma_y^2 */
    const double *rho;         /* [m] */
    /* object points */
    int32_t nPt;
    const double *xyz;         /* [3*nPt] */
    const int32_t *pt_col;     /* [3*nPt] */
    /* scale bars */
    int32_t nBar;
    const int32_t *bar_a, *bar_b;
    const double *bar_len, *bar_var;
} orc_problem;

static inline int active(int32_t c) { return c >= 0 && c != COL_FIXED; }

/* MathExtension.binomial, ME:53-64 */
static long long orc_binomial(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    long long result = 1;
    for (int i = 1; i <= k; i++) result = result * (n - k + i) / i;
    return result;
}

/* Collinearity equations, PDF:94-190 */
typedef struct {
    double cosOmega, sinOmega, cosPhi, sinPhi, cosKappa, sinKappa;
    double r11, r12, r13, r21, r22, r23, r31, r32, r33;
    double xs, ys, x, y, dX, dY, dZ, kx, ky, N, kxN, kyN;
    double pxs[12], pys[12]; /* slot order: X,Y,Z,x0,y0,c,X0,Y0,Z0,omega,phi,kappa */
} coll_t;

static void collinearity(coll_t *q, const double *io, const double *eo, const double *P) {
    double x0 = io[0], y0 = io[1], c = io[2];
    double X0 = eo[0], Y0 = eo[1], Z0 = eo[2], omega = eo[3], phi = eo[4], kappa = eo[5];
    double X = P[0], Y = P[1], Z = P[2];
    q->cosOmega = cos(omega); q->sinOmega = sin(omega);
    q->cosPhi = cos(phi);     q->sinPhi = sin(phi);
    q->cosKappa = cos(kappa); q->sinKappa = sin(kappa);
    q->r11 = q->cosPhi * q->cosKappa;
    q->r12 = -q->cosPhi * q->sinKappa;
    q->r13 = q->sinPhi;
    q->r21 = q->cosOmega * q->sinKappa + q->sinOmega * q->sinPhi * q->cosKappa;
    q->r22 = q->cosOmega * q->cosKappa - q->sinOmega * q->sinPhi * q->sinKappa;
    q->r23 = -q->sinOmega * q->cosPhi;
    q->r31 = q->sinOmega * q->sinKappa - q->cosOmega * q->sinPhi * q->cosKappa;
    q->r32 = q->sinOmega * q->cosKappa + q->cosOmega * q->sinPhi * q->sinKappa;
    q->r33 = q->cosOmega * q->cosPhi;
    q->dX = X - X0; q->dY = Y - Y0; q->dZ = Z - Z0;
    q->kx = q->r11 * q->dX + q->r21 * q->dY + q->r31 * q->dZ;
    q->ky = q->r12 * q->dX + q->r22 * q->dY + q->r32 * q->dZ;
    q->N  = q->r13 * q->dX + q->r23 * q->dY + q->r33 * q->dZ;
    q->kxN = q->kx / q->N; q->kyN = q->ky / q->N;
    q->xs = -c * q->kxN; q->ys = -c * q->kyN;
    q->x = x0 + q->xs; q->y = y0 + q->ys;
    /* PDF:157-171 */
    q->pxs[0] = -(q->r13 * q->xs + c * q->r11) / q->N;
    q->pxs[1] = -(q->r23 * q->xs + c * q->r21) / q->N;
    q->pxs[2] = -(q->r33 * q->xs + c * q->r31) / q->N;
    q->pxs[3] = 1.0; q->pxs[4] = 0.0; q->pxs[5] = -q->kxN;
    q->pxs[6] = -q->pxs[0]; q->pxs[7] = -q->pxs[1]; q->pxs[8] = -q->pxs[2];
    q->pxs[9]  = (q->xs * (q->r33 * q->dY - q->r23 * q->dZ) + c * (q->r31 * q->dY - q->r21 * q->dZ)) / q->N;
    q->pxs[10] = (q->xs * (q->ky * q->sinKappa - q->kx * q->cosKappa) + c * q->N * q->cosKappa) / q->N;
    q->pxs[11] = q->ys;
    /* PDF:175-189 */
    q->pys[0] = -(q->r13 * q->ys + c * q->r12) / q->N;
    q->pys[1] = -(q->r23 * q->ys + c * q->r22) / q->N;
    q->pys[2] = -(q->r33 * q->ys + c * q->r32) / q->N;
    q->pys[3] = 0.0; q->pys[4] = 1.0; q->pys[5] = -q->kyN;
    q->pys[6] = -q->pys[0]; q->pys[7] = -q->pys[1]; q->pys[8] = -q->pys[2];
    q->pys[9]  = (q->ys * (q->r33 * q->dY - q->r23 * q->dZ) + c * (q->r32 * q->dY - q->r22 * q->dZ)) / q->N;
    q->pys[10] = (q->ys * (q->ky * q->sinKappa - q->kx * q->cosKappa) - c * q->N * q->sinKappa) / q->N;
    q->pys[11] = -q->xs;
}

/* Compact view of one 2 x n Jacobian row pair: slot s has column col[s] and entries a0[s], a1[s].
 * Slots 0..11 = X,Y,Z,x0,y0,c,X0,Y0,Z0,omega,phi,kappa; slots 12.. = the camera's coefficients. */
typedef struct {
    int ns;
    int32_t col[ORC_MAXSLOT];
    double a0[ORC_MAXSLOT], a1[ORC_MAXSLOT];
    double w[2];
} rows_t;

/* DistortionModelFactory.apply, DMF:33-101: chain rule onto X,Y,Z,c,EO (never x0,y0); base partials */
static void dmf_apply(const coll_t *q, rows_t *r, double deltaX, double deltaY,
                      double dXxs, double dXys, double dYxs, double dYys) {
    static const int slots[10] = {0, 1, 2, 5, 6, 7, 8, 9, 10, 11};
    r->w[0] += -deltaX;
    r->w[1] += -deltaY;
    for (int i = 0; i < 10; i++) {
        int s = slots[i];
        if (active(r->col[s])) {
            r->a0[s] += dXxs * q->pxs[s] + dXys * q->pys[s];
            r->a1[s] += dYxs * q->pxs[s] + dYys * q->pys[s];
        }
    }
}

static void set_own(rows_t *r, int s, double v0, double v1) {
    if (active(r->col[s])) { r->a0[s] = v0; r->a1[s] = v1; }
}

/* ZernikeCoefficient.ZernikePolynomial, ZC:40-56 */
typedef struct { int n, m, nterms; long long p[64], c[64]; double length; } zern_t;
static void zernike_poly(zern_t *z, int order) {
    z->n = (int)ceil((-3 + sqrt(9 + 8 * order)) / 2);
    z->m = 2 * order - z->n * (z->n + 2);
    int halfnm = (z->n - abs(z->m)) / 2;
    z->nterms = halfnm + 1;
    for (int k = 0; k <= halfnm; k++) {
        z->p[k] = z->n - 2 * k;
        z->c[k] = ((k % 2 == 0) ? 1 : -1) * orc_binomial(z->n - k, k) * orc_binomial(z->n - 2 * k, halfnm - k);
    }
    z->length = sqrt((1 + ((z->m != 0) ? 1 : 0)) * (z->n + 1) / M_PI);
}

/* Evaluate one image point: PDF:285-442. Fills rows (unsorted slots), P (p00,p01,p11) and the
 * diagonalWeighting flag (PDF:300). */
static void eval_image_point(const orc_problem *pb, int32_t img, int64_t j, double sigma2apriori,
                             rows_t *r, coll_t *q, double P[3], int *diag) {
    int32_t cam = pb->cam_of_img[img];
    int32_t p = pb->obj_idx[j];
    const double *io = pb->io_val + 3 * cam;
    const double *eo = pb->eo_val + 6 * (int64_t)img;
    collinearity(q, io, eo, pb->xyz + 3 * (int64_t)p);

    double varianceX = pb->var[2 * j], varianceY = pb->var[2 * j + 1], corr = pb->rho[j];
    *diag = (corr == 0);
    if (*diag) {
        P[0] = sigma2apriori / varianceX; P[2] = sigma2apriori / varianceY; P[1] = 0.0;
    } else {
        double invDet = sigma2apriori / ((1.0 - corr * corr) * varianceX * varianceY);
        P[0] = invDet * varianceY; P[2] = invDet * varianceX;
        P[1] = -invDet * corr * sqrt(varianceX * varianceY);
    }
    r->w[0] = pb->xy[2 * j] - q->x;
    r->w[1] = pb->xy[2 * j + 1] - q->y;

    int c0 = pb->coef_ptr[cam], c1 = pb->coef_ptr[cam + 1];
    r->ns = 12 + (c1 - c0);
    for (int s = 0; s < 3; s++) r->col[s] = pb->pt_col[3 * (int64_t)p + s];
    for (int s = 0; s < 3; s++) r->col[3 + s] = pb->io_col[3 * cam + s];
    for (int s = 0; s < 6; s++) r->col[6 + s] = pb->eo_col[6 * (int64_t)img + s];
    for (int k = c0; k < c1; k++) r->col[12 + k - c0] = pb->coef_col[k];
    for (int s = 0; s < r->ns; s++) { r->a0[s] = 0.0; r->a1[s] = 0.0; }
    for (int s = 0; s < 12; s++)
        if (active(r->col[s])) { r->a0[s] = q->pxs[s]; r->a1[s] = q->pys[s]; }

    const double xs = q->xs, ys = q->ys;
    double r0 = pb->r0[cam];
    int k = c0;
    while (k < c1) {
        int t = pb->coef_type[k];
        if (t == PT_AFF_CX) { /* AFF:37-81; list holds Cx, Cy consecutively */
            double cx = pb->coef_val[k], cy = pb->coef_val[k + 1];
            double deltaX = cx * xs + cy * ys, deltaY = 0.0;
            dmf_apply(q, r, deltaX, deltaY, cx, cy, 0.0, 0.0);
            set_own(r, 12 + k - c0, xs, 0.0);
            set_own(r, 12 + k + 1 - c0, ys, 0.0);
            k += 2;
        } else if (t == PT_TANG_BX) { /* TAN:39-134; list holds Bx, By, then Bi... */
            double bx = pb->coef_val[k], by = pb->coef_val[k + 1];
            double r2 = xs * xs + ys * ys;
            double xxs2 = 2.0 * xs * xs, yys2 = 2.0 * ys * ys, xys2 = 2.0 * xs * ys;
            double sum = 1.0;
            double deltaX = bx * (r2 + xxs2) + by * xys2;
            double deltaY = by * (r2 + yys2) + bx * xys2;
            double dXxs = 2.0 * (3.0 * bx * xs + by * ys);
            double dXys = 2.0 * (by * xs + bx * ys);
            double dYxs = 2.0 * (by * xs + bx * ys);
            double dYys = 2.0 * (bx * xs + 3.0 * by * ys);
            dmf_apply(q, r, deltaX, deltaY, dXxs, dXys, dYxs, dYys);
            int kb = k + 2;
            while (kb < c1 && pb->coef_type[kb] == PT_TANG_B) {
                double bi = pb->coef_val[kb];
                int expi = pb->coef_order[kb];
                double ri = pow(r2, expi);
                double dTani = bi * ri;
                sum += dTani;
                double deltaXi = deltaX * dTani, deltaYi = deltaY * dTani;
                double par_xs_Bi = deltaX * ri, par_ys_Bi = deltaY * ri;
                double constTani = 2.0 * bi * expi * pow(r2, expi - 1);
                double constTanXi = deltaX * constTani, constTanYi = deltaY * constTani;
                double dXxsi = dTani * dXxs + xs * constTanXi;
                double dXysi = dTani * dXys + ys * constTanXi;
                double dYxsi = dTani * dYxs + xs * constTanYi;
                double dYysi = dTani * dYys + ys * constTanYi;
                dmf_apply(q, r, deltaXi, deltaYi, dXxsi, dXysi, dYxsi, dYysi);
                set_own(r, 12 + kb - c0, par_xs_Bi, par_ys_Bi);
                kb++;
            }
            set_own(r, 12 + k - c0, sum * (r2 + xxs2), sum * xys2);
            set_own(r, 12 + k + 1 - c0, sum * xys2, sum * (r2 + yys2));
            k = kb;
        } else if (t == PT_RADIAL_A) { /* RAD:39-90 */
            double r2 = xs * xs + ys * ys, r02 = r0 * r0;
            double xxs2 = 2.0 * xs * xs, yys2 = 2.0 * ys * ys, xys2 = 2.0 * xs * ys;
            while (k < c1 && pb->coef_type[k] == PT_RADIAL_A) {
                double ai = pb->coef_val[k];
                int expi = pb->coef_order[k];
                double dRi = pow(r2, expi) - pow(r02, expi);
                double dRadi = ai * dRi;
                double deltaX = xs * dRadi, deltaY = ys * dRadi;
                double constRadi = ai * expi * pow(r2, expi - 1);
                double dXxs = xxs2 * constRadi + dRadi, dXys = xys2 * constRadi;
                double dYxs = xys2 * constRadi, dYys = yys2 * constRadi + dRadi;
                dmf_apply(q, r, deltaX, deltaY, dXxs, dXys, dYxs, dYys);
                set_own(r, 12 + k - c0, xs * dRi, ys * dRi);
                k++;
            }
        } else if (t == PT_DIST_D) { /* DIST:39-161 */
            double r2 = xs * xs + ys * ys, r02 = r0 * r0;
            double xxs2 = 2.0 * xs * xs, yys2 = 2.0 * ys * ys, xys2 = 2.0 * xs * ys;
            while (k < c1 && pb->coef_type[k] == PT_DIST_D) {
                double di = pb->coef_val[k];
                int expi = pb->coef_order[k];
                double dRi = pow(r2, expi) - pow(r02, expi);
                double dDisti = (di * dRi) / q->N;
                double deltaX = xs * dDisti, deltaY = ys * dDisti;
                double pN[12] = {q->r13, q->r23, q->r33, 0, 0, 0, -q->r13, -q->r23, -q->r33,
                                 -q->r33 * q->dY + q->r23 * q->dZ,
                                 q->kx * q->cosKappa - q->ky * q->sinKappa, 0.0};
                double constRadi = (di * expi * pow(r2, expi - 1)) / q->N;
                double dXxs = xxs2 * constRadi + dDisti, dXys = xys2 * constRadi;
                double par_dDistX_N = -deltaX / q->N;
                double dYxs = xys2 * constRadi, dYys = yys2 * constRadi + dDisti;
                double par_dDistY_N = -deltaY / q->N;
                dmf_apply(q, r, deltaX, deltaY, dXxs, dXys, dYxs, dYys);
                set_own(r, 12 + k - c0, (xs * dRi) / q->N, (ys * dRi) / q->N);
                static const int slots[9] = {0, 1, 2, 6, 7, 8, 9, 10, 11};
                for (int i = 0; i < 9; i++) {
                    int s = slots[i];
                    if (active(r->col[s])) {
                        r->a0[s] += pN[s] * par_dDistX_N;
                        r->a1[s] += pN[s] * par_dDistY_N;
                    }
                }
                k++;
            }
        } else if (t == PT_ZERN_Z) { /* ZER:41-137 gradient model */
            double r02 = r0 * r0;
            double xxs = xs * xs, yys = ys * ys, xys = xs * ys;
            double phi = atan2(ys, xs);
            double r2 = xxs + yys, rn2 = r2 / r02, const2rnr0 = 2.0 / rn2 / r02;
            while (k < c1 && pb->coef_type[k] == PT_ZERN_Z) {
                zern_t z; zernike_poly(&z, pb->coef_order[k]);
                double zi = pb->coef_val[k], m = z.m;
                double sinmphi = sin(m * phi), cosmphi = cos(m * phi);
                double par_xs_Zi = 0, par_ys_Zi = 0;
                for (int jt = 0; jt < z.nterms; jt++) {
                    long long pj = z.p[jt];
                    long long constExp = (pj / 2 - 1);
                    double cj = z.length * z.c[jt];
                    double constC = cj / r02 * pow(rn2, (double)constExp);
                    if (m < 0) {
                        double cX = (-pj * xs * sinmphi + m * ys * cosmphi);
                        double cY = (-pj * ys * sinmphi - m * xs * cosmphi);
                        double dX_ = zi * constC * cX, dY_ = zi * constC * cY;
                        double a = zi * constC * (constExp * xs * const2rnr0 * cX - pj * sinmphi + m / r2 * (pj * xys * cosmphi + m * yys * sinmphi));
                        double b = zi * constC * (constExp * ys * const2rnr0 * cX + m * cosmphi - m / r2 * (pj * xxs * cosmphi + m * xys * sinmphi));
                        double c_ = zi * constC * (constExp * xs * const2rnr0 * cY - m * cosmphi + m / r2 * (pj * yys * cosmphi - m * xys * sinmphi));
                        double d = zi * constC * (constExp * ys * const2rnr0 * cY - pj * sinmphi - m / r2 * (pj * xys * cosmphi - m * xxs * sinmphi));
                        dmf_apply(q, r, dX_, dY_, a, b, c_, d);
                        par_xs_Zi += constC * cX; par_ys_Zi += constC * cY;
                    } else {
                        double cX = (pj * xs * cosmphi + m * ys * sinmphi);
                        double cY = (pj * ys * cosmphi - m * xs * sinmphi);
                        double dX_ = zi * constC * cX, dY_ = zi * constC * cY;
                        double a = zi * constC * (constExp * xs * const2rnr0 * cX + pj * cosmphi + m / r2 * (pj * xys * sinmphi - m * yys * cosmphi));
                        double b = zi * constC * (constExp * ys * const2rnr0 * cX + m * sinmphi - m / r2 * (pj * xxs * sinmphi - m * xys * cosmphi));
                        double c_ = zi * constC * (constExp * xs * const2rnr0 * cY - m * sinmphi + m / r2 * (pj * yys * sinmphi + m * xys * cosmphi));
                        double d = zi * constC * (constExp * ys * const2rnr0 * cY + pj * cosmphi - m / r2 * (pj * xys * sinmphi + m * xxs * cosmphi));
                        dmf_apply(q, r, dX_, dY_, a, b, c_, d);
                        par_xs_Zi += constC * cX; par_ys_Zi += constC * cY;
                    }
                }
                set_own(r, 12 + k - c0, par_xs_Zi, par_ys_Zi);
                k++;
            }
        } else if (t == PT_ZERN_X || t == PT_ZERN_Y) { /* ZER:147-227 scalar models (integer pj/2, verbatim) */
            double r02 = r0 * r0;
            double xxs = xs * xs, yys = ys * ys;
            double phi = atan2(ys, xs);
            double r2 = xxs + yys, rn2 = r2 / r02;
            int type = t;
            while (k < c1 && pb->coef_type[k] == type) {
                zern_t z; zernike_poly(&z, pb->coef_order[k]);
                double zi = pb->coef_val[k], m = z.m;
                double sinmphi = sin(m * phi), cosmphi = cos(m * phi);
                double par_Zi = 0;
                for (int jt = 0; jt < z.nterms; jt++) {
                    long long pj = z.p[jt];
                    double cj = z.length * z.c[jt];
                    double constC = cj * pow(rn2, (double)(pj / 2));
                    double constZ = zi * cj / r02 * pow(rn2, (double)(pj / 2 - 1));
                    double delta = 0, pdxs = 0, pdys = 0;
                    if (m < 0) {
                        double cX = (-pj * xs * sinmphi + m * ys * cosmphi);
                        double cY = (-pj * ys * sinmphi - m * xs * cosmphi);
                        pdxs = constZ * cX; pdys = constZ * cY;
                        delta = -zi * constC * sinmphi;
                        par_Zi += -constC * sinmphi;
                    } else {
                        double cX = (pj * xs * cosmphi + m * ys * sinmphi);
                        double cY = (pj * ys * cosmphi - m * xs * sinmphi);
                        pdxs = constZ * cX; pdys = constZ * cY;
                        delta = +zi * constC * cosmphi;
                        par_Zi += +constC * cosmphi;
                    }
                    if (type == PT_ZERN_X) dmf_apply(q, r, delta, 0, pdxs, pdys, 0, 0);
                    else dmf_apply(q, r, 0, delta, 0, 0, pdxs, pdys);
                }
                int s = 12 + k - c0;
                if (active(r->col[s])) {
                    if (type == PT_ZERN_X) r->a0[s] = par_Zi; else r->a1[s] = par_Zi;
                }
                k++;
            }
        } else {
            k++; /* unknown type id: ignored */
        }
    }
}

/* sort the active slots by column (Collections.sort(columns), PDF:476) */
static int sorted_active(const rows_t *r, int *order) {
    int n = 0;
    for (int s = 0; s < r->ns; s++) if (active(r->col[s])) order[n++] = s;
    for (int i = 1; i < n; i++) { /* insertion sort */
        int s = order[i], jx = i - 1;
        while (jx >= 0 && r->col[order[jx]] > r->col[s]) { order[jx + 1] = order[jx]; jx--; }
        order[jx + 1] = s;
    }
    return n;
}

static inline int64_t pidx(int64_t r, int64_t c) { return r + c * (c + 1) / 2; } /* r <= c */

/* stackNormalEquationSystem for a 2-row group, PDF:475-505 (NEQ may be NULL) */
static void stack2(double *NEQ, double *neq, const rows_t *r, const double P3[3], int diag) {
    int order[ORC_MAXSLOT];
    int nc = sorted_active(r, order);
    const double P[2][2] = {{P3[0], P3[1]}, {P3[1], P3[2]}};
    for (int row = 0; row < 2; row++) {
        const double *Ar = row == 0 ? r->a0 : r->a1;
        for (int ia = 0; ia < nc; ia++) {
            int sa = order[ia];
            int64_t colAT = r->col[sa];
            double aT = Ar[sa];
            if (diag) neq[colAT] += aT * P[row][row] * r->w[row];
            else for (int colP = 0; colP < 2; colP++) neq[colAT] += aT * P[row][colP] * r->w[colP];
            if (NEQ) {
                for (int ib = ia; ib < nc; ib++) {
                    int sb = order[ib];
                    int64_t colA = r->col[sb];
                    if (diag) NEQ[pidx(colAT, colA)] += aT * P[row][row] * Ar[sb];
                    else {
                        NEQ[pidx(colAT, colA)] += aT * P[row][0] * r->a0[sb];
                        NEQ[pidx(colAT, colA)] += aT * P[row][1] * r->a1[sb];
                    }
                }
            }
        }
    }
}

/* ---- exported API ------------------------------------------------------------------------- */

/* Per-entry view of one image point for K1 parity: slot-ordered (12 + ncoef) columns/entries. */
int orc_eval_point(const orc_problem *pb, int32_t img, int64_t j, double sigma2apriori,
                   int32_t *cols, double *a0, double *a1, double *w, double *P3) {
    rows_t r; coll_t q; int diag;
    eval_image_point(pb, img, j, sigma2apriori, &r, &q, P3, &diag);
    for (int s = 0; s < r.ns; s++) { cols[s] = r.col[s]; a0[s] = r.a0[s]; a1[s] = r.a1[s]; }
    w[0] = r.w[0]; w[1] = r.w[1];
    return r.ns;
}

/* Image-point and scale-bar part of createNormalEquation, BA:795-797, in observation-group order
 * (image points camera->image->point, then scale bars; directly observed groups are added by the
 * Python driver, which owns their weight matrices). N/n must be zeroed by the caller. */
void orc_stack_image_points(const orc_problem *pb, double sigma2apriori, double *NEQ, double *neq,
                            int64_t j_begin, int64_t j_end) {
    rows_t r; coll_t q; double P3[3]; int diag;
    for (int32_t img = 0; img < pb->nImg; img++) {
        for (int64_t j = pb->pt_ptr[img]; j < pb->pt_ptr[img + 1]; j++) {
            if (j < j_begin || j >= j_end) continue;
            eval_image_point(pb, img, j, sigma2apriori, &r, &q, P3, &diag);
            stack2(NEQ, neq, &r, P3, diag);
        }
    }
}

/* getPartialDerivativeScaleBar, PDF:210-283 */
static void eval_bar(const orc_problem *pb, int b, double sigma2apriori, int32_t cols[6], double a[6], double *w, double *P) {
    const double *A = pb->xyz + 3 * (int64_t)pb->bar_a[b], *B = pb->xyz + 3 * (int64_t)pb->bar_b[b];
    double dX = B[0] - A[0], dY = B[1] - A[1], dZ = B[2] - A[2];
    double lengthAB = sqrt(dX * dX + dY * dY + dZ * dZ);
    double ax = dX / lengthAB, ay = dY / lengthAB, az = dZ / lengthAB;
    *P = sigma2apriori / pb->bar_var[b];
    *w = pb->bar_len[b] - lengthAB;
    for (int s = 0; s < 3; s++) { cols[s] = pb->pt_col[3 * (int64_t)pb->bar_a[b] + s]; cols[3 + s] = pb->pt_col[3 * (int64_t)pb->bar_b[b] + s]; }
    a[0] = -ax; a[1] = -ay; a[2] = -az; a[3] = ax; a[4] = ay; a[5] = az;
}

void orc_stack_scale_bars(const orc_problem *pb, double sigma2apriori, double *NEQ, double *neq) {
    for (int b = 0; b < pb->nBar; b++) {
        int32_t cols[6]; double a[6], w, P;
        eval_bar(pb, b, sigma2apriori, cols, a, &w, &P);
        int order[6], nc = 0;
        for (int s = 0; s < 6; s++) if (active(cols[s])) order[nc++] = s;
        for (int i = 1; i < nc; i++) { int s = order[i], jx = i - 1; while (jx >= 0 && cols[order[jx]] > cols[s]) { order[jx + 1] = order[jx]; jx--; } order[jx + 1] = s; }
        for (int ia = 0; ia < nc; ia++) {
            double aT = a[order[ia]];
            int64_t colAT = cols[order[ia]];
            neq[colAT] += aT * P * w;
            if (NEQ) for (int ib = ia; ib < nc; ib++) NEQ[pidx(colAT, cols[order[ib]])] += aT * P * a[order[ib]];
        }
    }
}

/* getOmega, BA:472-491, image points + scale bars: v = w - A*dx (dgemv 'N' order: ascending column),
 * Pv = P*v (dspmv/dsbmv order), omega += v.Pv */
double orc_omega_image_points(const orc_problem *pb, double sigma2apriori, const double *dx) {
    rows_t r; coll_t q; double P3[3]; int diag; double omega = 0.0;
    int order[ORC_MAXSLOT];
    for (int32_t img = 0; img < pb->nImg; img++) {
        for (int64_t j = pb->pt_ptr[img]; j < pb->pt_ptr[img + 1]; j++) {
            eval_image_point(pb, img, j, sigma2apriori, &r, &q, P3, &diag);
            int nc = sorted_active(&r, order);
            double v0 = r.w[0], v1 = r.w[1];
            for (int i = 0; i < nc; i++) {
                int s = order[i];
                double temp = -1.0 * dx[r.col[s]];
                v0 += temp * r.a0[s];
                v1 += temp * r.a1[s];
            }
            double Pv0, Pv1;
            if (diag) { Pv0 = v0 * P3[0]; Pv1 = v1 * P3[2]; }
            else { Pv0 = v0 * P3[0]; Pv0 += v1 * P3[1]; Pv1 = v1 * P3[2] + P3[1] * v0; }
            omega += v0 * Pv0 + v1 * Pv1;
        }
    }
    return omega;
}

double orc_omega_scale_bars(const orc_problem *pb, double sigma2apriori, const double *dx) {
    double omega = 0.0;
    for (int b = 0; b < pb->nBar; b++) {
        int32_t cols[6]; double a[6], w, P;
        eval_bar(pb, b, sigma2apriori, cols, a, &w, &P);
        int order[6], nc = 0;
        for (int s = 0; s < 6; s++) if (active(cols[s])) order[nc++] = s;
        for (int i = 1; i < nc; i++) { int s = order[i], jx = i - 1; while (jx >= 0 && cols[order[jx]] > cols[s]) { order[jx + 1] = order[jx]; jx--; } order[jx + 1] = s; }
        double v = w;
        for (int i = 0; i < nc; i++) v += (-1.0 * dx[cols[order[i]]]) * a[order[i]];
        omega += v * (v * P);
    }
    return omega;
}

/* applyPrecondition, NES:82-91 on packed upper storage: M[r,c] = V[c]*M[r,c]*V[r]; m[r] = V[r]*m[r] */
void orc_apply_precondition(int64_t n, const double *V, double *M, double *m) {
    for (int64_t row = 0; row < n; row++) {
        if (m) m[row] = V[row] * m[row];
        if (M) for (int64_t col = row; col < n; col++) M[pidx(row, col)] = V[col] * M[pidx(row, col)] * V[row];
    }
}

/* Preconditioner, BA:824-828 */
void orc_preconditioner(int64_t n, const double *N, double *V, double eps) {
    for (int64_t c = 0; c < n; c++) { double v = N[pidx(c, c)]; V[c] = v > eps ? 1.0 / sqrt(v) : 1.0; }
}

/* packed upper -> dense (both triangles), helper for tests and the fast route */
void orc_unpack(int64_t n, const double *Np, double *D) {
    for (int64_t c = 0; c < n; c++) for (int64_t r = 0; r <= c; r++) { double v = Np[pidx(r, c)]; D[r * n + c] = v; D[c * n + r] = v; }
}

/* dense symmetric (upper triangle read) -> packed upper, helper of the blocked route (oracle/fast_oracle.py) */
void orc_pack(int64_t n, const double *D, double *Np) {
    for (int64_t c = 0; c < n; c++) for (int64_t r = 0; r <= c; r++) Np[pidx(r, c)] = D[r * n + c];
}
