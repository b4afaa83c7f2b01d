/*
 * jaicov_b200.h -- C ABI of the B200-native adjustment hot path for JAICOV
 * (applied-geodesy/bundle-adjustment).
 *
 * The library replaces what BundleAdjustment.estimateModel() does between "parameters are indexed"
 * and "results are exported": the body of the do/while loop at
 *   JAICOV/src/org/applied_geodesy/adjustment/bundle/BundleAdjustment.java:228-355
 * i.e. createNormalEquation (:789-834), NormalEquationSystem.applyPrecondition
 * (adjustment/NormalEquationSystem.java:82-91), MathExtension.solve = dspsv [+ dsptri]
 * (adjustment/MathExtension.java:338-366), getOmega (:472-491), updateUnknownParameters (:450-462), the
 * convergence test (:327-350) and the centroid shift (:115-201).
 *
 * The Java host keeps prepareUnknownParameters / detectRankDefect (integer bookkeeping, :667-782, :836-1042),
 * all setters/getters and the result writers; it hands the flattened object graph to this library once per
 * estimateModel() call (see INTEGRATION.md for the FFM/JNI stub).  All entry points are plain C: pointers and
 * sizes only, caller owns every buffer it passes and may free it when the call returns.
 *
 * Conventions
 *   - column indices are the reference's UnknownParameter columns AFTER the "+= d" renumbering
 *     (BundleAdjustment.java:776-781): -1 = unset, INT32_MAX = fixed (parameter/UnknownParameter.java:27),
 *     otherwise a column in [d, u+d).  Rows/columns [0,d) of N and Qxx are the datum border (:493-635).
 *   - return value of every function: an EstimationStateType id (adjustment/EstimationStateType.java:24-42)
 *     or JAICOV_OK (0) for calls that are not an estimation; never throws, never aborts.
 *   - Qxx is returned in MTJ UpperSymmPackMatrix layout: column-major packed upper,
 *     element (r,c), r<=c, at r + c(c+1)/2 -- what BundleAdjustment.getCofactorMatrix() (:1177-1179) hands to
 *     util/io/writer/MatlabResultWriter.java:92-223.
 *   - one handle = one adjustment = one host thread at a time.  There is no CPU fallback: every function
 *     that computes fails with JAICOV_NOT_INITIALISED when no sm_100 device is usable.
 */
#ifndef JAICOV_B200_H
#define JAICOV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* EstimationStateType ids (EstimationStateType.java:25-41) */
#define JAICOV_OK 0
#define JAICOV_ERROR_FREE_ESTIMATION 1
#define JAICOV_INTERRUPT (-1)
#define JAICOV_SINGULAR_MATRIX (-2)
#define JAICOV_NO_CONVERGENCE (-4)
#define JAICOV_NOT_INITIALISED (-5)
#define JAICOV_OUT_OF_MEMORY (-7)
/* not a reference id: bad argument / unsupported option (the Java side would throw IllegalArgumentException) */
#define JAICOV_ILLEGAL_ARGUMENT (-100)

/* progress states passed to the callback (names of EstimationStateType) */
#define JAICOV_STATE_BUSY 100
#define JAICOV_STATE_ITERATE 101
#define JAICOV_STATE_CONVERGENCE 102
#define JAICOV_STATE_INVERT_NORMAL_EQUATION_MATRIX 103
#define JAICOV_STATE_ESTIMATE_STOCHASTIC_PARAMETERS 104
#define JAICOV_STATE_LEVENBERG_MARQUARDT_STEP 105 /* old = previous, new = adapted damping value (BA:417-418) */

/* BundleAdjustment.MatrixInversion (BundleAdjustment.java:65-70) */
#define JAICOV_INVERT_NONE 0
#define JAICOV_INVERT_FULL 1
#define JAICOV_INVERT_PRE_ELIMINATION 2 /* BA:261-291, :1197-1453: needs jaicov_set_reduced_rows */
#define JAICOV_INVERT_REDUCED 3         /* idem; what the reference's three example programs use (ExampleReport.java:89) */

/* EstimationType (adjustment/EstimationType.java): only these two are accepted by the reference (:1132-1137) */
#define JAICOV_L2NORM 0
#define JAICOV_SIMULATION 1

/* Solver route.  The reference has one (LAPACK on the packed n x n system, MathExtension.java:338-366); both routes here
 * return its results.  STRUCTURED applies when no observation couples two object points (no scale bars, no directly
 * observed point groups): the object-coordinate block of N is then block diagonal and dx / Qxx follow from the reduced
 * camera system.  AUTO takes it when it applies, DENSE never, STRUCTURED fails if it does not. */
#define JAICOV_SOLVER_AUTO 0
#define JAICOV_SOLVER_DENSE 1
#define JAICOV_SOLVER_STRUCTURED 2

#define JAICOV_COL_UNSET (-1)
#define JAICOV_COL_FIXED 2147483647

/* ParameterType ids of distortion coefficients (bundle/parameter/ParameterType.java:27-56) */
#define JAICOV_PT_RADIAL_A 121
#define JAICOV_PT_TANGENTIAL_B 131
#define JAICOV_PT_TANGENTIAL_BX 132
#define JAICOV_PT_TANGENTIAL_BY 133
#define JAICOV_PT_AFFINITY_CX 141
#define JAICOV_PT_AFFINITY_CY 142
#define JAICOV_PT_DISTANCE_D 151
#define JAICOV_PT_ZERNIKE_X 161
#define JAICOV_PT_ZERNIKE_Y 162
#define JAICOV_PT_ZERNIKE_Z 163

typedef struct jaicov_handle jaicov_handle;

typedef struct {
    int32_t invert_mode;        /* JAICOV_INVERT_*; setInvertNormalEquation, BundleAdjustment.java:1146 (default FULL, :91) */
    int32_t estimation_type;    /* JAICOV_L2NORM / JAICOV_SIMULATION; setEstimationType, :1132 */
    int32_t max_iterations;     /* DefaultValue.getMaximalNumberOfIterations() = 5000 (DefaultValue.java:25) */
    int32_t use_centroid;       /* useCentroidedCoordinates, :1181 (default 1, :87) */
    int32_t apply_aposteriori;  /* applyAposterioriVarianceOfUnitWeight, :1185 (default 1, :86) */
    int32_t device;             /* CUDA device ordinal this handle binds to */
    int32_t solver;             /* JAICOV_SOLVER_* (default AUTO); the environment variable JAICOV_SOLVER=dense|structured overrides */
    int32_t n_devices;          /* 0 / 1: one GPU.  k > 1: ONE handle, ONE process, k GPUs [device, device + k) -- what a single JVM thread
                                 * can drive (BundleAdjustment.estimateModel runs on one thread, :203); see "multi-GPU" below */
    double sigma2apriori;       /* min(1, min variance) as accumulated by addObservationGroup, :637-643; <=0 -> 1 (:221) */
    double damping_value;       /* Levenberg-Marquardt lambda >= 0, setLevenbergMarquardtDampingValue :1189; 0 = plain Gauss-Newton (:96) */
} jaicov_options;

typedef struct {
    int32_t status;             /* last EstimationStateType id */
    int32_t iterations;         /* number of loop passes executed (incl. the final one) */
    int32_t iteration_step;     /* the reference's iterationStep = maxIter - runs (:230) at the last pass */
    int32_t n_unknowns;         /* u */
    int32_t n_datum;            /* d */
    int32_t n_observations;     /* 2 m + bars + observed rows */
    int32_t dof;                /* n_observations - u + d (:1080-1082) */
    int32_t solver_used;        /* JAICOV_SOLVER_DENSE or JAICOV_SOLVER_STRUCTURED: the route the last pass took */
    double omega;               /* v'Pv of the final pass (:429-430) */
    double max_abs_dx;          /* of the last pass (:432) */
    double sigma2apriori;
    double sigma2aposteriori;   /* getVarianceFactorAposteriori, :1090-1093 */
    double ms_assembly, ms_factor, ms_solve, ms_inverse, ms_omega, ms_total; /* device time of the LAST pass, CUDA events */
} jaicov_stats;

/* progress listener: replaces PropertyChangeSupport.firePropertyChange (:72, :205, :232, ...); called on the
 * calling thread only */
typedef void (*jaicov_progress_cb)(void *user, int32_t state, double old_value, double new_value);

/* ---- life cycle ------------------------------------------------------------------------------------------------- */
int32_t jaicov_default_options(jaicov_options *opt);
int32_t jaicov_create(const jaicov_options *opt, jaicov_handle **out);
void jaicov_destroy(jaicov_handle *h);
const char *jaicov_last_error(const jaicov_handle *h);
/* number of usable sm_100 devices (0 => every compute call fails loudly) */
int32_t jaicov_device_count(void);
/* Device buffers of destroyed handles are kept for the next handle of this process (cudaMalloc / cudaFree of tens of GB
 * cost hundreds of milliseconds; a failed allocation empties the cache and retries).  This call returns them to the
 * driver now; result = bytes released. */
int64_t jaicov_release_cached_memory(void);
/* diagnostic: number of CUDA kernels this library has launched in this process so far */
int64_t jaicov_launch_count(void);

/* Arithmetic of the big tile products (process-wide): 0 = FP64 tensor-core tiles (mma.sync DMMA) for every launch; 4..8 = launches of
 * at least 148 tiles and K >= 1024 are computed from exact int8 digit products on the tcgen05 tensor cores with FP64 recombination
 * (Ozaki scheme, csrc/ozaki.cu; 8 digits reproduce the FP64 results within the parity tolerances -- DESIGN.md section 9, profiles/).
 * digits < 0 only queries.  Returns the previous setting (-1: not decided yet = environment variable JAICOV_GEMM_OZAKI or the default). */
int32_t jaicov_set_gemm_digits(int32_t digits);

/* ---- multi-GPU ------------------------------------------------------------------------------------------------------
 * Two ways to put one adjustment on several GPUs of one box; both run the same device code.  Dense route: every GPU stores only
 * its own block-column panels of the system (about 1/P of it, jaicov_get_device_bytes), evaluates all observations and keeps what it
 * owns, and the block-column-cyclic Cholesky streams every factored panel over NVLink (NCCL broadcast) to where the trailing
 * updates, the solves and the GPU's own column tiles of Qxx consume it (JAICOV_DIST_STORAGE=replica: round 1's whole copy per
 * GPU).  Structured route: image-sharded assembly + all-reduce of the shared pieces, per-GPU column tiles of the Qxx products:
 *
 * (1) jaicov_options.n_devices = k: a SINGLE handle in a SINGLE process -- the form a Java host uses.  Every call is the
 *     same as on one GPU; internally each GPU gets its own host thread and its own communicator of one ncclCommInitAll clique
 *     (the progress listener is still only called on the calling thread).  All getters return complete results:
 *     jaicov_get_qxx_packed gathers the column tiles over NVLink (peer-to-peer) onto the first device and ships the MTJ-packed
 *     array in one stream; jaicov_get_qxx_block / _submatrix / jaicov_propagate_eo_transform sum the devices' parts.
 *
 * (2) one process per GPU (torchrun, MPI): rank 0 calls jaicov_nccl_unique_id and ships the 128 bytes to the other processes;
 *     every process then calls jaicov_dist_init on its handle BEFORE the set_* calls take effect.  Every rank passes the
 *     SAME full problem.  jaicov_get_qxx_block then returns PARTIAL blocks (zeros for entries owned elsewhere; the sum over
 *     ranks is the block), jaicov_get_qxx_local the rank's column tiles as stored.  An interrupt flag must be passed by every
 *     rank or by none (the ranks agree on the decision with one scalar all-reduce per check).
 *
 * Failure semantics (both forms): errors that follow from the problem (illegal arguments, a matrix that is not positive definite, the
 * iteration limit, an interrupt) are detected identically or agreed upon by all devices, and every device returns the same id.  An
 * error that strikes ONE device only (out of memory because another process holds that GPU, a device fault) is not agreed upon: the
 * other devices wait in their next collective, as in any NCCL program -- give every device of the handle the same free memory. */
/* pure host function: the contiguous image range [img_begin, img_end) rank `rank` of `world` works on (boundaries at
 * the images where the cumulative observation count crosses rank * m / world) */
int32_t jaicov_shard_images(int32_t n_img, const int64_t *pt_ptr, int32_t world, int32_t rank, int32_t *img_begin, int32_t *img_end);
int32_t jaicov_nccl_unique_id(void *out128);
int32_t jaicov_dist_init(jaicov_handle *h, int32_t rank, int32_t world, const void *nccl_id128);
/* n_tiles <- number of 128-wide column tiles of Qxx this rank owns; tile_first_col[i] <- reference column of tile i;
 * dst (may be NULL, host, pinned preferred) <- the tiles one after the other, tile i as a row-major block of
 * (np - e_i) rows x 128 columns with e_i = tile_first_col[i] - d and np = u rounded up to 128:
 * block_i[r][k] = Qxx[tile_first_col[i] + r][tile_first_col[i] + k], i.e. the rows on and below the tile's diagonal
 * (the part above the diagonal inside the first 128 rows is unspecified; Qxx is symmetric) */
int32_t jaicov_get_qxx_local(jaicov_handle *h, int32_t *n_tiles, int32_t *tile_first_col, int32_t tile_cap, double *dst);

/* ---- problem description (flattened object graph) ---------------------------------------------------------------- */
/* cameras: io_val/io_col hold x0, y0, c per camera (iterator order camera/orientation/InteriorOrientation.java:70-79);
 * coefficients of camera k are entries [coef_ptr[k], coef_ptr[k+1]) listed in the reference's evaluation order:
 * models by enum ordinal AFFINITY, TANGENTIAL, RADIAL, DISTANCE, ZERNIKE_X, ZERNIKE_Y, ZERNIKE_GRADIENT
 * (camera/Camera.java:50, camera/distortion/DistortionModel.java:29-37); inside TANGENTIAL: Bx, By, then Bi;
 * inside AFFINITY: Cx, Cy.  Fixed coefficients (col = JAICOV_COL_FIXED) must still be listed: they distort.
 * coef_order = PolynomialCoefficient.getOrder() (Zernike: the single index j), 0 for Bx/By/Cx/Cy.
 * Capacity: any number of cameras; at most 62 distortion coefficients in ONE camera (its Gram row [EO | IO | coefficients | w] is held in
 * nine 8-column tensor-core tiles), else the first computing call returns JAICOV_ILLEGAL_ARGUMENT.  (The reference has no limit; its
 * cameras carry at most a few dozen coefficients: 2 + 2 + n_A + n_B + n_D + Zernike terms.) */
int32_t jaicov_set_cameras(jaicov_handle *h, int32_t n_cam, const double *io_val, const int32_t *io_col, const double *r0,
                           const int32_t *coef_ptr, const int32_t *coef_type, const int32_t *coef_order,
                           const double *coef_val, const int32_t *coef_col);
/* images in camera->image iteration order; eo = X0,Y0,Z0,omega,phi,kappa
 * (camera/orientation/ExteriorOrientation.java:39-45); pt_ptr is a CSR over image points */
int32_t jaicov_set_images(jaicov_handle *h, int32_t n_img, const int32_t *cam_of_img, const double *eo_val,
                          const int32_t *eo_col, const int64_t *pt_ptr);
/* image points in the reference's observation order (rows 2j, 2j+1; :670-676); var = sigma_x^2, sigma_y^2
 * (camera/ImageCoordinate.java:50-51); rho = correlation coefficient (:54) */
int32_t jaicov_set_image_points(jaicov_handle *h, int64_t m, const int32_t *obj_idx, const double *xy, const double *var,
                                const double *rho);
/* object points; is_datum[i] != 0 iff the point is in BundleAdjustment.objectCoordinates and isDatum() (:501-513) */
int32_t jaicov_set_object_points(jaicov_handle *h, int32_t n_pt, const double *xyz, const int32_t *col,
                                 const uint8_t *is_datum);
/* scale bars (ScaleBar.java:34-39): end points, length, variance = sigma^2 */
int32_t jaicov_set_scale_bars(jaicov_handle *h, int32_t n_bar, const int32_t *a, const int32_t *b, const double *length,
                              const double *var);
/* one DirectlyObservedParameterGroup (parameter/DirectlyObservedParameterGroup.java:37-60).  target_kind:
 * 0 object point (index = point, comp 0..2), 1 interior orientation (index = camera, comp 0..2),
 * 2 distortion coefficient (index = position in the global coefficient list, comp ignored),
 * 3 exterior orientation (index = image, comp 0..5).  Either var (diagonal) or sigma_packed_upper
 * (r(r+1)/2, MTJ packed upper) must be given; with the latter P = sigma0^2 * Sigma^-1 (:80-90). */
int32_t jaicov_add_observed_group(jaicov_handle *h, int32_t r, const int32_t *target_kind, const int32_t *target_index,
                                  const int32_t *target_comp, const double *obs, const double *var,
                                  const double *sigma_packed_upper);
/* EXTENSION (north_star (2), BASELINE.json configs[3] "per-image dense Sigma_ll blocks"; not expressible in the reference, whose
 * image-coordinate groups are always the two rows of one point, camera/ImageCoordinate.java:102-104 -- parity unpinned beyond the
 * block-diagonal case): fully populated dispersion of the 2m image coordinates (x_0, y_0, x_1, y_1, ... in observation order) of
 * image `image`, MTJ packed upper, n_rows = 2m.  The image's points are then weighted with P = sigma0^2 Sigma^-1 as ONE
 * observation group of 2m rows (the contribution PDF:475-505 would stack for such a group); var / rho of those points are
 * ignored.  n_rows = 0 removes it.  Couples all points of the image: dense solver route, one device. */
int32_t jaicov_set_image_dispersion(jaicov_handle *h, int32_t image, int64_t n_rows, const double *sigma_packed_upper);
/* result of detectRankDefect (:836-1042): free_flags in the order tx,ty,tz,rx,ry,rz,scale (1 = FREE);
 * n_unknowns = numberOfUnknownParameters (:80), n_observations = numberOfObservations (:81) */
int32_t jaicov_set_datum(jaicov_handle *h, const int32_t free_flags[7], int32_t n_unknowns, int32_t n_observations);

/* MatrixInversion.REDUCED / PRE_ELIMINATION: numRows of the reduced system = numberOfInteriorOrientations +
 * numberOfDistortionParameters + 3 * objectCoordinates.size() + d (BundleAdjustment.java:262).  The library does not
 * form the reference's Schur complement on the 6x6 EO blocks (:1225-1342): the reduced solution and the reduced
 * cofactor matrix are the leading numRows block of the full ones, which the tensor-core path computes anyway, so both
 * modes run the full path and expose rows/columns [0, numRows) of Qxx only (jaicov_get_qxx_packed then returns
 * numRows(numRows+1)/2 doubles).  The exact dx of the EO parameters replaces the reference's final-pass leftovers
 * (:273 applies the un-solved right-hand side as dx_EO; below the parity tolerance at convergence). */
int32_t jaicov_set_reduced_rows(jaicov_handle *h, int32_t num_rows);

/* ---- the adjustment ------------------------------------------------------------------------------------------------ */
/* Runs the whole loop on the device and returns the EstimationStateType id. interrupt_flag (may be NULL) is
 * polled twice per pass like BundleAdjustment.interrupt (:240, :320). */
int32_t jaicov_estimate(jaicov_handle *h, jaicov_progress_cb cb, void *user, volatile int32_t *interrupt_flag);

/* One pass of the loop body on the CURRENT values (benchmark / stage tests): assembly + datum + precondition,
 * factor + solve, and if final_pass != 0 also the inversion (if invert_mode FULL) and Omega.  If apply_update
 * != 0 the unknowns are updated (x += dx) as the reference does. */
int32_t jaicov_iterate(jaicov_handle *h, int32_t final_pass, int32_t apply_update);

/* ---- results ----------------------------------------------------------------------------------------------------- */
int32_t jaicov_get_stats(jaicov_handle *h, jaicov_stats *out);
/* current parameter values (any pointer may be NULL) in the layouts of the set_* calls */
int32_t jaicov_get_values(jaicov_handle *h, double *xyz, double *io_val, double *coef_val, double *eo_val);
/* last solution vector dx in reference column order, length u+d (border entries first) */
int32_t jaicov_get_dx(jaicov_handle *h, double *dx);
/* Qxx (u+d) in MTJ packed-upper layout, (u+d)(u+d+1)/2 doubles; dst is host memory (pinned or pageable) */
int32_t jaicov_get_qxx_packed(jaicov_handle *h, double *dst);
/* rectangular tile rows [r0,r1) x cols [c0,c1) of the full symmetric Qxx, row-major with leading dimension ld */
int32_t jaicov_get_qxx_block(jaicov_handle *h, int32_t r0, int32_t r1, int32_t c0, int32_t c1, double *dst, int64_t ld);
/* dst[i * n_idx + j] = scale * Qxx[idx[i]][idx[j]] (reference column numbers, border included): the sub-matrix the
 * result writers export -- MatlabResultWriter "dispersion" (scale 1, util/io/writer/MatlabResultWriter.java:209-223) and
 * DefaultResultWriter ".cxx" (scale sigma0^2 a posteriori, util/io/writer/DefaultResultWriter.java:126-155) -- gathered on
 * the device so that the full Qxx never has to travel to the host.  Distributed handle: partial (sum over ranks). */
int32_t jaicov_get_qxx_submatrix(jaicov_handle *h, int32_t n_idx, const int32_t *idx, double scale, double *dst);
/* diagonal of Qxx, length u+d */
int32_t jaicov_get_qxx_diag(jaicov_handle *h, double *dst);

/* ---- stage access for parity tests and profiling (same kernels the loop uses) ---------------------------------- */
/* DirectLinearTransformation.RestrictionType ordinals (dlt/DirectLinearTransformation.java:50-57) */
#define JAICOV_DLT_IDENTICAL_PRINCIPLE_DISTANCE 0
#define JAICOV_DLT_ROTATION_WITHOUT_SHEAR 1
#define JAICOV_DLT_FIXED_PRINCIPLE_DISTANCE_X 2
#define JAICOV_DLT_FIXED_PRINCIPLE_DISTANCE_Y 3
#define JAICOV_DLT_FIXED_PRINCIPAL_POINT_X 4
#define JAICOV_DLT_FIXED_PRINCIPAL_POINT_Y 5

/* Batched direct linear transformation: initial interior / exterior orientation of n_img images at once,
 * replacing one DirectLinearTransformation.adjust call per image (dlt/DirectLinearTransformation.java:67-184; normal
 * equations dlt/DLTPartialDerivativeFactory.java:239-337, restrictions :68-236, expansion DLT:186-266).  The caller
 * gathers the homologous points (image point + object coordinates matched by name, DLT:78-94): image i owns the
 * observations [pt_ptr[i], pt_ptr[i+1]) of xy (2 doubles each) and xyz (3 doubles each); io = (c, x0, y0) of the image's
 * camera (used by the FIXED_* restrictions); the restrictions apply to every image.  out20 per image: b11 b12 b13 b14 b21
 * b22 b23 b24 b31 b32 b33 (back-scaled), c = (cx + cy)/2, x0, y0, X0, Y0, Z0, omega, phi, kappa.  status per image: 1 =
 * adjust() returned true, 0 = no convergence, -1 = failed (fewer than 6 points, singular system); passes (may be NULL) =
 * normal-equation solves.  No handle: like jaicov_spd_solve_invert the call is self-contained. */
int32_t jaicov_dlt_batch(int32_t device, int32_t n_img, const int64_t *pt_ptr, const double *xy, const double *xyz, const double *io,
                         int32_t n_restrictions, const int32_t *restrictions, int32_t max_iterations, double *out20,
                         int32_t *status, int32_t *passes);

/* Covariance propagation of object points transformed into the frame of a reference image, replacing
 * CoordinateTransformationExteriorOrientation.transform (tranformation/CoordinateTransformationExteriorOrientation.java:49-121):
 * for every triple (point, source image, target image)  X_trg = X0_trg + R_trg R_src' (X - X0_src)  (:209-215; source ==
 * target: identity, :141-149), J = d X_trg / d (EO_trg, EO_src, X)  (:223-279) and
 *   cov = sigma2 * J Qxx J'   (:110-114), packed upper like MTJ's UpperSymmPackMatrix, (3 n_points)(3 n_points + 1)/2 doubles.
 * The caller enumerates the triples (the visibility loops of :57-98 stay on the host side); point / image indices are those
 * of jaicov_set_object_points / jaicov_set_images; fixed parameters contribute nothing.  Qxx is contracted where it lives
 * (15 x 15 gathers per 3 x 3 block); needs a final pass with a cofactor matrix.  xyz_out (3 n_points) or cov_packed may be
 * NULL.  Multi-GPU handles return the rank's partial sums (the sum over ranks is cov). */
int32_t jaicov_propagate_eo_transform(jaicov_handle *h, int32_t n_points, const int32_t *point, const int32_t *src_image,
                                      const int32_t *trg_image, double sigma2, double *xyz_out, double *cov_packed);

/* K1: residuals and compact Jacobian of every image point at the current values.  Per point j the 2 x ns
 * entries are written slot-ordered: slots 0..11 = X,Y,Z,x0,y0,c,X0,Y0,Z0,omega,phi,kappa, slots 12.. = the
 * coefficients of the point's camera in list order; ns_max = 12 + max coefficients per camera.
 * a: [m][2][ns_max], w: [m][2], p: [m][3] (P00,P01,P11); host buffers. */
int32_t jaicov_eval_residual_jacobian(jaicov_handle *h, int32_t ns_max, double *a, double *w, double *p);
/* Normal equations of the current values BEFORE datum/preconditioning: N (u+d, MTJ packed upper, border rows
 * as written by addDatumConditionRows) and n (u+d); host buffers; either may be NULL */
int32_t jaicov_get_normal_equations(jaicov_handle *h, double *n_packed, double *rhs);
/* Omega = v'Pv for a given dx (reference column order, length u+d) at the current values (getOmega, :472-491) */
int32_t jaicov_omega(jaicov_handle *h, const double *dx, double *omega);

/* Verification at any size: the product with the bordered normal matrix K = [[0, B], [B', N]] of createNormalEquation
 * (:789-834, N = A'PA over image points PDF:285-445 + scale bars PDF:210-283 + directly observed groups PDF:447-473, B = the
 * datum rows :493-635; no Levenberg-Marquardt damping), evaluated MATRIX-FREE from the observations at the current values:
 * y_v = K x_v for nvec vectors x, y = [nvec][u+d] (reference column order, host buffers), rhs (may be NULL, u+d) <- [0; A'Pw],
 * wpw (may be NULL) <- w'Pw.  Nothing of the assembled matrix, its factor or Qxx is read, so a caller can check
 *   K [lambda; dx] = [0; n],   K Qxx e_c = e_c (columns from jaicov_get_qxx_block),   Omega = w'Pw - 2 n'dx + dx'N dx
 * on configurations no CPU reference can reach (bench.py does, at every N).  Distributed handles sum their image shards
 * (NCCL all-reduce); every rank receives the complete product.  Does not disturb Qxx / dx of the last pass. */
int32_t jaicov_normal_product(jaicov_handle *h, int32_t nvec, const double *x, double *y, double *rhs, double *wpw);
/* Device time (CUDA events, ms) of the three observation sweeps of the LAST pass: the by-image sweep (model evaluation + per-image
 * Gram + the unique EO x point blocks), the by-point sweep (first camera group) and the Omega sweep (0 unless it was a final pass).
 * These are the kernels whose algorithmic traffic is 44 B per image point; bench.py turns them into GB/s. */
int32_t jaicov_get_sweep_times(jaicov_handle *h, double *ms_by_image, double *ms_by_point, double *ms_omega);
/* Device memory of the big buffers of this handle, in bytes: out[0] the system matrix N -> L (whole lower-triangular square on one GPU;
 * ONLY THE RANK'S OWN block-column panels -- about 1/P of it -- on the distributed dense route), out[1] the second square of the
 * single-GPU inverse (0 otherwise), out[2] the rank's column tiles of Qxx (distributed), out[3] the two panel staging slots.
 * Single-process multi-GPU handle: the largest device's figures. */
int32_t jaicov_get_device_bytes(jaicov_handle *h, int64_t out[4]);
/* Jacobi preconditioner V of the last pass (:824-828), length u+d (1 on the border) */
int32_t jaicov_get_preconditioner(jaicov_handle *h, double *v);

/* The tensor-core tile product behind every O(n^3) stage (factor, inverse, the structured route's products), exposed
 * for parity tests and profiling:  C = alpha * op(A) op(B)' + beta * C  on 128 x 128 tiles, host buffers.
 * op(A) is (128 mt) x K: a_layout 0 = stored row-major with leading dimension lda (k contiguous), 1 = stored as K rows of
 * leading dimension lda (the tile-row index contiguous); op(B) is (128 nt) x K, b_layout likewise; C is (128 mt) x (128 nt)
 * row-major (ldc).  kmode: 0 full contraction; 1 = op(B)'[k][n] lower triangular (k >= n: starts at the tile column's
 * diagonal); 2 = op(A)[m][k] lower triangular (k <= m: stops after the tile row's diagonal); 3 = both operands stored
 * [k][.] lower triangular (k >= max(m, n)) -- the hints the schedule of csrc/dense_driver.hpp passes.  tri_out != 0 (mt == nt):
 * only tiles on or below the diagonal are computed.  reps >= 1 launches are timed (forced to 1 when beta != 0);
 * ms (may be NULL) <- average device time of one launch.  No handle; fails with JAICOV_NOT_INITIALISED without a device. */
int32_t jaicov_gemm_tiles(int32_t device, int32_t a_layout, int32_t b_layout, int32_t mt, int32_t nt, int64_t K, double alpha,
                          double beta, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
                          int32_t tri_out, int32_t kmode, int32_t reps, double *ms);

/* Level-1 seam (the routines BundleAdjustment needs from LAPACK, MathExtension.java:304-366), dense SPD route:
 * in: a = symmetric positive definite n x n, row-major, host; b = nrhs right-hand sides [nrhs][n] (may be NULL).
 * out: a <- a^-1 (full symmetric) if invert != 0, b <- solutions.  Returns SINGULAR_MATRIX if not SPD. */
int32_t jaicov_spd_solve_invert(int32_t device, int64_t n, double *a, int32_t nrhs, double *b, int32_t invert,
                                double *ms_factor, double *ms_inverse);

#ifdef __cplusplus
}
#endif
#endif /* JAICOV_B200_H */
