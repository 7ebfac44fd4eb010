/* ezcalib_b200.h -- C ABI of the B200-native EZCalib hot path.
 *
 * The reference (ferreram/ezcalib) has no FFI layer: its hot path sits behind C++ interfaces
 * (CalibDetector, DistParam, IntrinsicsParam, AutoDiffLocalLeftSE3) that ezcalibrator.cpp and
 * camera.cpp program against.  Each entry point below names the reference call it replaces;
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions: plain C, POD only; caller owns every host buffer; the library owns device memory
 * behind opaque handles; in/out arrays are overwritten only on success; every function returns an
 * ezc_status and never aborts (the reference exit(-1)s; see SURVEY.md 8b "Error conventions").
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * EZC_ERR_CUDA.
 */
#ifndef EZCALIB_B200_H_
#define EZCALIB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  EZC_OK = 0,
  EZC_ERR_INVALID = 1,   /* bad argument */
  EZC_ERR_CUDA = 2,      /* CUDA runtime failure (incl. no device) */
  EZC_ERR_NUMERIC = 3,   /* non-finite cost at the starting point (Ceres: FAILURE) */
  EZC_ERR_CAPACITY = 4,  /* an internal per-image capacity was exceeded */
  EZC_ERR_COLLECTIVE = 5 /* the multi-GPU exchange failed (callback returned non-zero) */
} ezc_status;

typedef struct ezc_context ezc_context;

/* Collective hook: all-reduce `count` doubles IN PLACE in DEVICE memory, enqueued on `stream` (it must not block: the
 * solver enqueues whole groups of LM iterations without synchronising).  op: 0 = sum, 1 = max.  Returns 0 on success; any
 * other value makes the running ezc_lm_* call fail with EZC_ERR_COLLECTIVE -- the caller must then abort the process group
 * (its peers are blocked in the same collective).  Production: ezc_comm_init's own exchange (see below) or
 * torch.distributed (NCCL over NVLink); tests: gloo. */
typedef int32_t (*ezc_allreduce_fn)(void* user, void* device_buf, int32_t count, int32_t op, void* stream);

ezc_status ezc_create(int32_t device, ezc_context** out);
void ezc_destroy(ezc_context* ctx);
/* Thread-local message of the last failure on this thread. */
const char* ezc_last_error(void);
/* Run all work of this context on an existing cudaStream_t (e.g. torch's current stream). */
ezc_status ezc_set_stream(ezc_context* ctx, void* cuda_stream);
/* View-sharded multi-GPU: each rank owns a contiguous slice of the pose blocks. */
ezc_status ezc_set_allreduce(ezc_context* ctx, ezc_allreduce_fn fn, void* user, int32_t rank, int32_t world);
/* The library's own exchange for the view-sharded solve (one process per GPU on one NVLink / NVSwitch node): a one-shot
 * all-reduce through peer memory -- every rank stores its <= 256-double packet into a mailbox slot of every peer, publishes a
 * sequence flag, waits for the world's flags in its own mailbox and sums the slots in FIXED RANK ORDER (bit-identical on all
 * ranks, independent of arrival order).  It replaces the two NCCL launches per LM iteration behind ezc_set_allreduce.
 *   1. every rank: ezc_comm_mailbox(ctx, world, NULL, handle)     allocates the mailbox, returns its 64-byte CUDA IPC handle
 *   2. the caller all-gathers the handles with whatever it has (MPI, torch.distributed, a socket)
 *   3. every rank: ezc_comm_init(ctx, rank, world, handles)       maps the peers' mailboxes and installs the exchange
 * ezc_comm_attach is the same for mailboxes that are plain device pointers in THIS process (several contexts in one process). */
ezc_status ezc_comm_mailbox(ezc_context* ctx, int32_t world, void** mailbox_out, uint8_t* ipc_handle_out /*[64] or NULL*/);
ezc_status ezc_comm_init(ezc_context* ctx, int32_t rank, int32_t world, const uint8_t* all_ipc_handles /*[world][64]*/);
ezc_status ezc_comm_attach(ezc_context* ctx, int32_t rank, int32_t world, void* const* mailboxes /*[world]*/);
/* Number of kernels of this library launched on the context so far (bench.py "gpu_launches"). */
uint64_t ezc_kernel_launches(const ezc_context* ctx);
/* Give the cached device blocks of `device` (-1: every device) back to the driver.  The cache is per device and is
 * also flushed when the last context of a device is destroyed. */
ezc_status ezc_trim_cache(int32_t device);
ezc_status ezc_synchronize(ezc_context* ctx);
/* Roofline denominator for the FP64-bound LM kernels: DFMA-saturating microbenchmark (TFLOP/s). */
ezc_status ezc_measure_fp64_peak(ezc_context* ctx, double* tflops_out);
/* Same for the FP64 tensor-core path (mma.sync m8n8k4 f64) used by the Gram accumulation. */
ezc_status ezc_measure_dmma_peak(ezc_context* ctx, double* tflops_out);

/* ------------------------------------------------------------------------------------------
 * Half (a): batched calibration-target detection.
 * ---------------------------------------------------------------------------------------- */

/* A batch of 8-bit gray images of identical size: what cv::imread(path, IMREAD_GRAYSCALE) yields in
 * EZCalibrator::runCalibration (/root/reference/src/ezcalibrator.cpp:44).  row_stride / image_stride in
 * bytes; 0 means densely packed. */
typedef struct {
  const uint8_t* data;
  int32_t n_images, width, height;
  int64_t row_stride, image_stride;
} ezc_image_batch;

typedef struct ezc_images ezc_images;

/* Copy a HOST batch into HBM (the detectors then work on the device-resident copy). */
ezc_status ezc_images_upload(ezc_context* ctx, const ezc_image_batch* host_batch, ezc_images** out);
/* The same, but returns as soon as the copies are QUEUED on a second stream; the detectors start on the first
 * sub-batch while later images are still crossing PCIe (what the reference's loop does implicitly by reading image i+1
 * after processing image i, ezcalibrator.cpp:40-59).  The host buffer must stay valid and unmodified until the first
 * call that consumes the whole batch has returned (or ezc_images_destroy); pinned host memory is needed for real overlap. */
ezc_status ezc_images_upload_async(ezc_context* ctx, const ezc_image_batch* host_batch, ezc_images** out);
/* Borrow a batch that already lives in DEVICE memory (batch->data is a device pointer). */
ezc_status ezc_images_wrap_device(ezc_context* ctx, const ezc_image_batch* device_batch, ezc_images** out);
void ezc_images_destroy(ezc_images* imgs);

/* Copy images [first, first+count) back to HOST memory (densely packed). */
ezc_status ezc_images_download(ezc_context* ctx, const ezc_images* imgs, int32_t first, int32_t count, uint8_t* host_out);

/* Synthetic data generator (not part of the reference's path; the reference ships no images): renders n_images views
 * of a planar target directly into HBM.  pattern 0 = chessboard (rows x cols SQUARES of `size` metres),
 * 1 = AprilGrid 36h11 (rows x cols tags of `size` metres, gap ratio `space`).  cams12: HOST [n][12] = fx fy cx cy
 * dist[8] (reference coefficient order); poses7: HOST [n][7] world->camera. */
ezc_status ezc_render_targets(ezc_context* ctx, int32_t pattern, int32_t rows, int32_t cols, double size, double space,
                              int32_t model, int32_t width, int32_t height, int32_t n_images, const double* cams12,
                              const double* poses7, double noise_sigma, uint64_t seed, ezc_images** out);

/* cv::cornerSubPix(img, corners, Size(win_w,win_h), Size(-1,-1), {COUNT+EPS, max_iter, eps}) for n corners;
 * corner i lies in image image_index[i].  Replaces /root/reference/include/target_detectors/
 * chessboard_detector.hpp:39-40 and aprilgrid_detector.hpp:52-53 (win 5x5, 100 iterations, eps 0.05).
 * corners: HOST float [n][2], refined in place.  Corners must lie inside the image (OpenCV asserts it). */
ezc_status ezc_corner_subpix(ezc_context* ctx, const ezc_images* imgs, float* corners, const int32_t* image_index,
                             int32_t n, int32_t win_w, int32_t win_h, int32_t max_iter, double eps);

/* ChessboardCalibDetector::detectTarget (/root/reference/include/target_detectors/chessboard_detector.hpp:27-50) for
 * every image of the batch: cv::findChessboardCorners(img, Size(nb_cols-1, nb_rows-1), ADAPTIVE_THRESH + NORMALIZE_IMAGE +
 * FILTER_QUADS + FAST_CHECK) followed by cv::cornerSubPix (5,5)/(100,0.05).  nb_rows / nb_cols count SQUARES, like the
 * reference's constructor.  corners_out: HOST float [n_images][(nb_cols-1)*(nb_rows-1)][2] in OpenCV's corner order
 * (zeros when not found); found_out: HOST [n_images]. */
ezc_status ezc_detect_chessboard_batch(ezc_context* ctx, const ezc_images* imgs, int32_t nb_rows, int32_t nb_cols,
                                       float* corners_out, uint8_t* found_out, int32_t max_images_in_flight);
/* The same search with an explicit inner-corner pattern; final_subpix = 0 returns what cv::findChessboardCorners itself
 * returns (after its internal (2,2) refinement), for parity tests against cv2.  fast_check != 0 applies the
 * CALIB_CB_FAST_CHECK gate (checkChessboardBinary / checkChessboard) before the search, like the reference's flags do. */
ezc_status ezc_find_chessboard_corners_batch(ezc_context* ctx, const ezc_images* imgs, int32_t pattern_w, int32_t pattern_h,
                                             int32_t final_subpix, int32_t fast_check, float* corners_out, uint8_t* found_out,
                                             int32_t max_images_in_flight);

/* AprilTagCalibDetector::detectTarget (/root/reference/include/target_detectors/aprilgrid_detector.hpp:30-71) for
 * every image of the batch: AprilTags::TagDetector(tagCodes36h11).extractTags
 * (/root/reference/thirdparty/ethz_apriltag2/src/TagDetector.cc:40-586), cv::cornerSubPix (5,5)/(100,0.05) on the
 * 4 corners of every detection with id < nb_rows*nb_cols, scatter into slot id*4+i.
 * corners_out: HOST float [n_images][4*nb_tags][2], (-1,-1) = not detected;
 * n_detections_out: HOST [n_images] = vdet.size(); success_out: HOST [n_images] = vdet.size() > nb_tags/2 - 1.
 * max_images_in_flight <= 0 lets the library size its sub-batches. */
ezc_status ezc_detect_aprilgrid_batch(ezc_context* ctx, const ezc_images* imgs, int32_t nb_rows, int32_t nb_cols,
                                      float* corners_out, int32_t* n_detections_out, uint8_t* success_out,
                                      int32_t max_images_in_flight);

/* Test hooks: stage products of extractTags for ONE image (theta/mag images, merge-order edge list, union-find
 * representatives, segments, quads), compared stage by stage against the reference objects in tests/. */
typedef struct ezc_apriltag_debug ezc_apriltag_debug;
ezc_status ezc_apriltag_debug_run(ezc_context* ctx, const ezc_images* imgs, int32_t image, int32_t nb_rows, int32_t nb_cols,
                                  ezc_apriltag_debug** out, float* corners_out, int32_t* n_detections_out,
                                  uint8_t* success_out);
/* what: 0 pixels, 1 edges, 2 segments, 3 quads */
int32_t ezc_apriltag_debug_count(const ezc_apriltag_debug* d, int32_t what);
ezc_status ezc_apriltag_debug_get(const ezc_apriltag_debug* d, float* theta, float* mag, int32_t* edges /*[n][3]*/,
                                  int32_t* rep, int32_t* size, float* segments /*[n][6]*/, float* quads /*[n][9]*/);
void ezc_apriltag_debug_free(ezc_apriltag_debug* d);

/* Test hook: the device restatements of the host libm functions the AprilTag path needs bit for bit (csrc/libm_exact.cuh):
 * which = 0: atan2f(a[i], b[i]) (glibc / fdlibm), 1: sinf(a[i]), 2: cosf(a[i]) (glibc's ARM optimised routines).  HOST pointers. */
ezc_status ezc_debug_libm(ezc_context* ctx, int32_t which, const float* a, const float* b, int64_t n, float* out);

/* ------------------------------------------------------------------------------------------
 * Between the halves: initial poses.  EZCalibrator::computeP3PRansac
 * (/root/reference/src/ezcalibrator.cpp:560-628) for every view of a batch, one warp per view:
 * pixels inside the image (float isPointInImage, calib_models.hpp:107-115) are normalised with K^-1 of the PRIOR
 * intrinsics (:580) and, with the model points cast to float (:582), go through
 * cv::solvePnPRansac(identity K, no distortion, max_iterations = 1000, reproj_threshold = 4 / mean focal,
 * confidence = 0.999, SOLVEPNP_AP3P) (:592-606): RANSAC over 4-point samples with cv::RNG's own sequence, then EPnP on
 * the inliers; a view fails with fewer than min_points (32) in-image points or inliers when the target has that many
 * (:586-590, :610).  The result is converted like :616-625 (Rodrigues -> float -> Sophus::SO3d::fitToSO3).
 * corners: HOST float [n_views][n_pts][2]; target: HOST double [n_pts][3]; focal[n_focal = 1|2], pp[2]: prior intrinsics.
 * ok_out: HOST [n_views]; poses7_out: HOST [n_views][7] = qx qy qz qw tx ty tz (zeros when !ok);
 * optional (may be NULL): n_inliers_out [n_views], inlier_mask_out [n_views][n_pts] (per TARGET point),
 * rvec_tvec_out [n_views][6] = what cv::solvePnPRansac itself returns, n_iterations_out [n_views] RANSAC iterations run. */
ezc_status ezc_p3p_ransac_batch(ezc_context* ctx, const float* corners, const double* target, int32_t n_views, int32_t n_pts,
                                int32_t n_target_pts, const double* focal, int32_t n_focal, const double* pp, int32_t width,
                                int32_t height, int32_t max_iterations, double reproj_threshold, double confidence,
                                int32_t min_points, uint8_t* ok_out, double* poses7_out, int32_t* n_inliers_out,
                                uint8_t* inlier_mask_out, double* rvec_tvec_out, int32_t* n_iterations_out);

/* ------------------------------------------------------------------------------------------
 * Half (b): Levenberg-Marquardt calibration solve.
 * Replaces the ceres::Problem built in EZCalibrator::computeCalibration
 * (/root/reference/src/ezcalibrator.cpp:105-229) and runMultiCameraCalib (:631-844), the
 * cost functions from DistParam::createCeresCostFunction (/root/reference/include/dist_models/
 * dist_model_base.hpp:73-74) and the AutoDiffLocalLeftSE3 manifold
 * (/root/reference/include/ceres_params/se3_param.hpp:12-41).
 * ---------------------------------------------------------------------------------------- */

/* dist_model ids, in the order of Camera::setupInitialDistortion (/root/reference/src/camera.cpp:67-94) */
enum { EZC_RAD1 = 0, EZC_RAD2 = 1, EZC_RAD3 = 2, EZC_RADTAN4 = 3, EZC_RADTAN5 = 4, EZC_RADTAN8 = 5, EZC_KB4 = 6 };

/* Number of distortion coefficients of a model (DistParam::getNumberOfParameters), -1 if unknown. */
int32_t ezc_num_dist(int32_t model);

typedef struct {
  int32_t model;             /* EZC_RAD1 .. EZC_KB4 */
  int32_t use_mono_focal;    /* 1: one focal, 0: fx,fy */
  int32_t n_blocks;          /* SE3 blocks owned by THIS rank: views (mono) or extrinsics (multi-cam) */
  int32_t obs_per_block;     /* observations per block; mono: number of target points */
  int32_t pts_period;        /* 3-D point of local observation l is pts[l % pts_period] */
  int32_t cams_per_block;    /* 0: one shared (focal,pp,dist); 1: one constant set per block (multi-cam) */
  double img_w, img_h;       /* IntrinsicsParam::isPointInImage bounds (calib_models.hpp:107-115) */
  const double* pts;         /* [pts_period][3] */
  const float* obs;          /* [n_blocks][obs_per_block][2]; cv::Point2f; (-1,-1) = missing */
  int32_t n_blocks_global;   /* total blocks over all ranks (== n_blocks when world == 1) */
  int32_t collective;        /* on a context with an exchange installed (world > 1): 1 = this solve is view-sharded over the
                                ranks (every rank must make the same calls), -1 = local to this rank (e.g. the per-camera mono
                                calibrations of a rig), 0 = decide from n_blocks_global != n_blocks */
} ezc_lm_problem;

typedef struct {
  int32_t opt_focal, opt_pp, opt_dist;   /* which shared blocks are variable (SetParameterBlockConstant) */
  int32_t max_num_iterations;            /* reference: 1000 (ezcalibrator.cpp:172) */
  double function_tolerance;             /* Ceres defaults below */
  double gradient_tolerance;
  double parameter_tolerance;
  double initial_trust_region_radius;
  double max_trust_region_radius;
  double min_trust_region_radius;
  double min_relative_decrease;
  double min_lm_diagonal, max_lm_diagonal;
  int32_t max_num_consecutive_invalid_steps;
  int32_t trace_capacity;                /* entries available in the three trace arrays (may be 0) */
  double* trace_cost;                    /* per iteration: cost after the iteration */
  double* trace_radius;                  /*                trust-region radius */
  int32_t* trace_flag;                   /* 0 iteration zero, 1 successful, 2 unsuccessful, 3 invalid */
} ezc_lm_options;

typedef struct {
  int32_t iterations;        /* iteration summaries incl. iteration 0 (ceres Summary::iterations.size()) */
  int32_t num_successful;    /* incl. iteration 0 */
  int32_t termination;       /* 0 function tol, 1 parameter tol, 2 gradient tol, 3 max iterations,
                                4 min radius, 5 failure */
  int32_t n_residual_blocks;
  double initial_cost, final_cost;
  int32_t jacobian_evaluations;
  int32_t linear_solves;
} ezc_lm_summary;

typedef struct ezc_lm_solver ezc_lm_solver;

/* Ceres 2.2.0 defaults + the reference's max_num_iterations = 1000; all shared blocks variable. */
void ezc_lm_default_options(ezc_lm_options* opt);

/* Upload observations / points (HOST pointers) and allocate the device state. */
ezc_status ezc_lm_create(ezc_context* ctx, const ezc_lm_problem* problem, ezc_lm_solver** out);
/* Same, but problem->pts / problem->obs are DEVICE pointers that the solver borrows (no copy). */
ezc_status ezc_lm_create_device(ezc_context* ctx, const ezc_lm_problem* problem, ezc_lm_solver** out);
void ezc_lm_destroy(ezc_lm_solver* s);

/* focal[1|2], pp[2], dist[nd] (or [n_blocks][..] when cams_per_block), poses[n_blocks][7] = qx qy qz qw tx ty tz
 * (Sophus::SE3d::data() order).  HOST pointers. */
ezc_status ezc_lm_set_params(ezc_lm_solver* s, const double* focal, const double* pp, const double* dist,
                             const double* poses);
ezc_status ezc_lm_get_params(ezc_lm_solver* s, double* focal, double* pp, double* dist, double* poses);

/* One ceres::Solve(options, &problem, &summary) on the device-resident state. */
ezc_status ezc_lm_solve(ezc_lm_solver* s, const ezc_lm_options* opt, ezc_lm_summary* summary);

/* Per-block RMS reprojection error, CalibFrame::rmse_err (/root/reference/src/ezcalibrator.cpp:236-274).
 * Blocks without in-image observation get DBL_MAX.  rmse_out: HOST [n_blocks]. */
ezc_status ezc_lm_frame_rmse(ezc_lm_solver* s, double* rmse_out);

/* Covariance of the shared parameters (ceres::Covariance on focal, pp, dist;
 * /root/reference/src/ezcalibrator.cpp:293-331).  cov_out: HOST [k][k] row-major, k = nf+2+nd. */
ezc_status ezc_lm_covariance(ezc_lm_solver* s, double* cov_out);

/* The whole of EZCalibrator::computeCalibration's optimisation: 3 staged solves
 * (focal+poses; +dist; +pp) and the per-frame RMSE.  All pointers HOST; host<->device copies inside. */
ezc_status ezc_lm_mono(ezc_context* ctx, const ezc_lm_problem* problem, double* focal, double* pp, double* dist,
                       double* poses, double* frame_rmse_out, ezc_lm_summary summaries[3]);

/* The optimisation of EZCalibrator::runMultiCameraCalib (/root/reference/src/ezcalibrator.cpp:631-844): block b is the
 * extrinsic T_c(b+1)_c0 of camera b+1 with its own CONSTANT intrinsics (problem->cams_per_block = 1;
 * focal [n_blocks][1|2], pp [n_blocks][2], dist [n_blocks][nd]); problem->pts are cam0's target points T_c0_w * p_w of
 * every usable cam0 frame (:731-736), problem->obs the matching corners of camera b+1 ((-1,-1) where there is no pair).
 * One ceres::Solve over all blocks (one radius, one termination).  extrinsics: HOST [n_blocks][7] in/out.
 * View-sharded multi-GPU: each rank passes the blocks (cameras) it owns, n_blocks_global = all of them. */
ezc_status ezc_lm_multicam(ezc_context* ctx, const ezc_lm_problem* problem, const double* focal, const double* pp, const double* dist,
                           double* extrinsics, ezc_lm_summary* summary);

/* Debug / test entry: raw residual r[2] and Jacobian J[2][nf+2+nd+6] (columns focal|pp|dist|pose6) of one
 * observation, evaluated by the same device code the solver uses. */
ezc_status ezc_lm_residual_jacobian(ezc_context* ctx, int32_t model, int32_t use_mono_focal, const double* focal,
                                    const double* pp, const double* dist, const double* pose7, const double* wpt3,
                                    double u, double v, double* r, double* J);

/* Debug / test entry: robustified normal equations at the current parameters for the given
 * variable pattern.  H_cc[k*k], g_c[k], per block H_pp[36], H_pc[6*k], g_p[6]; cost.  HOST pointers. */
ezc_status ezc_lm_normal_equations(ezc_lm_solver* s, int32_t opt_focal, int32_t opt_pp, int32_t opt_dist,
                                   double* H_cc, double* g_c, double* H_pp, double* H_pc, double* g_p, double* cost);

/* Bench / profiling entry: enqueue ONLY the Jacobian + normal-equation build (lm_accumulate and its
 * fixed-order reductions) at the current parameters, asynchronously on the context stream. */
ezc_status ezc_lm_accumulate_only(ezc_lm_solver* s, int32_t opt_focal, int32_t opt_pp, int32_t opt_dist);

#ifdef __cplusplus
}
#endif
#endif /* EZCALIB_B200_H_ */
