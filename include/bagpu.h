/*
 * libbagpu -- B200-native sparse bundle adjustment behind ORB-SLAM3's Optimizer.
 *
 * C ABI (plain pointers and sizes; no C++ types, no exceptions, no torch types).
 * This is the drop-in boundary for the reference's BA entry points
 * (reference paths are relative to the ORB_SLAM3 tree):
 *
 *   Optimizer::LocalBundleAdjustment(KeyFrame*,bool*,Map*,...)   src/Optimizer.cc:1116-1498
 *   Optimizer::LocalBundleAdjustment(KeyFrame*,vector,vector,bool*) src/Optimizer.cc:3506-3953
 *   Optimizer::BundleAdjustment / GlobalBundleAdjustemnt          src/Optimizer.cc:53-390
 *   Optimizer::PoseOptimization(Frame*)                           src/Optimizer.cc:815-1114
 *
 * The replacement bodies of those functions gather KeyFrame/MapPoint/Frame data
 * into the flat arrays below, make ONE call, and classify / erase / scatter from
 * the returned per-edge chi2 and depth flags exactly as the reference does
 * (see INTEGRATION.md for the binding a maintainer adds to Optimizer.cc).
 *
 * Everything between "graph built" and "estimates + per-edge chi2 returned" --
 * i.e. the whole of Thirdparty/g2o (SparseOptimizer, OptimizationAlgorithmLevenberg,
 * BlockSolver_6_3, LinearSolverEigen/Dense), OptimizableTypes.cpp, the stereo
 * edges of types_six_dof_expmap.cpp and CameraModels project/projectJac --
 * happens on the GPU (sm_100a). There is no CPU fallback.
 */
#ifndef BAGPU_H
#define BAGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAGPU_VERSION 2

/* ---- status codes (SURVEY 8b "Errors") ---------------------------------- */
#define BAGPU_OK                 0
#define BAGPU_TERMINATE_TRIALS   1  /* g2o Terminate: 10 failed trials or rho==0  (optimization_algorithm_levenberg.cpp:151-155) */
#define BAGPU_TERMINATE_NBAD     2  /* g2o Terminate: ORB-SLAM "_nBad>=3" rule      (optimization_algorithm_levenberg.cpp:157-166) */
#define BAGPU_STOPPED            3  /* caller's stop flag seen                      (sparse_optimizer.cpp:376, sparse_optimizer.h:188) */
#define BAGPU_ERR_CUDA          -1
#define BAGPU_ERR_ARG           -2
#define BAGPU_ERR_NCCL          -3
#define BAGPU_ERR_NO_DEVICE     -4
#define BAGPU_ERR_ALLOC         -5

/* ---- camera models (include/CameraModels/GeometricCamera.h:97-101) ------ */
#define BAGPU_CAM_PINHOLE 0         /* src/CameraModels/Pinhole.cpp:35-41,71-81       */
#define BAGPU_CAM_KB8     1         /* src/CameraModels/KannalaBrandt8.cpp:46-65,145-175 */

typedef struct bagpu_camera {
    int32_t type;                   /* BAGPU_CAM_*                                       */
    float   p[8];                   /* fx fy cx cy [k0 k1 k2 k3]  (std::vector<float> mvParameters) */
    float   bf;                     /* KeyFrame::mbf, used by stereo edges only           */
} bagpu_camera;

/* Right-camera extrinsics Trl for "ToBody" edges (include/OptimizableTypes.h:143). */
typedef struct bagpu_rig {
    double qt[7];                   /* tx ty tz qx qy qz qw (g2o SE3Quat::toVector order, se3quat.h:138-148) */
} bagpu_rig;

/* ---- edge kinds --------------------------------------------------------- */
#define BAGPU_EDGE_MONO   0  /* ORB_SLAM3::EdgeSE3ProjectXYZ(OnlyPose)        OptimizableTypes.h:31-55,89-115  */
#define BAGPU_EDGE_STEREO 1  /* g2o::EdgeStereoSE3ProjectXYZ(OnlyPose)        types_six_dof_expmap.h:146-175,208-236 */
#define BAGPU_EDGE_BODY   2  /* ORB_SLAM3::EdgeSE3ProjectXYZ(OnlyPose)ToBody  OptimizableTypes.h:57-87,117-144 */

/* obs_flags bits */
#define BAGPU_FLAG_ROBUST 1  /* edge carries a RobustKernelHuber when the call starts */

/*
 * A BA problem = what Optimizer.cc puts into g2o::SparseOptimizer.
 * Poses are VertexSE3Expmap in ascending vertex-id order, points are
 * VertexSBAPointXYZ (marginalized) in ascending id order, observations are the
 * edges in insertion order (g2o's activeEdges order, sparse_optimizer.cpp:482-487).
 * The caller owns every pointer for the duration of the call.
 */
typedef struct bagpu_problem {
    int32_t        n_poses;
    const double  *pose_qt;         /* [n_poses][7] tx ty tz qx qy qz qw; normalised on entry like SE3Quat(q,t) (se3quat.h:62-64) */
    const uint8_t *pose_fixed;      /* [n_poses] setFixed()                               */
    int32_t        n_points;
    const double  *points;          /* [n_points][3]                                      */
    int32_t        n_cameras;
    const bagpu_camera *cameras;
    int32_t        n_rigs;
    const bagpu_rig    *rigs;       /* may be NULL when n_rigs==0                         */
    int64_t        n_obs;
    const int32_t *obs_pose;        /* [n_obs] index into poses                            */
    const int32_t *obs_point;       /* [n_obs] index into points                           */
    const int16_t *obs_cam;         /* [n_obs] index into cameras                          */
    const int16_t *obs_rig;         /* [n_obs] index into rigs, -1 unless kind==BODY      */
    const uint8_t *obs_kind;        /* [n_obs] BAGPU_EDGE_*                                */
    const uint8_t *obs_flags;       /* [n_obs] BAGPU_FLAG_*                                */
    const double  *obs_u;           /* [n_obs] measurement (float pixel widened)           */
    const double  *obs_v;
    const double  *obs_ur;          /* [n_obs] right-image u, read for STEREO only; may be NULL if no stereo edge */
    const double  *obs_inv_sigma2;  /* [n_obs] information = inv_sigma2 * I                */
} bagpu_problem;

/* ---- schedule: the sequence of optimize() calls and gates --------------- */
#define BAGPU_GATE_NONE 0
#define BAGPU_GATE_LBA  1  /* chi2 > gate || !isDepthPositive -> level 1, doubles (Optimizer.cc:3745-3776) */
#define BAGPU_GATE_POSE 2  /* recompute outliers, (float)chi2 > (float)gate -> level 1 else level 0 (Optimizer.cc:1013-1100) */

typedef struct bagpu_round {
    int32_t iterations;             /* optimizer.optimize(iterations)                      */
    int32_t gate_after;             /* BAGPU_GATE_*  applied after this round              */
    double  gate_mono;              /* 5.991  (body edges use gate_mono)                   */
    double  gate_stereo;            /* 7.815                                               */
    int32_t drop_kernel_after;      /* 1: setRobustKernel(0) on every edge after the gate  */
    int32_t reset_pose;             /* 1: reset poses to the call's initial poses before this round (Optimizer.cc:1008-1009) */
} bagpu_round;

typedef struct bagpu_schedule {
    int32_t            n_rounds;
    const bagpu_round *rounds;
    double  delta_mono;             /* Huber delta as the reference stores it: (double)(float)sqrt(5.991) etc. */
    double  delta_stereo;
    double  lambda_init;            /* <=0: tau*max diag (tau=1e-5); >0: setUserLambdaInit  (Optimizer.cc:1197-1198) */
    const volatile uint8_t *stop_flag; /* the caller's bool* pbStopFlag, may be NULL; polled where g2o polls terminate() (small maps run the LM
                                        * loop chained on the device and poll it every four trials) */
    int32_t linear_solver;          /* BAGPU_SOLVER_* */
    int32_t max_trace;              /* capacity of bagpu_result.trace (entries)            */
    double  pcg_tolerance;          /* BAGPU_SOLVER_PCG: stop at |r| <= tol |b| (<= 0: 1e-10) */
    int32_t pcg_max_iterations;     /* BAGPU_SOLVER_PCG: give up (= failed solve, rejected trial) after this many (<= 0: 20000) */
} bagpu_schedule;

/* Reduced-camera-system solver (the seam of g2o::LinearSolver::solve, core/linear_solver.h:50-58).
 * AUTO = CHOLESKY: the envelope (band) Cholesky family -- one front, two fronts from both ends, or P fronts with spikes and a
 * cyclic-reduced separator system -- picked from the system's size; it is exact like the reference's SimplicialLDLT and, on the
 * keyframe chains of this path, faster than PCG at every BASELINE size (DESIGN.md section 4 has the measurements).
 * PCG = block-Jacobi preconditioned conjugate gradients on the band-stored Hschur, an opt-in alternative. */
#define BAGPU_SOLVER_AUTO     0
#define BAGPU_SOLVER_CHOLESKY 1
#define BAGPU_SOLVER_PCG      2

/* One entry per LM iteration = one OptimizationAlgorithmLevenberg::solve call
 * (this is also g2o's G2OBatchStatistics view, core/batch_stats.h:40-62). */
typedef struct bagpu_trace {
    int32_t round;
    int32_t iteration;
    double  chi2_before;            /* activeRobustChi2 at the linearisation point          */
    double  chi2_after;             /* currentChi when solve() returns                      */
    double  lambda;                 /* _currentLambda when solve() returns                  */
    int32_t trials;                 /* _levenbergIterations (qmax)                          */
    int32_t status;                 /* BAGPU_OK / BAGPU_TERMINATE_*                         */
    /* per-iteration phase record (G2OBatchStatistics: numEdges, timeLinearize + timeQuadraticForm + timeSchurComplement,
     * timeLinearSolver, timeUpdate + timeResiduals, timeIteration); device times from CUDA events, summed over the trials */
    int64_t active_edges;           /* numEdges: level-0 edges of this round                */
    double  linearise_schur_us;     /* stage_kernel + tile_diag_kernel + pair_tile_mma_kernel (all trials of the iteration) */
    double  linear_solve_us;        /* reduced camera system; the two-front solver of a mid-sized map runs BESIDE the pass, so the phases overlap there */
    double  update_us;              /* back-substitution, update, evaluation                 */
    double  iteration_us;           /* host wall clock of the whole iteration                */
} bagpu_trace;

typedef struct bagpu_result {
    double  *pose_qt;               /* [n_poses][7] out                                     */
    double  *points;                /* [n_points][3] out                                    */
    double  *edge_chi2;             /* [n_obs] e->chi2() as the caller would read it after the last round */
    uint8_t *edge_depth_pos;        /* [n_obs] e->isDepthPositive() on the final estimates  */
    uint8_t *edge_level;            /* [n_obs] e->level() after the last gate               */
    bagpu_trace *trace;             /* [max_trace]                                          */
    int32_t  n_trace;
    int32_t  status;                /* status of the last optimize()                        */
} bagpu_result;

/* ---- batched PoseOptimization (Optimizer.cc:815-1114) ------------------- */
typedef struct bagpu_pose_batch {
    int32_t        n_frames;
    const double  *pose_qt;         /* [n_frames][7] initial Tcw                            */
    const int64_t *frame_ptr;       /* [n_frames+1] CSR offsets into the edge arrays        */
    int32_t        n_cameras;
    const bagpu_camera *cameras;
    int32_t        n_rigs;
    const bagpu_rig    *rigs;
    int64_t        n_obs;
    const double  *xw;              /* [n_obs][3] MapPoint world position (fixed)           */
    const int16_t *obs_cam;
    const int16_t *obs_rig;
    const uint8_t *obs_kind;
    const double  *obs_u, *obs_v, *obs_ur;
    const double  *obs_inv_sigma2;
    double delta_mono, delta_stereo;  /* (double)(float)sqrt(5.991), (double)(float)sqrt(7.815) */
    float  gate_mono, gate_stereo;    /* 5.991f, 7.815f                                     */
} bagpu_pose_batch;

typedef struct bagpu_pose_result {
    double  *pose_qt;               /* [n_frames][7]                                        */
    uint8_t *outlier;               /* [n_obs]  Frame::mvbOutlier                           */
    int32_t *n_inliers;             /* [n_frames] nInitialCorrespondences - nBad (0 if <3 correspondences) */
    double  *final_chi2;            /* [n_frames] currentChi of the last LM iteration run (diagnostic) */
} bagpu_pose_result;

/* ---- timings of the last call (CUDA events on the library's stream) ----- */
typedef struct bagpu_timing {
    double h2d_ms, solve_ms, d2h_ms;   /* whole phases                                      */
    double build_ms;                   /* sum over launches of the linearise+Schur kernel   */
    double linsolve_ms;                /* reduced camera system                             */
    double update_ms;                  /* back-substitution + update + evaluation kernel    */
    int64_t build_launches, update_launches, linsolve_launches, total_launches;
    int64_t lm_iterations, lm_trials;
    int64_t edge_linearisations;       /* active edges x linearise passes                   */
    int64_t edge_evaluations;          /* active edges x evaluation passes                  */
    int64_t h2d_bytes, d2h_bytes;
    int32_t pcg_iterations;            /* total, when the PCG path ran                      */
    int32_t schur_blocks;              /* upper-triangular 6x6 blocks of Hschur (inside the stored band) */
    int32_t solve_retries;             /* trials re-run because the overlapped solve starved (see INTEGRATION.md "Threading") */
    int32_t solver_parts;              /* partitions of the reduced-system factorisation (1 = one front, 2 = two-way, >2 = partitioned) */
} bagpu_timing;

typedef struct bagpu_ctx bagpu_ctx;

/* One context per calling thread (Tracking / LocalMapping / LoopClosing / GBA):
 * owns a stream and a grow-only device arena. device_id < 0 -> current device. */
int  bagpu_init(int device_id, bagpu_ctx **out);
void bagpu_destroy(bagpu_ctx *ctx);
const char *bagpu_strerror(int code);
const char *bagpu_last_error(const bagpu_ctx *ctx);   /* detail of the last negative status */

/* Multi-GPU global BA (one process per GPU): rank 0 creates the id, the host
 * plumbing (torch.distributed / MPI / anything) broadcasts its 128 bytes. */
int  bagpu_comm_unique_id(uint8_t id_out[128]);
int  bagpu_comm_init(bagpu_ctx *ctx, int world_size, int rank, const uint8_t id[128]);

/* Optional: page-lock caller-owned gather buffers so the H2D/D2H copies run at full PCIe rate and
 * asynchronously (an adapter that keeps persistent gather arrays registers them once). */
int  bagpu_pin_host(void *p, size_t bytes);
int  bagpu_unpin_host(void *p);

/* Whole call with HOST buffers: H2D, solve, D2H. */
int  bagpu_solve_ba(bagpu_ctx *ctx, const bagpu_problem *p, const bagpu_schedule *s, bagpu_result *r);

/* Split form: keep the problem resident in HBM (bench "value" path; also lets a
 * caller overlap gather of the next window with the solve of this one).
 * With a communicator, each rank uploads ITS landmark shard (all poses, its
 * points, all observations of its points); solve reduces the camera system. */
int  bagpu_upload(bagpu_ctx *ctx, const bagpu_problem *p);
int  bagpu_solve_resident(bagpu_ctx *ctx, const bagpu_schedule *s, bagpu_result *r /* may be NULL: no D2H */);
int  bagpu_download(bagpu_ctx *ctx, bagpu_result *r);
int  bagpu_reset_resident(bagpu_ctx *ctx);   /* estimates, edge levels and kernels back to the uploaded state */
/* New estimates for the resident map (same poses / points / observations, same fixed set): pose_qt [n_poses][7], points [n_points][3],
 * either may be NULL (kept). They become the "uploaded state" of bagpu_reset_resident. The plan of the upload (tile records, envelope,
 * solver fronts) is kept: a caller that re-optimises one window -- GlobalBundleAdjustemnt re-run after a loop correction moved the
 * keyframes (src/LoopClosing.cc:2289 after :2330-2460), or local BA twice on one window -- skips the observation upload and the
 * re-planning (first step: estimates only; adding / removing observations still takes a bagpu_upload). */
int  bagpu_update_estimates(bagpu_ctx *ctx, const double *pose_qt, const double *points);

int  bagpu_pose_opt_batch(bagpu_ctx *ctx, const bagpu_pose_batch *b, bagpu_pose_result *r);
/* Split form of the above for device-resident timing. */
int  bagpu_pose_upload(bagpu_ctx *ctx, const bagpu_pose_batch *b);
int  bagpu_pose_solve_resident(bagpu_ctx *ctx, bagpu_pose_result *r /* may be NULL */);

int  bagpu_get_timing(const bagpu_ctx *ctx, bagpu_timing *out);

/* Device self-test hooks used by tests (libm parity of the float fossils). */
int  bagpu_test_atan2f(bagpu_ctx *ctx, const float *y, const float *x, float *out, int64_t n);
/* (A + lambda I) x = b with the production solver; A dense symmetric row-major, col_end[j] = last nonzero row of column j
 * (monotone). fail_out = 1 when a pivot was not positive. */
int  bagpu_test_solve(bagpu_ctx *ctx, int n, const int *col_end, const double *A, const double *b, double lambda, double *x, int *fail_out);
/* The same through the partitioned solver with `parts` factorisation fronts (>= 3; long keyframe chains). */
int  bagpu_test_solve_parts(bagpu_ctx *ctx, int n, const int *col_end, const double *A, const double *b, double lambda, int parts, double *x, int *fail_out);

/* Measured FP64 throughput of the device (TFLOP/s): plain DFMA and the m8n8k4 FP64 MMA; the denominators of the FP64 roofline. */
int  bagpu_test_fp64_peak(bagpu_ctx *ctx, double *dfma_tflops, double *dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* BAGPU_H */
