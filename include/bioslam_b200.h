/*
 * bioslam_b200.h -- C ABI of the B200-native Levenberg-Marquardt path for bioslam factor graphs.
 *
 * This is the drop-in boundary.  The reference has no FFI: the seam it replaces is
 * the call sequence
 *     gtsam::LevenbergMarquardtOptimizer optimizer(m_graph, m_initialValues, params);
 *     optimizer.iterate(); optimizer.error(); optimizer.iterations(); optimizer.values();
 * at reference src/lowerBodyPoseEstimator.cpp:644-698 and src/imuPoseEstimator.cpp:324-386.
 * A host adapter (bioslam_b200/host, or a GTSAM-side adapter described in INTEGRATION.md) flattens
 * m_graph / m_initialValues into a b200_problem_desc and then drives the entry points below.
 *
 * Conventions (GTSAM 4.2 semantics):
 *   - all arithmetic is fp64;
 *   - Pose3 value = rotation matrix R (row-major, 9 doubles, body->nav) followed by translation t (3);
 *     tangent = [rot(3), trans(3)], retraction T * Expmap(xi) (full SE(3) exponential);
 *   - Vec3 / Point3 value = 3 doubles, additive;
 *   - Bias6 (imuBias::ConstantBias) value = [acc(3), gyro(3)], additive;
 *   - Unit3 value = unit 3-vector, tangent dim 2 through the GTSAM basis B(p);
 *   - factor error = 0.5*||whitened residual||^2, graph error = sum over factors.
 *
 * No torch / C++ types cross this ABI: plain pointers, sizes and POD structs only.  Host pointers are borrowed
 * for the duration of a call.  Functions return a b200_status; nothing throws.
 */
#ifndef BIOSLAM_B200_H
#define BIOSLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- status codes */
typedef enum {
    B200_OK = 0,
    B200_BAD_ARG = 1,
    B200_CUDA_ERROR = 2,
    B200_INDETERMINATE_SYSTEM = 3, /* non-positive Cholesky pivot == gtsam::IndeterminantLinearSystemException
                                      (reference src/imuPoseEstimator.cpp:360) */
    B200_LAMBDA_MAXED = 4,         /* LM "giving up": lambda >= lambdaUpperBound, iterations() did not advance
                                      (reference src/imuPoseEstimator.cpp:344-348) */
    B200_NOT_IMPLEMENTED = 5,
    B200_NCCL_ERROR = 6
} b200_status;

/* ---------------------------------------------------------------- variable types */
typedef enum {
    B200_VAR_POSE3 = 0, /* value 12, tangent 6 */
    B200_VAR_VEC3 = 1,  /* value 3,  tangent 3  (gtsam::Vector3 / gtsam::Point3) */
    B200_VAR_BIAS6 = 2, /* value 6,  tangent 6  (gtsam::imuBias::ConstantBias, [acc, gyro]) */
    B200_VAR_UNIT3 = 3  /* value 3,  tangent 2  (gtsam::Unit3) */
} b200_var_type;

/* ---------------------------------------------------------------- factor types
 * Variable order of every factor is the key order of the reference constructor (wrap/bioslam.h:21-105).
 * "params" are slot-level constants (same for every keyframe), "consts" are per-keyframe constants. */
typedef enum {
    /* gtsam::CombinedImuFactor (reference src/imuPoseEstimator.cpp:215).
     * vars: pose_i, vel_i, pose_j, vel_j, bias_i, bias_j.   params[0..2] = n_gravity.
     * consts (190): zeta[9]=[theta,p,v], H_biasAcc[9x3 row-major], H_biasOmega[9x3], biasHat[6], deltaTij,
     *               sqrt-information R (15x15 upper triangle, packed row-major: row r holds cols r..14) */
    B200_F_IMU_COMBINED = 0,
    /* gtsam::ImuFactor, static bias (reference src/imuPoseEstimator.cpp:298).
     * vars: pose_i, vel_i, pose_j, vel_j, bias.  params[0..2] = n_gravity.
     * consts (115): zeta[9], H_biasAcc[27], H_biasOmega[27], biasHat[6], deltaTij, R (9x9 upper packed, 45) */
    B200_F_IMU_STATICBIAS = 1,
    /* bioslam::AngularVelocityFactor (src/factors/AngularVelocityFactor.cpp:11-33).
     * vars: angvel, bias.  params[0..2]=sigmas.  consts (3): measured gyro */
    B200_F_ANGVEL = 2,
    /* bioslam::HingeJointConstraintVecErrEstAngVel (src/factors/HingeJointFactors.cpp:14-109).
     * vars: poseA, omegaA, poseB, omegaB, axisA(Unit3).  params[0..2]=sigmas */
    B200_F_HINGE_VEC = 3,
    /* bioslam::HingeJointConstraintNormErrEstAngVel (src/factors/HingeJointFactors.cpp:118-173). params[0]=sigma */
    B200_F_HINGE_NORM = 4,
    /* bioslam::ConstrainedJointCenterPositionFactor (src/factors/ConstrainedJointCenterPositionFactor.cpp:12-62).
     * vars: poseA, poseB, sA, sB.  params[0..2]=sigmas */
    B200_F_JOINTPOS_VEC = 5,
    /* bioslam::ConstrainedJointCenterNormPositionFactor (:65-123). params[0]=sigma, params[1]=expectedDist */
    B200_F_JOINTPOS_NORM = 6,
    /* bioslam::ConstrainedJointCenterVelocityFactor (src/factors/ConstrainedJointCenterVelocityFactor.cpp:13-131).
     * vars: poseA, velA, omegaA, sA, poseB, velB, omegaB, sB.  params[0..2]=sigmas */
    B200_F_JOINTVEL_VEC = 7,
    /* bioslam::SegmentLengthMagnitudeFactor (src/factors/SegmentLengthMagnitudeFactor.cpp:12-65).
     * vars: v1, v2.  params[0]=sigma, params[1]=ideal length */
    B200_F_SEGLEN_MAG = 8,
    /* bioslam::SegmentLengthMaxMagnitudeFactor (:68-106). params[0]=sigma, [1]=max length, [2]=a */
    B200_F_SEGLEN_MAX = 9,
    /* bioslam::SegmentLengthMinMagnitudeFactor (:108-144). params[0]=sigma, [1]=min length, [2]=a */
    B200_F_SEGLEN_MIN = 10,
    /* bioslam::SegmentLengthDiscrepancyFactor (src/factors/SegmentLengthDiscrepancyFactor.cpp:12-72).
     * vars: aProx, aDist, bProx, bDist.  params[0]=sigma */
    B200_F_SEGLEN_DISCREPANCY = 11,
    /* bioslam::AngleBetweenAxisAndSegmentFactor (src/factors/AngleBetweenAxisAndSegmentFactor.cpp:13-92).
     * vars: axis(Unit3), v1, v2.  params[0]=sigma, params[1]=expected angle */
    B200_F_ANGLE_AXIS_SEG = 12,
    /* bioslam::Pose3TranslationPrior (src/factors/Pose3Priors.cpp:11-32). vars: pose. params[0..2]=sigmas, [3..5]=prior */
    B200_F_POSE3_TRANSLATION_PRIOR = 13,
    /* bioslam::Pose3CompassPrior (src/factors/Pose3Priors.cpp:35-86). vars: pose.
     * params[0]=sigma, [1]=expected angle, [2..4]=refVecLocal, [5..7]=refVecNav, [8..10]=upVecNav */
    B200_F_POSE3_COMPASS_PRIOR = 14,
    /* gtsam::PriorFactor<Vector3|Point3>. params[0..2]=sigmas, [3..5]=mean */
    B200_F_PRIOR_VEC3 = 15,
    /* gtsam::PriorFactor<imuBias::ConstantBias>. params[0..5]=sigmas, [6..11]=mean */
    B200_F_PRIOR_BIAS6 = 16,
    /* gtsam::PriorFactor<Unit3> (Jacobian approximated by I as in GTSAM). params[0..1]=sigmas, [2..4]=mean */
    B200_F_PRIOR_UNIT3 = 17,
    /* gtsam::PriorFactor<Pose3> (Jacobian approximated by I). params[0..5]=sigmas, [6..17]=mean (R row-major, t) */
    B200_F_PRIOR_POSE3 = 18,
    /* bioslam::MagPose3Factor (src/factors/MagPose3Factor.cpp:12-38). vars: pose.
     * params[0..2]=sigmas, [3..5]=B_global (scale*direction + bias).  consts (3): measured field in body frame */
    B200_F_MAG_POSE3 = 19,
    /* bioslam::ConstrainedJointCenterNormVelocityFactor (src/factors/ConstrainedJointCenterVelocityFactor.cpp:137-214): the norm of the
     * vector error of B200_F_JOINTVEL_VEC (gtsam::norm3 chain rule).  Same variables.  params[0]=sigma */
    B200_F_JOINTVEL_NORM = 20,
    B200_F_NUM_TYPES = 21
} b200_factor_type;

#define B200_MAX_FACTOR_VARS 8
#define B200_MAX_FACTOR_PARAMS 24
#define B200_MAX_FACTOR_ROWS 15
#define B200_MAX_FACTOR_COLS 30

/* where a factor variable lives */
typedef enum {
    B200_AT_K = 0,      /* time block of keyframe k */
    B200_AT_K1 = 1,     /* time block of keyframe k+1 */
    B200_AT_STATIC = 2  /* static block */
} b200_var_kind;

/* One factor "slot": the same factor template instantiated at every keyframe k in [k_begin, k_end).
 * A slot whose variables are all B200_AT_STATIC is a static factor and must have k_begin=0, k_end=1. */
typedef struct {
    int32_t type;                             /* b200_factor_type */
    int32_t k_begin, k_end;
    int32_t nvars;
    int32_t var_kind[B200_MAX_FACTOR_VARS];   /* b200_var_kind */
    int32_t var_index[B200_MAX_FACTOR_VARS];  /* index into time_var_type[] or static_var_type[] */
    int32_t const_offset;                     /* offset (in doubles) of this slot's per-keyframe constants inside the
                                                 keyframe constant record; -1 when the type has none */
    int32_t reserved;
    double params[B200_MAX_FACTOR_PARAMS];
} b200_factor_slot;

/* Flat description of a time-chain factor graph: K keyframes that all carry the same "time block" of variables,
 * plus one block of static variables.  Dense integer indices replace gtsam::Symbol keys. */
typedef struct {
    int32_t K;                      /* number of keyframes */
    int32_t n_time_vars;            /* variables per keyframe */
    const int32_t* time_var_type;   /* [n_time_vars] b200_var_type */
    int32_t n_static_vars;
    const int32_t* static_var_type; /* [n_static_vars] */
    int32_t n_slots;
    const b200_factor_slot* slots;  /* [n_slots] */
    int32_t const_stride;           /* doubles per keyframe constant record */
    const double* consts;           /* [K][const_stride] (host pointer) */
} b200_problem_desc;

/* LevenbergMarquardtParams subset that bioslam sets (reference src/lowerBodyPoseEstimator.cpp:644-647,
 * src/imuPoseEstimator.cpp:324-325); remaining GTSAM defaults are fixed: minModelFidelity=1e-3,
 * diagonalDamping=false, useFixedLambdaFactor=true. */
typedef struct {
    double lambda_initial;     /* GTSAM default 1e-5 */
    double lambda_factor;      /* 10 */
    double lambda_lower_bound; /* GTSAM default 0 */
    double lambda_upper_bound; /* GTSAM default 1e5 */
    double relative_error_tol; /* used only by tryLambda's stopSearchingLambda test */
    double absolute_error_tol; /* unused by iterate(); kept for completeness */
    double min_model_fidelity; /* 1e-3 */
    int32_t gauss_newton;      /* 1: lambda == 0 and every step accepted (matlab/src/lowerBodyPoseEstimator.m:1108) */
    int32_t reserved;
} b200_lm_params;

/* state after one optimizer.iterate() */
typedef struct {
    double error;           /* optimizer.error() */
    double lambda;          /* optimizer.lambda() */
    int32_t iterations;     /* optimizer.iterations(): accepted outer iterations so far */
    int32_t inner_trials;   /* lambda trials (solve+retract+error) spent inside this iterate() */
    int32_t total_trials;   /* cumulative */
    int32_t status;         /* b200_status of this iterate(): OK, LAMBDA_MAXED or INDETERMINATE_SYSTEM */
} b200_lm_state;

/* bioslam outer loop (reference src/lowerBodyPoseEstimator.cpp:651-667 when consec_small_iter>1,
 * src/imuPoseEstimator.cpp:338-359 when single_imu_rule=1) */
typedef struct {
    double abs_err_decrease_limit;
    double rel_err_decrease_limit;
    double abs_err_limit;
    int32_t max_iterations;
    int32_t converge_at_consec_small_iter; /* lowerBody: 6 */
    int32_t single_imu_rule;               /* 1: imuPoseEstimator::fastOptimize exit rule */
    int32_t max_history;                   /* capacity of error_history (entries) */
} b200_outer_params;

typedef struct {
    double initial_error;
    double final_error;
    double final_lambda;
    int32_t iterations;    /* optimizer.iterations() at exit */
    int32_t outer_calls;   /* number of iterate() calls */
    int32_t total_trials;
    int32_t n_history;     /* entries written to error_history (== m_optimizationTotalError) */
    int32_t status;
    int32_t reserved;
    double device_ms;      /* device time of the loop, CUDA events */
} b200_optimize_result;

/* IMU noise model (reference include/imu/imuNoiseModelHandler.h:19-35, continuous-time sigmas) */
typedef struct {
    double accel_noise_sigma;
    double accel_bias_rw_sigma;
    double gyro_noise_sigma;
    double gyro_bias_rw_sigma;
    double preint_bias_sigma;
    double integration_err_sigma;
} b200_imu_noise;

typedef struct b200_problem b200_problem;

/* ---------------------------------------------------------------- problem lifetime */
/* Upload the graph description (copied; desc may be freed afterwards).  device = CUDA ordinal. */
b200_status b200_problem_create(const b200_problem_desc* desc, int32_t device, b200_problem** out);
void b200_problem_destroy(b200_problem* p);
const char* b200_last_error(void);

/* dims implied by a description */
int32_t b200_time_value_dim(const b200_problem* p);    /* doubles per keyframe in the value array */
int32_t b200_static_value_dim(const b200_problem* p);
int32_t b200_time_tangent_dim(const b200_problem* p);  /* e.g. 126 */
int32_t b200_static_tangent_dim(const b200_problem* p); /* e.g. 58 */

/* ---------------------------------------------------------------- values (== gtsam::Values in / out) */
/* time_values[K][time_value_dim], static_values[static_value_dim]; variables packed in table order. */
b200_status b200_set_values(b200_problem* p, const double* time_values, const double* static_values);
b200_status b200_get_values(b200_problem* p, double* time_values, double* static_values);

/* ---------------------------------------------------------------- optimizer */
/* == LevenbergMarquardtOptimizer ctor: resets lambda/iterations and computes the initial error */
b200_status b200_lm_begin(b200_problem* p, const b200_lm_params* params, double* initial_error);
/* == optimizer.iterate() */
b200_status b200_lm_iterate(b200_problem* p, b200_lm_state* state);
/* == graph.error(values) at the current values */
b200_status b200_error(b200_problem* p, double* error);
/* bioslam's fastOptimize outer loop run against device-resident values; error_history may be NULL */
b200_status b200_optimize(b200_problem* p, const b200_lm_params* lm, const b200_outer_params* outer,
                          double* error_history, b200_optimize_result* result);

/* ---------------------------------------------------------------- preintegration (A0)
 * GTSAM-equivalent PreintegratedCombinedMeasurements (tangent preintegration) on device.
 * acc/gyro: [n_imu][n_samples][3] (host), bias_hat: [n_imu][6] = [acc,gyro].
 * Interval i of an IMU integrates samples [i*n_per, (i+1)*n_per) (reference src/imuPoseEstimator.cpp:200-213).
 * out: [n_imu][n_intervals][190] records laid out as B200_F_IMU_COMBINED consts (host). */
b200_status b200_preintegrate_combined(int32_t device, int32_t n_imu, int32_t n_samples, int32_t n_per,
                                       int32_t n_intervals, const double* acc, const double* gyro, double dt,
                                       const double* bias_hat, const b200_imu_noise* noise, double* out);

/* static-bias twin (gtsam::ImuFactor / PreintegratedImuMeasurements, reference src/imuPoseEstimator.cpp:271-296):
 * combined=0 -> 115-double records laid out as B200_F_IMU_STATICBIAS consts; combined=1 == b200_preintegrate_combined */
b200_status b200_preintegrate(int32_t device, int32_t n_imu, int32_t n_samples, int32_t n_per, int32_t n_intervals,
                              const double* acc, const double* gyro, double dt, const double* bias_hat,
                              const b200_imu_noise* noise, int32_t combined, double* out);

/* ---------------------------------------------------------------- derived joint angles
 * == lowerBodyPoseEstimator::setDerivedJointAnglesFromEstimatedStates (reference src/lowerBodyPoseEstimator.cpp:813-835):
 * segment frames from the estimated static vectors / axes (bioutils::get_R_SacrumImu_to_Pelvis, get_R_Segment_to_N,
 * reference src/bioutils.cpp:12-115) and clinical Grood-Suntay angles (bioutils::clinical3DofJcsAngles, :117-237), computed
 * on the device from the problem's CURRENT values, one thread per keyframe.
 * pose_var: time-variable indices of the Pose3 of the sacrum, right thigh, right shank, left thigh, left shank IMU.
 * static_var: static-variable indices of sacrumImuToRHipCtr, sacrumImuToLHipCtr, rkneeAxisThigh, rthighImuToHipCtr,
 *   rthighImuToKneeCtr, rkneeAxisShank, rshankImuToKneeCtr, rshankImuToAnkleCtr, lkneeAxisThigh, lthighImuToHipCtr,
 *   lthighImuToKneeCtr, lkneeAxisShank, lshankImuToKneeCtr, lshankImuToAnkleCtr.
 * angles: [K][12] = right knee, left knee, right hip, left hip x (flexion/extension, ab/adduction, int/ext rotation), rad. */
typedef struct b200_joint_angle_desc {
    int32_t pose_var[5];
    int32_t static_var[14];
} b200_joint_angle_desc;
b200_status b200_lower_body_joint_angles(b200_problem* p, const b200_joint_angle_desc* d, double* angles);

/* ---------------------------------------------------------------- solver partition
 * The time chain is cut into n_chunks contiguous windows that are eliminated in parallel (one CTA each) and joined through
 * a reduced system on the separator keyframes.  Default: ~sqrt(2.2 K), at most one per SM; 1 = plain sequential sweep. */
b200_status b200_set_chunks(b200_problem* p, int32_t n_chunks);
int32_t b200_get_chunks(const b200_problem* p);
/* nested partition: n_chunks chunks of keyframes; their n_chunks-1 separators are cut into n_groups groups that are
 * eliminated in parallel as well, and only the n_groups-1 group separators are swept sequentially.  0 = choose from the
 * built-in cost model, n_groups = 1 = two-level scheme. */
b200_status b200_set_partition(b200_problem* p, int32_t n_chunks, int32_t n_groups);
int32_t b200_get_groups(const b200_problem* p);
/* depth of the nested partition (0 = plain sequential sweep).  With n_groups = 0 on a single GPU the separators are
 * partitioned recursively (as many levels as the cost model finds worthwhile); sizes[l] receives the chunk count of level l
 * (up to max_levels entries, may be NULL). */
int32_t b200_get_levels(const b200_problem* p, int32_t* sizes, int32_t max_levels);

/* ---------------------------------------------------------------- multi-GPU (time axis sharded, section 8(e))
 * One process per GPU, every rank creates the problem with the FULL description and the same values.  Rank r owns a
 * contiguous range of chunks (== a window of keyframes): it linearizes and eliminates only that window.  The two
 * exchange steps are delegated to host-supplied collectives operating on DEVICE pointers and ordered on the given CUDA
 * stream (torch.distributed / NCCL in the Python host, ncclAllGather / ncclAllReduce in a C++ host):
 *   allgather: in place; buf holds world_size segments of bytes_per_rank, segment `rank` is this rank's contribution.
 *   allreduce: in-place sum of n doubles.
 * Both return 0 on success. */
typedef int (*b200_allgather_fn)(void* ctx, void* dev_buf, uint64_t bytes_per_rank, void* cuda_stream);
typedef int (*b200_allreduce_fn)(void* ctx, double* dev_buf, int32_t n, void* cuda_stream);
b200_status b200_comm_set(b200_problem* p, int32_t world_size, int32_t rank, b200_allgather_fn allgather,
                          b200_allreduce_fn allreduce, void* ctx);

/* NCCL inside the library (the production multi-GPU path; the callbacks above stay for hosts that own their communicators and for
 * the gloo tests).  Rank 0 calls b200_nccl_unique_id and distributes the 128 bytes; every rank calls b200_nccl_init with the
 * total rank count, its global rank and the number of lambda groups (see b200_spec_set; ranks are group-major).  libnccl.so.2 is
 * resolved at run time (dlopen), so the library itself has no link-time dependency on NCCL.  The exchanges of one solve are
 * issued as ncclGroupStart/End groups on the problem's stream: {window-boundary Schur complement + fill}, {level-1 outputs +
 * group-separator blocks}, {step}, plus one all-reduce of the cost scalars per lambda trial. */
b200_status b200_nccl_unique_id(char out[128]);
b200_status b200_nccl_init(b200_problem* p, int32_t n_ranks, int32_t global_rank, const char id_bytes[128], int32_t n_groups);
/* device time (ms) of the collectives and number of collective groups since the last reset; the time is measured only while
 * "phase_timing" is on (every group is then bracketed by CUDA events and waited for) */
b200_status b200_comm_ms(b200_problem* p, double* ms, int64_t* calls, int32_t reset);

/* Lambda speculation (optional, on top of the sharding above).  GTSAM's tryLambda loop (reference seam:
 * src/lowerBodyPoseEstimator.cpp:659 `optimizer.iterate()`) re-solves the SAME linearization with lambda, 10 lambda, ... until
 * a step is accepted; the rungs of that ladder are independent of one another.  With n_groups groups of GPUs, group g solves
 * rung g of the ladder at the same time and the verdicts are then read in ladder order, so the accepted step, the final
 * lambda and the trial count are exactly those of the sequential loop (bit-identical when world_size == 1).
 * Call after b200_comm_set (world_size = ranks per group); `allgather` spans all n_groups * world_size ranks in group-major
 * order (global rank = group * world_size + rank). */
b200_status b200_spec_set(b200_problem* p, int32_t n_groups, int32_t group, b200_allgather_fn allgather, void* ctx);

/* ---------------------------------------------------------------- parity/debug accessors (used by tests only)
 * External tangent order == variable table order (the library permutes internally). */
b200_status b200_debug_linearize(b200_problem* p);
b200_status b200_debug_get_gradient(b200_problem* p, double* g_time /*[K][nt]*/, double* g_static /*[ns]*/);
/* dense blocks of J^T J at keyframe k: D[nt*nt] (k,k), O[nt*nt] (k,k+1) (zeros for k=K-1), A[nt*ns] (k,static) */
b200_status b200_debug_get_hessian_blocks(b200_problem* p, int32_t k, double* D, double* O, double* A);
b200_status b200_debug_get_static_hessian(b200_problem* p, double* S /*[ns*ns]*/);
/* solve (J^T J + lambda I) delta = -J^T r on the last linearization */
b200_status b200_debug_solve(b200_problem* p, double lambda, double* delta_time /*[K][nt]*/, double* delta_static);
/* reduced blocks of keyframe k after the elimination of the keyframe-local dofs (angular velocities) at damping lambda -- the
 * Schur complement the chain solver works on.  dims[4] = {nl, nc, ns1, used}; chain_ext[nc] = external tangent index of every
 * chain dof; Dc[nc*nc] (symmetric), Ac[ns1*nc] (static rows, last row = -gradient), Sc[ns1*ns1].  Call with NULL outputs to
 * query the dimensions. */
b200_status b200_debug_local_elim(b200_problem* p, int32_t k, double lambda, int32_t dims[4], int32_t* chain_ext, double* Dc, double* Ac, double* Sc);
/* raw read of an internal solver buffer (development only; layouts in DESIGN.md section 2) */
b200_status b200_debug_read(b200_problem* p, const char* name, int64_t offset, int64_t n, double* out);
/* one factor instance of `type` evaluated on the device by the same device functions linearize_kernel uses: residual r[rows]
 * and Jacobian J[rows*cols] (row-major; whitened if whiten != 0).  params: B200_MAX_FACTOR_PARAMS doubles; consts: the factor's
 * per-keyframe constant record (or NULL); xcat: the variables' values concatenated in factor order.  Replaces, for tests, the
 * reference's per-factor evaluateError(...) call sites (src/factors/ all .cpp files). */
b200_status b200_debug_factor_eval(int32_t device, int32_t type, const double* params, const double* consts, const double* xcat,
                                   int32_t whiten, double* r, double* J);
/* scalars of the last lambda trial: out[0] = linear.error(0) (== the cost at the linearization point, all ranks summed),
 * out[1] = linear.error(delta), out[2] = cost of the retracted values */
b200_status b200_debug_get_scalars(b200_problem* p, double out[4]);
/* kernel launch counter (all launches issued by this library for p since creation) */
int64_t b200_launch_count(const b200_problem* p);
/* device time (CUDA events on the problem's stream) since the last reset.  out[0] linearize+assemble ms, out[1] solve ms,
 * out[2] model-error+retract+error ms, out[3] chain_sweep_kernel ms, out[4] chain_sweep_kernel launches,
 * out[5] top_solve_kernel ms, out[6] top_solve_kernel launches, out[7] chain_backsub_kernel ms, out[8] local elimination ms,
 * out[9] level-0 Schur complement + window-boundary exchange + reduce ms, out[10] level 1 (rank-aligned separators) ms,
 * out[11] levels >= 2 (computed redundantly on every rank of a time-sharded group) ms.
 * The first call switches the per-phase events on. */
b200_status b200_phase_ms(b200_problem* p, double out[12], int32_t reset);
/* runtime switches of the LM driver (b200_lm_iterate): "graph" 0/1 -- every lambda trial is one instantiated CUDA graph (default
 * 1, single rank); "spec_trials" 1..4 -- lambda trials queued on the device per host synchronisation (default 2: the accept /
 * reject / damping decisions are taken on the device by lm_judge_kernel, trials queued behind an accepted one return at once);
 * "phase_timing" 0/1 -- per-phase CUDA events (direct launches instead of graphs). */
b200_status b200_set_option(b200_problem* p, const char* name, int32_t value);
/* host <-> device synchronisations issued by b200_lm_iterate since creation (1 per iterate() when the ladder ends inside the
 * queued trials) */
int64_t b200_host_sync_count(const b200_problem* p);
/* CUDA-event stopwatch on the problem's stream (the stream every kernel of this problem is launched on) */
b200_status b200_timer_begin(b200_problem* p);
b200_status b200_timer_end(b200_problem* p, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* BIOSLAM_B200_H */
