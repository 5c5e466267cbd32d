/*
 * ckc.h -- C ABI of the B200-native batched state-validity library (libckc.so) for the closed-chain
 * dual ABB IRB120 + rod system of avishais/ABB_dualarm_planning.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point replaces one
 * member function of the reference's `StateValidityChecker` family (file:line cited per function, paths relative to
 * the reference root) with a BATCH call; the C++ classes in abb_dualarm_planning_b200/host/ keep the reference's
 * signatures on top of it.  All computation runs in hand-written sm_100a CUDA kernels; there is no CPU fallback:
 * every call fails with CKC_ERR_CUDA when no device/kernel image is usable.
 *
 * Conventions
 *   - configuration = 12 doubles [q1(6) | q2(6)], radians, the reference's `State` pair (retrieveStateVector,
 *     StateValidityCheckerPCS.cpp:23-32).  Host arrays are AoS: q[i*12 + j].
 *   - lengths in mm. env 1 = "3 poles" (D=900, L=300), env 2 = cones (D=1200, L=500)  (StateValidityCheckerPCS.h:23-26)
 *   - host entry points accept pageable or pinned host pointers; copies are issued on the context's stream and the
 *     call returns after the results are in the caller's buffers (synchronous, like the reference's calls).
 *   - `_dev` entry points take DEVICE pointers (structure of arrays: q[j*n + i]) and a cudaStream_t (as void*);
 *     they only enqueue work -- the caller synchronises.
 *   - return value: 0 = CKC_OK, negative = error (never throws, never exits).  Output flags are 0/1 bytes.
 *   - a context is bound to one CUDA device (ckc_create) or owns one sub-context per device (ckc_create_multi) and may be
 *     used from one host thread at a time.  Samples / edges are independent: no collective, no peer access.
 */
#ifndef CKC_H_
#define CKC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CKC_OK 0
#define CKC_ERR_ARG (-1)        /* bad argument */
#define CKC_ERR_IO (-2)         /* mesh file / directory unreadable (the reference exit(-1)s, collisionDetection.cpp:503) */
#define CKC_ERR_CUDA (-3)       /* CUDA failure (no device, no sm_100a image, out of memory, launch error) */
#define CKC_ERR_CAPACITY (-4)   /* internal work queue overflow that could not be recovered */

typedef struct ckc_ctx ckc_ctx;

typedef struct ckc_stats {          /* mirrors the reference's public counters (collisionDetection.h:30-37, apc_class.h:112-119) */
    int64_t ik_calls;               /* IK_counter: projection attempts */
    int64_t ik_feasible;            /* projections that succeeded */
    int64_t collision_calls;        /* collisionCheck_counter: collision_state evaluations */
    int64_t collisions;             /* evaluations that returned "in collision" */
    int64_t is_valid_calls;         /* isValid_counter: isValidRBS evaluations (RBS midpoints) */
    int64_t bv_tests;               /* bounding-volume pair tests executed by the kernels */
    int64_t tri_tests;              /* exact triangle-pair distance tests executed by the kernels */
    int64_t queue_overflows;        /* configurations re-run through the large-stack path */
    int64_t gd_iterations;          /* Newton iterations (GD flavour) */
    int64_t kernel_launches;        /* kernels launched by this context since creation */
} ckc_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------------- */
/* Replaces StateValidityChecker::StateValidityChecker(int env) (StateValidityCheckerPCS.h:44-51) + load_models()
 * (collisionDetection.cpp:496-963).  mesh_path: a directory holding the reference's .tris files (e.g. "./simulator")
 * or a packed .ckcmesh file (tools/pack_meshes.py).  device: CUDA ordinal. */
int  ckc_create(ckc_ctx** ctx, int env, const char* mesh_path, int device);
/* The same with ndev devices behind ONE context: every host entry point splits its batch into ndev contiguous
 * index ranges and drives each device from its own host thread (samples and edges are independent: no collective, no peer
 * access; meshes and hierarchies are replicated).  A planner process that owns one checker object thereby uses every GPU of the
 * box without a launcher.  Counters (ckc_get_stats) are summed over the devices; the `_dev` entry points need a single-device
 * context.  mesh limits: each mesh at most 1024 triangles (CKC_ERR_ARG otherwise; the reference's largest has 1000). */
int  ckc_create_multi(ckc_ctx** ctx, int env, const char* mesh_path, const int* devices, int ndev);
int  ckc_num_devices(const ckc_ctx* ctx);
void ckc_destroy(ckc_ctx* ctx);
const char* ckc_last_error(const ckc_ctx* ctx);     /* human readable text of the last failure on this context */
const char* ckc_version(void);
int  ckc_get_stats(ckc_ctx* ctx, ckc_stats* out);   /* LogPerf2file counters, StateValidityCheckerPCS.cpp:628-651 */
int  ckc_reset_stats(ckc_ctx* ctx);                 /* initiate_log_parameters, StateValidityCheckerPCS.h:186-202 */
int  ckc_num_pairs(const ckc_ctx* ctx);             /* 116 (env 1) / 103 (env 2) entries of the collision pair table */
/* pinned host buffers for callers that want full-rate PCIe copies (optional) */
int  ckc_host_alloc(void** p, uint64_t bytes);
int  ckc_host_free(void* p);
int  ckc_synchronize(ckc_ctx* ctx);

/* ---- PCS / APC flavour ---------------------------------------------------------------------------------------- */
/* IKproject(State& q1, State& q2, int active_chain, int ik_sol)  StateValidityCheckerPCS.cpp:140-162
 * = two_robots::calc_specific_IK_solution_R1/_R2 (apc_class.cpp:356-389: FKsolve_rob :51-76, IKsolve_rob :114-245).
 * ac[i] = 0: arm 1 active (q1 kept, q2 computed), 1: arm 2 active.  ac == NULL means all 0.
 * q_out[i] = projected configuration when ok[i] == 1, the input configuration otherwise. */
int ckc_pcs_project(ckc_ctx* ctx, int64_t n, const double* q, const int8_t* ac, const int8_t* ik,
                    double* q_out, uint8_t* ok);

/* isValidRBS(State& q1, State& q2, int active_chain, int ik_sol)  StateValidityCheckerPCS.cpp:455-466
 * (= isValid(state, ac, ik) :360-387): projection, then collision_state on the projected configuration.
 * valid[i] = 1 iff IK feasible and collision free.  q_out and ok may be NULL. */
int ckc_pcs_validate(ckc_ctx* ctx, int64_t n, const double* q, const int8_t* ac, const int8_t* ik,
                     double* q_out, uint8_t* ok, uint8_t* valid);

/* Same validity query with a COMPACT result: the mask (may be NULL) plus, for the valid samples only, their index and
 * projected configuration (what a planner keeps: isValidRBS only leaves a usable state behind when it returns true).
 * idx / q_valid have room for `cap` entries; *n_valid receives the number of valid samples (CKC_ERR_CAPACITY if > cap);
 * the order of the entries is unspecified.  Device-to-host traffic drops from 97 to about 2.4 bytes per sample. */
int ckc_pcs_validate_compact(ckc_ctx* ctx, int64_t n, const double* q, const int8_t* ac, const int8_t* ik, uint8_t* valid,
                             int64_t* n_valid, int32_t* idx, double* q_valid, int64_t cap);
int ckc_gd_validate_compact(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* valid, int64_t* n_valid, int32_t* idx, double* q_valid, int64_t cap);

/* identify_state_ik(State q1, State q2, State ik = {-1,-1})  StateValidityCheckerPCS.cpp:275-316
 * ik01[2*i] = branch for arm-1-active, ik01[2*i+1] = branch for arm-2-active, -1 = none. */
int ckc_identify_ik(ckc_ctx* ctx, int64_t n, const double* q, int8_t* ik01);

/* isValid(const ob::State*)  -- validity_checkers/StateValidityCheckerPCS.cpp:323-357 (the ob::StateValidityChecker
 * override OMPL calls on arbitrary states): identify both IK branches (:328), re-project arm 2 from arm 1 (branch ik[0])
 * and require ||q2 - IK|| <= 0.05 and no collision of (q1, IK) (:331-338), then the same from arm 2 with 0.12 (:342-349).
 * valid[i] = 1/0.  A sample whose branches cannot be identified (-1) is invalid: the reference indexes its branch tables
 * with -1 there (undefined behaviour). */
int ckc_pcs_is_valid(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* valid);

/* check_angle_limits(State q)  apc_class.cpp:304-334 / kdl_class.cpp:213-242 */
int ckc_check_angle_limits(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* ok);

/* checkMotionRBS(State qa1, qa2, qb1, qb2, int active_chain, int ik_sol, 0, 0)  StateValidityCheckerPCS.cpp:486-511
 * one edge = qa[i*12..], qb[i*12..].  ok[i] = 1 iff the whole bisection tree is valid.  nchecks (may be NULL) receives
 * the number of isValidRBS evaluations the batch path performed for the edge. */
int ckc_pcs_check_motion_rbs(ckc_ctx* ctx, int64_t n_edges, const double* qa, const double* qb,
                             const int8_t* ac, const int8_t* ik, uint8_t* ok, int32_t* nchecks);

/* reconstructRBS(State qa1, qa2, qb1, qb2, active_chain, ik_sol, Matrix& M, 0, 1, 1)  StateValidityCheckerPCS.cpp:531-577
 * (post-processing of a solved path: the configurations the bisection visits, in order along the edge).  ONE edge per call:
 * confs[0] = qa, confs[*n_confs - 1] = qb, the valid (projected) midpoints in between; *ok = 1 iff the edge is valid -- the only
 * case the reference uses the matrix in; for an invalid edge the reference stops at the first invalid midpoint of its depth-first
 * order and leaves a partial matrix, here every midpoint evaluated before the frontier died is returned.  CKC_ERR_CAPACITY if
 * the edge needs more than `cap` rows (*n_confs then holds the number required). */
int ckc_pcs_reconstruct_rbs(ckc_ctx* ctx, const double* qa, const double* qb, int active_chain, int ik_sol, double* confs, int64_t cap,
                            int64_t* n_confs, uint8_t* ok);

/* ---- single-arm kinematics, one to one (the batched entry points above fuse them; the drop-in classes expose these members) ---- */
/* two_robots::FKsolve_rob(State q, int robot_num) + get_FK_solution_T1/T2  apc_class.cpp:51-76, :78-112.  robot[i] = 1 or 2;
 * T16[i] = 4x4 row-major tool pose in the world frame. */
int ckc_fk(ckc_ctx* ctx, int64_t n, const double* q6, const int8_t* robot, double* T16);
/* two_robots::IKsolve_rob(Matrix T, int robot_num, int solution_num) + get_IK_solution_q1/q2  apc_class.cpp:114-245.
 * ok[i] = 0 on the reference's early returns (joint limit, unreachable); q6[i] is meaningful when ok[i] = 1. */
int ckc_ik(ckc_ctx* ctx, int64_t n, const double* T16, const int8_t* robot, const int8_t* sol, double* q6, uint8_t* ok);

/* ---- collision ------------------------------------------------------------------------------------------------ */
/* int collisionDetection::collision_state(Matrix P, State q1, State q2)  collisionDetection.cpp:31-494
 * hit[i] = 1 iff any of the pair-table tolerance queries (8 mm) reports closer-than-tolerance (plus the env-2
 * "T7z > 800" rule, :427-428).  The rod matrix P is the one setP() builds (StateValidityCheckerPCS.h:150-158). */
int ckc_collide(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* hit);
/* per-pair flags of the same query without early exit: flags[i*116 + p] (diagnostics / parity tests) */
int ckc_collide_pairs(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* flags);

/* ---- GD / KDL flavour ----------------------------------------------------------------------------------------- */
/* kdl::GD(State q) + get_GD_result()  kdl_class.cpp:68-140 (ChainIkSolverPos_NR, pinv velocity solver, eps 1e-5,
 * 10000 iterations max).  ok[i] = converged AND inside joint limits; converged[i] (may be NULL) = NR converged;
 * q_out[i] = projected configuration when converged, else the input; iters (may be NULL) = Newton iterations. */
int ckc_gd_project(ckc_ctx* ctx, int64_t n, const double* q, double* q_out, uint8_t* ok, uint8_t* converged, int32_t* iters);
/* isValidRBS(State& q)  StateValidityCheckerGD.cpp:180-196 : GD then collision_state */
int ckc_gd_validate(ckc_ctx* ctx, int64_t n, const double* q, double* q_out, uint8_t* valid);
/* checkMotionRBS(State s1, State s2, 0, 0)  StateValidityCheckerGD.cpp:214-241 */
int ckc_gd_check_motion_rbs(ckc_ctx* ctx, int64_t n_edges, const double* qa, const double* qb, uint8_t* ok, int32_t* nchecks);

/* reconstructRBS(State q1, State q2, Matrix& M, 0, 1, 1)  StateValidityCheckerGD.cpp:258-296 -- see ckc_pcs_reconstruct_rbs */
int ckc_gd_reconstruct_rbs(ckc_ctx* ctx, const double* qa, const double* qb, double* confs, int64_t cap, int64_t* n_confs, uint8_t* ok);
/* kdl::FK(State q) + get_FK_solution()  kdl_class.cpp:251-285: tip frame of the 12-joint chain base 1 -> base 2 (4x4 row-major);
 * a closed configuration gives Trans(D, 0, 0) */
int ckc_kdl_fk(ckc_ctx* ctx, int64_t n, const double* q12, double* T16);

/* ---- RX (relaxed constraint) flavour -------------------------------------------------------------------------- */
/* isValidRBS(State q)  StateValidityCheckerRX.cpp:201-214 : check_relax_constraint (:105-135) && joint limits && !collision.
 * resid (may be NULL) receives sqrt(Ka d^2 + roll^2 + pitch^2 + yaw^2). */
int ckc_rx_validate(ckc_ctx* ctx, int64_t n, const double* q, double epsilon, uint8_t* valid, double* resid);

/* ---- device-resident entry points (inputs already in HBM; used for throughput runs and by GPU-side callers) ---- */
/* d_q: SoA [12][n] doubles; d_ik: n int8; outputs: d_q_out SoA [12][n] (may be NULL), d_ok, d_valid n bytes (d_ok may be NULL) */
int ckc_pcs_validate_dev(ckc_ctx* ctx, int64_t n, const double* d_q, const int8_t* d_ac, const int8_t* d_ik,
                         double* d_q_out, uint8_t* d_ok, uint8_t* d_valid, void* stream);
/* S-PCS samples generated on the device from the counter-based generator of DESIGN.md (seed, first index i0):
 * writes only the masks; optionally the generated/projected configurations (SoA, may be NULL) */
int ckc_spcs_validate_seeded_dev(ckc_ctx* ctx, uint64_t seed, uint64_t i0, int64_t n, double* d_q_out,
                                 uint8_t* d_ok, uint8_t* d_valid, void* stream);
int ckc_gd_validate_dev(ckc_ctx* ctx, int64_t n, const double* d_q, double* d_q_out, uint8_t* d_valid, void* stream);
int ckc_sgd_validate_seeded_dev(ckc_ctx* ctx, uint64_t seed, uint64_t i0, int64_t n, double* d_q_out,
                                uint8_t* d_ok, uint8_t* d_valid, void* stream);
/* A device-resident call runs its kernels in order on the CALLER's stream with one of four internal scratch sets (rotated per call;
 * a set's previous user is awaited through an event, whatever stream it ran on).  Calls issued on different caller streams therefore
 * overlap on the device -- the way to get the collision tail of one batch hidden behind the projection of the next is to issue
 * consecutive batches on two streams.  (CKC_PCS_FAN / CKC_GD_FAN = 1 instead cut ONE call into chunks on internal streams, forked
 * from / joined to the caller's stream; only one such call may be in flight per context.)
 * ckc_set_overlap(0): the host pipelines stop moving the collision stage to their high-priority stream and nothing fans out --
 * every kernel runs alone and in order (used when single kernels are timed).  Default: enabled. */
int ckc_set_overlap(ckc_ctx* ctx, int enable);
/* Per-stage device timing (CUDA events recorded on the launching stream around each kernel stage; resolved lazily).
 * stage 0 = projection kernel (PCS / GD / RX), 1 = unused since the link frames are computed inside the collision kernel (with
 * CKC_TIMELINE set: the uploads of the host pipeline), 2 = collision, 3 = RBS expand / push / tail.
 * ms[s] = accumulated device milliseconds, launches[s] = number of timed stage executions since the last reset. */
int ckc_set_stage_timing(ckc_ctx* ctx, int enable);
int ckc_get_stage_times(ckc_ctx* ctx, double ms[4], int64_t launches[4]);
/* FMA-chain peak probes for the roofline denominator (DESIGN.md "Measurement"): runs `iters` dependent-chain
 * FMAs per thread on every SM and returns the achieved TFLOP/s (2 flop per FMA). precision: 32 or 64 */
int ckc_measure_fma_peak(ckc_ctx* ctx, int precision, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CKC_H_ */
