/*
 * ckc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain scalar FP64 restatement of the reference's batched-state-validity hot path
 * (avishais/ABB_dualarm_planning), used ONLY by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker the CUDA path is compared with.
 * Nothing under abb_dualarm_planning_b200/ may include, link or call this.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Pinning status (see DESIGN.md "Oracle"):
 *   - APC/PCS projection, FK, joint limits, identify_state_ik : pinned against apc_class.cpp compiled from
 *     source (oracle/_ref/libapc_ref.so) and against the prebuilt tests/trbspcs binary (oracle/refshim).
 *   - collision_state (frames, pair table, PQP tolerance semantics, TriDist): pinned against the PQP v1.3
 *     copy statically linked into tests/trbspcs (mask + TriDist values), golden fixtures in tests/golden/.
 *   - RBS: pinned against the binary's checkMotionRBS.
 *   - KDL Newton projection (GD) and RX residual: PARITY UNPINNED against real KDL (Orocos KDL 1.4 is not
 *     in /root/reference and not installed); pinned only by the reference's fixtures (paths/path.txt closure,
 *     tests/results/data/rlx_rand_confs_*.txt) and its published statistics.
 */
#ifndef CKC_ORACLE_H_
#define CKC_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants of the reference (apc_class.h:10, kdl_class.h:24) ---- */
#define CKCO_PI_ 3.1416
#define CKCO_NUM_FRAMES 21   /* 8 per robot (base..link6 + EE/rod pose) + 2 mixed-up quirk frames + 3 static obstacle poses */
#define CKCO_MAX_PAIRS 116

/* ---- counter-based sample generator shared (by specification) with the CUDA path ---- */
uint64_t ckco_splitmix64(uint64_t x);
double   ckco_uniform(uint64_t seed, uint64_t i, uint32_t k);          /* in [0,1) */
/* S-PCS sample i: 12 joints uniform in the rand_q_ambient box (apc_class.cpp:643-661), ik = floor(u*8) */
void     ckco_sample_pcs(uint64_t seed, uint64_t i, double q[12], int* ik);
/* S-GD sample i: 12 joints uniform in [-3.1416, 3.1416] (StateValidityCheckerGD.cpp:60-61) */
void     ckco_sample_gd(uint64_t seed, uint64_t i, double q[12]);

/* ---- kinematics (apc_class.cpp) ---- */
typedef struct ckco_kin {
    double b, l1, l2, l3a, l3b, l4, l5, lee;      /* apc_class.cpp:5-12 */
    double L;                                     /* rod length */
    double D;                                     /* distance between the robot bases */
    double lim[6][2];                             /* joint limits [min,max] rad, apc_class.cpp:19-25 */
    double pose1[3], pose2[3];                    /* (x,y,theta) of each base, StateValidityCheckerPCS.h:37 */
} ckco_kin;

void ckco_kin_init(ckco_kin* k, int env);                                  /* env 1: D=900,L=300; env 2: 1200/500 */
void ckco_fk(const ckco_kin* k, const double q[6], int robot, double T[16]);       /* apc_class.cpp:51-76 (row-major 4x4) */
int  ckco_ik(const ckco_kin* k, const double T[16], int robot, int sol, double q[6]); /* apc_class.cpp:114-245 */
/* calc_specific_IK_solution_R1 / _R2 (apc_class.cpp:356-389): active arm joints -> passive arm joints */
int  ckco_project_r1(const ckco_kin* k, const double q1[6], int sol, double q2[6]);
int  ckco_project_r2(const ckco_kin* k, const double q2[6], int sol, double q1[6]);
int  ckco_check_angle_limits(const ckco_kin* k, const double q[12]);      /* apc_class.cpp:304-334 */
/* StateValidityChecker::IKproject(q1,q2,ac,ik) PCS.cpp:140-162. q in/out (12). returns 1 when projected */
int  ckco_pcs_project(const ckco_kin* k, double q[12], int active_chain, int ik_sol);
/* identify_state_ik(q1,q2) PCS.cpp:275-316: ik[0], ik[1] (-1 = none) */
void ckco_identify_state_ik(const ckco_kin* k, const double q[12], int ik[2]);

/* ---- meshes, hierarchy and the collision world (collisionDetection.cpp) ---- */
typedef struct ckco_world ckco_world;

/* mesh ids (order is part of the packed-mesh file format) */
enum { CKCO_M_BASE = 0, CKCO_M_LINK1, CKCO_M_LINK2, CKCO_M_LINK3, CKCO_M_LINK4, CKCO_M_LINK5, CKCO_M_LINK6,
       CKCO_M_EE, CKCO_M_TABLE, CKCO_M_OBS, CKCO_M_CONE, CKCO_M_ROD, CKCO_NUM_MESHES };

/* Load from a directory of .tris files (the reference's format, collisionDetection.cpp:502-518) */
ckco_world* ckco_world_create_from_tris(int env, const char* tris_dir);
/* Load from the packed mesh file shipped with the repo (tools/pack_meshes.py) */
ckco_world* ckco_world_create_from_pack(int env, const char* pack_path);
void        ckco_world_destroy(ckco_world* w);
int         ckco_world_num_pairs(const ckco_world* w);
int         ckco_world_mesh_ntris(const ckco_world* w, int mesh);
const double* ckco_world_mesh_tris(const ckco_world* w, int mesh);        /* ntris*9 doubles */

/* 21 frames: frame f = R (9, row-major) followed by T (3): collisionDetection.cpp:72-260.
 * f 0..7 = robot 1 base, link1, link2, link3, link4, link5, link6, EE/rod (R6,T7);
 * f 8..15 = robot 2 same; f 16 = (R22, T2_t) quirk of res[78]; f 17 = (R2, T2_t2) quirk of res[97]/res[110];
 * f 18..20 = static obstacle poses obs1/obs2/obs3 (collisionDetection.cpp:363-406 env 1, :431-456 env 2). */
void ckco_link_frames(const ckco_world* w, const double q[12], double frames[CKCO_NUM_FRAMES][12]);

typedef struct ckco_pair { int mesh_a, frame_a, mesh_b, frame_b; } ckco_pair;
/* the pair table of collisionDetection.cpp:267-355 (+ :372-420 env 1, :443-472 env 2) */
const ckco_pair* ckco_world_pairs(const ckco_world* w);

/* PQP v1.3 TriDist / SegPoints semantics (external library, restated from its published algorithm;
 * pinned against the compiled copy in tests/trbspcs).  S, T: 3 vertices x 3 coords. returns distance */
double ckco_tri_dist(double P[3], double Q[3], const double S[3][3], const double T[3][3]);

typedef struct ckco_collide_stats { int64_t bv_tests, tri_tests, pairs_hit; } ckco_collide_stats;
/* collision_state (collisionDetection.cpp:31-494): returns 0 free / 1 collision. All pairs are evaluated
 * (quirk 6) when stats != NULL or early_exit == 0; per_pair (116 bytes) receives each pair's flag */
int ckco_collision_state(const ckco_world* w, const double q[12], int early_exit,
                         ckco_collide_stats* stats, uint8_t* per_pair);
/* brute force over all triangle pairs of each table entry (validates the hierarchy's conservativeness) */
int ckco_collision_state_brute(const ckco_world* w, const double q[12], uint8_t* per_pair);
/* min over the pair table of the exact mesh-mesh distance (for contact-margin analysis of mismatches) */
double ckco_min_pair_distance(const ckco_world* w, const double q[12]);

/* ---- checker-level functions ---- */
/* isValidRBS PCS.cpp:455-466 (project passive arm, then collision_state). q in/out */
int ckco_pcs_is_valid_rbs(const ckco_kin* k, const ckco_world* w, double q[12], int ac, int ik);
/* isValid(const ob::State*) PCS.cpp:323-357: identify both IK branches, both active-chain projections must reproduce the
 * other arm within 0.05 / 0.12 rad (Euclid) and be collision free.  A configuration whose branches cannot be identified
 * is invalid (the reference indexes its branch tables with -1 there: undefined behaviour) */
int ckco_pcs_is_valid_full(const ckco_kin* k, const ckco_world* w, const double q[12]);
/* checkMotionRBS PCS.cpp:486-511. nchecks (may be NULL) counts isValidRBS calls (isValid_counter) */
int ckco_pcs_check_motion_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12],
                              int ac, int ik, int64_t* nchecks);

/* ---- GD / KDL chain (kdl_class.cpp + Orocos KDL 1.4 semantics; PARITY UNPINNED vs real KDL) ---- */
void ckco_kdl_fk(const ckco_kin* k, const double q[12], double T[16]);               /* kdl_class.cpp:251-278 */
/* kdl::GD (kdl_class.cpp:68-140). returns 1 if converged and inside joint limits; 0 otherwise.
 * converged (may be NULL) = NR converged; iters = Newton iterations used. q_out always written when converged */
int  ckco_gd_project(const ckco_kin* k, const double q_in[12], double q_out[12], int* converged, int* iters);
int  ckco_gd_is_valid(const ckco_kin* k, const ckco_world* w, const double q_in[12], double q_out[12]); /* GD.cpp:180-196 */
int  ckco_gd_check_motion_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12],
                              int64_t* nchecks);                                      /* GD.cpp:199-241 */
/* RX check_relax_constraint RX.cpp:105-135: residual (returned through *resid) < eps */
int  ckco_rx_check(const ckco_kin* k, const double q[12], double eps, double* resid);

/* ---- members added in round 2 (ckc_oracle_more.cpp) ---- */
/* reconstructRBS PCS.cpp:531-577 / GD.cpp:246-296: rows (cap x 12) receives the matrix {a, ordered valid midpoints, b} the reference
 * builds (partial when the edge is invalid: its recursion stops at the first invalid midpoint); returns the reference's bool */
int ckco_pcs_reconstruct_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], int ac, int ik,
                             double* rows, int cap, int* nrows);
int ckco_gd_reconstruct_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], double* rows, int cap, int* nrows);
/* checkMotion PCS.cpp:389-432 (fixed grid, breadth-first bisection order) / GD.cpp:136-175 (serial grid) */
int ckco_pcs_check_motion(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], int ac, int ik, int* nchecks);
int ckco_gd_check_motion(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], int* nchecks);
/* IKproject(q1, q2, ik, active_chain, ik_nn) PCS.cpp:94-138 */
int ckco_pcs_ikproject5(const ckco_kin* k, const ckco_world* w, double q[12], int ik_out[2], int* ac, const int ik_nn[2]);

/* ---- batch drivers (OpenMP-free, single thread; bench.py forks processes for multi-core) ---- */
/* S-PCS validate: out_ok[i] = IK feasible, out_valid[i] = feasible && !collision; q_out (n*12, AoS) optional */
void ckco_batch_pcs_validate(const ckco_kin* k, const ckco_world* w, int64_t n, const double* q_aos,
                             const int8_t* ac, const int8_t* ik, uint8_t* out_ok, uint8_t* out_valid,
                             double* q_out, ckco_collide_stats* stats_sum);

#ifdef __cplusplus
}
#endif
#endif
