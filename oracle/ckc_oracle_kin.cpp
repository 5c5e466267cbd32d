/*
 * ckc_oracle_kin.cpp -- CPU ORACLE (test infrastructure, NOT product code): sample generator + APC/PCS kinematics.
 * Restates proj_classes/apc_class.cpp and the projection part of validity_checkers/StateValidityCheckerPCS.cpp.
 * Pinned against apc_class.cpp compiled from source (oracle/_ref/libapc_ref.so) and the prebuilt tests/trbspcs.
 */
#include "ckc_oracle.h"
#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Counter-based generator (specified in DESIGN.md "Synthetic inputs"; the CUDA path implements the
 * same specification independently): u(seed,i,k) = (splitmix64(splitmix64(seed) ^ (16*i + k)) >> 11) * 2^-53
 * ------------------------------------------------------------------------------------------------ */
extern "C" uint64_t ckco_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

extern "C" double ckco_uniform(uint64_t seed, uint64_t i, uint32_t k) {
    uint64_t key = ckco_splitmix64(seed);
    uint64_t x = ckco_splitmix64(key ^ (i * 16ULL + (uint64_t)k));
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

static double deg2rad(double deg) { return deg * CKCO_PI_ / 180.0; }      /* apc_class.cpp:591-593 */

/* rand_q_ambient box, apc_class.cpp:643-661 (q6 sampled in [-PI_, PI_], not its +-400 deg limit) */
extern "C" void ckco_sample_pcs(uint64_t seed, uint64_t i, double q[12], int* ik) {
    const double lo[6] = { -deg2rad(165), -deg2rad(110), deg2rad(-110), -deg2rad(160), -deg2rad(120), -CKCO_PI_ };
    const double hi[6] = {  deg2rad(165),  deg2rad(110), deg2rad(70),    deg2rad(160),  deg2rad(120),  CKCO_PI_ };
    for (int j = 0; j < 12; j++) {
        double u = ckco_uniform(seed, i, (uint32_t)j);
        q[j] = lo[j % 6] + u * (hi[j % 6] - lo[j % 6]);
    }
    int s = (int)(ckco_uniform(seed, i, 12) * 8.0);
    *ik = s > 7 ? 7 : s;
}

extern "C" void ckco_sample_gd(uint64_t seed, uint64_t i, double q[12]) {
    for (int j = 0; j < 12; j++)
        q[j] = -CKCO_PI_ + ckco_uniform(seed, i, (uint32_t)j) * (2 * CKCO_PI_);   /* GD.cpp:60-61 */
}

/* ------------------------------------------------------------------------------------------------
 * Kinematic constants: apc_class.cpp:5-25, StateValidityCheckerPCS.h:23-26,37
 * ------------------------------------------------------------------------------------------------ */
extern "C" void ckco_kin_init(ckco_kin* k, int env) {
    k->b = 166; k->l1 = 124; k->l2 = 270; k->l3a = 70; k->l3b = 149.6; k->l4 = 152.4;
    k->l5 = 72 - 13. / 2;             /* = 65.5  (apc_class.cpp:11, floating division) */
    k->lee = 60 + 13. / 2;            /* = 66.5 */
    k->D = (env == 1) ? 900. : 1200.;
    k->L = (env == 1) ? 300. : 500.;
    k->lim[0][0] = -deg2rad(165); k->lim[0][1] = deg2rad(165);
    k->lim[1][0] = -deg2rad(110); k->lim[1][1] = deg2rad(110);
    k->lim[2][0] = deg2rad(-110); k->lim[2][1] = deg2rad(70);
    k->lim[3][0] = -deg2rad(160); k->lim[3][1] = deg2rad(160);
    k->lim[4][0] = -deg2rad(120); k->lim[4][1] = deg2rad(120);
    k->lim[5][0] = -deg2rad(400); k->lim[5][1] = deg2rad(400);
    k->pose1[0] = -k->D / 2; k->pose1[1] = 0; k->pose1[2] = 0;
    k->pose2[0] =  k->D / 2; k->pose2[1] = 0; k->pose2[2] = CKCO_PI_;
}

/* 4x4 row-major helpers; product association follows MatricesMult (apc_class.cpp:427-436): sum starts at 0
 * and accumulates k = 0..3 in order */
static void mat4_mul(const double A[16], const double B[16], double C[16]) {
    double R[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int kk = 0; kk < 4; kk++) s += A[4 * i + kk] * B[4 * kk + j];
            R[4 * i + j] = s;
        }
    memcpy(C, R, sizeof(R));
}
static void mat4_identity(double A[16]) { memset(A, 0, 16 * sizeof(double)); A[0] = A[5] = A[10] = A[15] = 1; }
static void mat4_trans(double A[16], double x, double y, double z) { mat4_identity(A); A[3] = x; A[7] = y; A[11] = z; }
static void mat4_rotz(double A[16], double t) { mat4_identity(A); double c = cos(t), s = sin(t); A[0] = c; A[1] = -s; A[4] = s; A[5] = c; }
static void mat4_roty(double A[16], double t) { mat4_identity(A); double c = cos(t), s = sin(t); A[0] = c; A[2] = s; A[8] = -s; A[10] = c; }
static void mat4_rotx(double A[16], double t) { mat4_identity(A); double c = cos(t), s = sin(t); A[5] = c; A[6] = -s; A[9] = s; A[10] = c; }

/* FKsolve_rob (apc_class.cpp:51-76).  The reference evaluates the fully expanded symbolic entries of the
 * serial chain below; SURVEY.md appendix A verified equality to 7e-16 (rotation) / 7e-13 mm (position), and
 * tests/test_oracle_ref.py re-checks this against apc_class.cpp compiled from source:
 *   T = Trans(x0,y0,0) Rz(theta) Tz(b+l1) Rz(q1) Ry(q2) Tz(l2) Ry(q3) T(l3b+l4,0,l3a) Rx(q4) Ry(q5) Tx(l5+lee) Rx(q6) */
extern "C" void ckco_fk(const ckco_kin* k, const double q[6], int robot, double T[16]) {
    const double* pose = (robot == 1) ? k->pose1 : k->pose2;
    double A[16], B[16];
    mat4_trans(T, pose[0], pose[1], 0);
    mat4_rotz(A, pose[2]);                         mat4_mul(T, A, T);
    mat4_trans(A, 0, 0, k->b + k->l1);             mat4_mul(T, A, T);
    mat4_rotz(A, q[0]);                            mat4_mul(T, A, T);
    mat4_roty(A, q[1]);                            mat4_mul(T, A, T);
    mat4_trans(A, 0, 0, k->l2);                    mat4_mul(T, A, T);
    mat4_roty(A, q[2]);                            mat4_mul(T, A, T);
    mat4_trans(A, k->l3b + k->l4, 0, k->l3a);      mat4_mul(T, A, T);
    mat4_rotx(A, q[3]);                            mat4_mul(T, A, T);
    mat4_roty(A, q[4]);                            mat4_mul(T, A, T);
    mat4_trans(A, k->l5 + k->lee, 0, 0);           mat4_mul(T, A, T);
    mat4_rotx(B, q[5]);                            mat4_mul(T, B, T);
}

/* IKsolve_rob (apc_class.cpp:114-245), joint limits always included (apc_class.h:121) */
extern "C" int ckco_ik(const ckco_kin* k, const double T[16], int robot, int sol, double q[6]) {
    static const double q1_add_sol[8] = { 0, 0, 0, 0, CKCO_PI_, CKCO_PI_, CKCO_PI_, CKCO_PI_ };   /* :117 */
    static const double q2_sign[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };                              /* :118 */
    static const double q3_sign[8] = { 1, -1, 1, -1, 1, -1, 1, -1 };                              /* :119 */
    static const double sign456[8] = { 1, -1, -1, 1, -1, 1, 1, -1 };                              /* :120 */
    const double* V = (robot == 1) ? k->pose1 : k->pose2;
    const double cth = cos(V[2]), sth = sin(V[2]);
    const double b = k->b, l1 = k->l1, l2 = k->l2, l3a = k->l3a, l3b = k->l3b, l4 = k->l4, l5 = k->l5, lee = k->lee;
#define TT(i, j) T[4 * (i) + (j)]
    double R[3][3], p[3];
    for (int j = 0; j < 3; j++) {          /* :128-130 target rotation in the robot base frame */
        R[0][j] = cth * TT(0, j) + sth * TT(1, j) - cth * TT(3, j) * V[0] - sth * TT(3, j) * V[1];
        R[1][j] = cth * TT(1, j) - sth * TT(0, j) - cth * TT(3, j) * V[1] + sth * TT(3, j) * V[0];
        R[2][j] = TT(2, j);
    }
    p[0] = cth * TT(0, 3) + sth * TT(1, 3) - cth * TT(3, 3) * V[0] - sth * TT(3, 3) * V[1];      /* :132-134 */
    p[1] = cth * TT(1, 3) - sth * TT(0, 3) - cth * TT(3, 3) * V[1] + sth * TT(3, 3) * V[0];
    p[2] = TT(2, 3);
#undef TT
    const double x6[3] = { R[0][0], R[1][0], R[2][0] };
    const double p5[3] = { p[0] - x6[0] * (l5 + lee), p[1] - x6[1] * (l5 + lee), p[2] - x6[2] * (l5 + lee) };   /* :137 */

    /* q1 :141-148 */
    q[0] = atan2(p5[1], p5[0]) + q1_add_sol[sol];
    if (q[0] > CKCO_PI_) q[0] -= 2 * CKCO_PI_;
    if (q[0] < -CKCO_PI_) q[0] += 2 * CKCO_PI_;
    if (fabs(q[0]) > k->lim[0][1]) return 0;

    const double k2 = p5[0] * p5[0] + p5[1] * p5[1];                      /* :152 */
    const double D2 = k2 + (p5[2] - (l1 + b)) * (p5[2] - (l1 + b));       /* :153 */
    const double l34 = sqrt((l3a * l3a + (l3b + l4) * (l3b + l4)));       /* :154 */

    /* q3 :157-175 */
    double sinphi;
    {
        const double alpha = atan2(l3b + l4, l3a);
        const double cosphi = (D2 - l2 * l2 - l34 * l34) / (2 * l2 * l34);
        sinphi = (1 - cosphi * cosphi);
        if (sinphi < 0) return 0;
        sinphi = q3_sign[sol] * sqrt(sinphi);
        q[2] = -(atan2(sinphi, cosphi) + alpha);
        if (q[2] > CKCO_PI_) q[2] -= 2 * CKCO_PI_;
        if (q[2] < -CKCO_PI_) q[2] += 2 * CKCO_PI_;
        if (q[2] < k->lim[2][0] || q[2] > k->lim[2][1]) return 0;
    }
    /* q2 :178-189 */
    {
        const double sinalpha1 = -l34 * (sinphi) / sqrt(D2);
        const double alpha1 = atan2(-q2_sign[sol] * sinalpha1, sqrt(1 - sinalpha1 * sinalpha1));
        const double alpha2 = atan2(p5[2] - l1 - b, sqrt(k2));
        q[1] = q2_sign[sol] * (alpha1 + alpha2 - CKCO_PI_ / 2);
        if (q[1] > CKCO_PI_) q[1] -= 2 * CKCO_PI_;
        if (q[1] < -CKCO_PI_) q[1] += 2 * CKCO_PI_;
        if (fabs(q[1]) > k->lim[1][1]) return 0;
    }
    const double c23 = cos(q[1] + q[2]), s23 = sin(q[1] + q[2]);          /* :199-202 */
    const double c1 = cos(q[0]), s1 = sin(q[0]);
    /* q5 :205-211 */
    {
        const double S = R[0][0] * c23 * c1 - R[2][0] * s23 + R[1][0] * c23 * s1;
        q[4] = atan2(sign456[sol] * sqrt(1 - S * S), S);
        q[4] = fabs(q[4]) < 1e-4 ? 0 : q[4];
        if (fabs(q[4]) > k->lim[4][1]) return 0;
    }
    /* q4, q6 :215-233 */
    if (q[4]) {
        const double sinq4 = R[1][0] * c1 - R[0][0] * s1;
        const double cosq4 = R[2][0] * c23 + R[0][0] * s23 * c1 + R[1][0] * s23 * s1;
        q[3] = atan2(sign456[sol] * sinq4, -sign456[sol] * cosq4);
        const double sinq6 = R[0][1] * c23 * c1 - R[2][1] * s23 + R[1][1] * c23 * s1;
        const double cosq6 = R[0][2] * c23 * c1 - R[2][2] * s23 + R[1][2] * c23 * s1;
        q[5] = atan2(sign456[sol] * sinq6, sign456[sol] * cosq6);
    } else {
        q[3] = 0;
        q[5] = atan2(R[2][1], R[2][2]);
    }
    if (fabs(q[3]) > k->lim[3][1] || fabs(q[5]) > k->lim[5][1]) return 0;
    for (int i = 0; i < 6; i++)                                            /* :236-238 */
        if (fabs(q[i]) < 1e-5) q[i] = 0;
    return 1;
}

/* calc_specific_IK_solution_R1 (apc_class.cpp:356-371): T2 = FK1(q1) * Q * diag(-1,-1,1,1), Q = Trans(L,0,0) */
extern "C" int ckco_project_r1(const ckco_kin* k, const double q1[6], int sol, double q2[6]) {
    double T[16], Q[16], F[16];
    ckco_fk(k, q1, 1, T);
    mat4_trans(Q, k->L, 0, 0);
    mat4_mul(T, Q, T);
    mat4_identity(F); F[0] = -1; F[5] = -1;
    mat4_mul(T, F, T);
    return ckco_ik(k, T, 2, sol, q2);
}

/* calc_specific_IK_solution_R2 (apc_class.cpp:373-389): T1 = FK2(q2) * diag(-1,-1,1,1) * Q^-1.
 * InvertMatrix (:439-575, cofactor expansion) of Trans(L,0,0) is exactly Trans(-L,0,0) (all products of 0/1/L). */
extern "C" int ckco_project_r2(const ckco_kin* k, const double q2[6], int sol, double q1[6]) {
    double T[16], Qi[16], F[16];
    ckco_fk(k, q2, 2, T);
    mat4_identity(F); F[0] = -1; F[5] = -1;
    mat4_mul(T, F, T);
    mat4_trans(Qi, -k->L, 0, 0);
    mat4_mul(T, Qi, T);
    return ckco_ik(k, T, 1, sol, q1);
}

extern "C" int ckco_check_angle_limits(const ckco_kin* k, const double q[12]) {      /* apc_class.cpp:304-334 */
    for (int r = 0; r < 2; r++) {
        const double* a = q + 6 * r;
        if (fabs(a[0]) > k->lim[0][1]) return 0;
        if (fabs(a[1]) > k->lim[1][1]) return 0;
        if (a[2] < k->lim[2][0] || a[2] > k->lim[2][1]) return 0;
        if (fabs(a[3]) > k->lim[3][1]) return 0;
        if (fabs(a[4]) > k->lim[4][1]) return 0;
        if (fabs(a[5]) > k->lim[5][1]) return 0;
    }
    return 1;
}

/* IKproject(q1,q2,active_chain,ik_sol) StateValidityCheckerPCS.cpp:140-162 */
extern "C" int ckco_pcs_project(const ckco_kin* k, double q[12], int active_chain, int ik_sol) {
    double out[6];
    if (!active_chain) {
        if (!ckco_project_r1(k, q, ik_sol, out)) return 0;
        memcpy(q + 6, out, sizeof(out));
    } else {
        if (!ckco_project_r2(k, q + 6, ik_sol, out)) return 0;
        memcpy(q, out, sizeof(out));
    }
    return 1;
}

static double norm_dist6(const double* a, const double* b) {     /* StateValidityChecker::normDistance PCS.cpp:434-439 */
    double s = 0;
    for (int i = 0; i < 6; i++) s += pow(a[i] - b[i], 2);
    return sqrt(s);
}

/* identify_state_ik(q1,q2,{-1,-1}) StateValidityCheckerPCS.cpp:275-316 with calc_all_IK_solutions_{1,2}
 * (apc_class.cpp:249-277): first feasible branch (ascending index) within Euclid distance 0.1 */
extern "C" void ckco_identify_state_ik(const ckco_kin* k, const double q[12], int ik[2]) {
    ik[0] = ik[1] = -1;
    double cand[6];
    int n = 0;
    for (int s = 0; s < 8; s++)
        if (ckco_project_r1(k, q, s, cand)) {
            n++;
            if (norm_dist6(cand, q + 6) < 1e-1) { ik[0] = s; break; }
        }
    if (n == 0) return;                 /* :283-284 early return leaves ik[1] unknown */
    for (int s = 0; s < 8; s++)
        if (ckco_project_r2(k, q + 6, s, cand)) {
            if (norm_dist6(cand, q) < 1e-1) { ik[1] = s; break; }
        }
}
