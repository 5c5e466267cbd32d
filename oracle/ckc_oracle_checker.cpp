/*
 * ckc_oracle_checker.cpp -- CPU ORACLE (test infrastructure, NOT product code): checker-level functions.
 * Restates validity_checkers/StateValidityChecker{PCS,GD,RX}.cpp (isValidRBS, checkMotionRBS, check_relax_constraint)
 * and proj_classes/kdl_class.cpp + the Orocos KDL 1.4 solvers it calls (ChainIkSolverPos_NR, ChainIkSolverVel_pinv,
 * ChainJntToJacSolver, diff(Frame,Frame), Rotation::GetRot).  KDL is an external dependency that is NOT under
 * /root/reference and is not installed: its algorithm is restated from the published library source as recalled,
 * so GD joint solutions are PARITY UNPINNED against real KDL (DESIGN.md "Oracle").
 */
#include "ckc_oracle.h"
#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* PCS: isValidRBS / checkMotionRBS                                                                   */
/* ------------------------------------------------------------------------------------------------ */
extern "C" int ckco_pcs_is_valid_rbs(const ckco_kin* k, const ckco_world* w, double q[12], int ac, int ik) {
    if (!ckco_pcs_project(k, q, ac, ik)) return 0;                 /* PCS.cpp:459 */
    if (!ckco_collision_state(w, q, 1, NULL, NULL)) return 1;      /* PCS.cpp:462 */
    return 0;
}

extern "C" int ckco_pcs_is_valid_full(const ckco_kin* k, const ckco_world* w, const double q[12]) {
    int ik[2];
    ckco_identify_state_ik(k, q, ik);                               /* PCS.cpp:328 */
    if (ik[0] < 0 || ik[1] < 0) return 0;
    double c[12], out[6], s;
    if (!ckco_project_r1(k, q, ik[0], out)) return 0;               /* :331 */
    s = 0; for (int i = 0; i < 6; i++) s += pow(q[6 + i] - out[i], 2);
    if (sqrt(s) > 0.5e-1) return 0;                                 /* :333 */
    memcpy(c, q, 48); memcpy(c + 6, out, 48);
    if (ckco_collision_state(w, c, 1, NULL, NULL)) return 0;        /* :335 */
    if (!ckco_project_r2(k, q + 6, ik[1], out)) return 0;           /* :342 */
    s = 0; for (int i = 0; i < 6; i++) s += pow(q[i] - out[i], 2);
    if (sqrt(s) > 0.12) return 0;                                   /* :344 */
    memcpy(c, out, 48); memcpy(c + 6, q + 6, 48);
    if (ckco_collision_state(w, c, 1, NULL, NULL)) return 0;        /* :346 */
    return 1;
}

static double norm12(const double* a, const double* b) {          /* normDistanceDuo PCS.cpp:513-518 */
    double sum = 0;
    for (int i = 0; i < 6; i++) sum += pow(a[i] - b[i], 2) + pow(a[i + 6] - b[i + 6], 2);
    return sqrt(sum);
}

static int pcs_rbs(const ckco_kin* k, const ckco_world* w, const double* a, const double* b, int ac, int ik, int depth, int64_t* n) {
    const double d = norm12(a, b);                                 /* PCS.cpp:491-493 */
    if (d < 0.05) return 1;
    if (depth > 150) return 0;                                     /* :495 */
    double m[12];
    for (int i = 0; i < 12; i++) m[i] = (a[i] + b[i]) / 2;         /* midpoint :520-526 */
    if (n) (*n)++;
    if (!ckco_pcs_is_valid_rbs(k, w, m, ac, ik)) return 0;         /* :500 */
    return pcs_rbs(k, w, a, m, ac, ik, depth + 1, n) && pcs_rbs(k, w, m, b, ac, ik, depth + 1, n);   /* :506 */
}

extern "C" int ckco_pcs_check_motion_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12],
                                         int ac, int ik, int64_t* nchecks) {
    if (nchecks) *nchecks = 0;
    return pcs_rbs(k, w, qa, qb, ac, ik, 0, nchecks);
}

/* ------------------------------------------------------------------------------------------------ */
/* KDL chain (kdl_class.cpp:32-45): 13 segments, 12 joints, from base 1 to base 2                     */
/* ------------------------------------------------------------------------------------------------ */
struct seg { int axis; double tip[3]; };       /* axis: -1 none, 0 X, 1 Y, 2 Z; pose(q) = Rot(axis,q) then translate tip */

static void kdl_chain(const ckco_kin* k, seg s[13]) {
    /* kdl_class.cpp:11-12: l5 = 72-13/2 = 66, lee = 60+13/2 = 66 (integer division) */
    const double b = 166, l1 = 124, l2 = 270, l3a = 70, l3b = 149.6, l4 = 152.4, l5 = 72 - 13 / 2, lee = 60 + 13 / 2;
    const double L = k->L;
    const seg c[13] = {
        { -1, { 0, 0, b } }, { 2, { 0, 0, l1 } }, { 1, { 0, 0, l2 } }, { 1, { l3b, 0, l3a } }, { 0, { l4, 0, 0 } },
        { 1, { l5, 0, 0 } }, { 0, { 2 * lee + L, 0, 0 } },
        { 0, { l5, 0, 0 } }, { 1, { l4, 0, 0 } }, { 0, { l3b, 0, -l3a } }, { 1, { 0, 0, -l2 } }, { 1, { 0, 0, -l1 } }, { 2, { 0, 0, -b } } };
    memcpy(s, c, sizeof(c));
}

static void rot_axis(double R[9], int axis, double t) {
    const double c = cos(t), s = sin(t);
    memset(R, 0, 9 * sizeof(double));
    if (axis == 0) { R[0] = 1; R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
    else if (axis == 1) { R[0] = c; R[2] = s; R[4] = 1; R[6] = -s; R[8] = c; }
    else { R[0] = c; R[1] = -s; R[3] = s; R[4] = c; R[8] = 1; }
}
static void m3mul(double C[9], const double A[9], const double B[9]) {
    double R[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, R, sizeof(R));
}

/* FK of the chain in chain joint order; optionally the 6x12 geometric Jacobian at the tip (base frame):
 * column j = [ axis_j x (p_tip - p_j) ; axis_j ]  (KDL ChainJntToJacSolver) */
static void chain_fk(const ckco_kin* k, const double qc[12], double R[9], double p[3], double J[6][12]) {
    seg s[13]; kdl_chain(k, s);
    double ax[12][3], org[12][3];
    memset(R, 0, 9 * sizeof(double)); R[0] = R[4] = R[8] = 1; p[0] = p[1] = p[2] = 0;
    int j = 0;
    for (int i = 0; i < 13; i++) {
        if (s[i].axis >= 0) {
            for (int r = 0; r < 3; r++) { ax[j][r] = R[3 * r + s[i].axis]; org[j][r] = p[r]; }
            double Rq[9]; rot_axis(Rq, s[i].axis, qc[j]);
            m3mul(R, R, Rq);
            j++;
        }
        for (int r = 0; r < 3; r++) p[r] += R[3 * r] * s[i].tip[0] + R[3 * r + 1] * s[i].tip[1] + R[3 * r + 2] * s[i].tip[2];
    }
    if (J)
        for (int c = 0; c < 12; c++) {
            const double d[3] = { p[0] - org[c][0], p[1] - org[c][1], p[2] - org[c][2] };
            J[0][c] = ax[c][1] * d[2] - ax[c][2] * d[1];
            J[1][c] = ax[c][2] * d[0] - ax[c][0] * d[2];
            J[2][c] = ax[c][0] * d[1] - ax[c][1] * d[0];
            J[3][c] = ax[c][0]; J[4][c] = ax[c][1]; J[5][c] = ax[c][2];
        }
}

/* user order <-> chain order (kdl_class.cpp:73-78 and :124-128) */
static void to_chain_order(const double q[12], double qc[12]) {
    for (int i = 0; i < 6; i++) qc[i] = q[i];
    for (int i = 11, j = 6; i >= 6; i--, j++) qc[j] = q[i];
    qc[11] = -qc[11];
}
static void from_chain_order(const double qc[12], double q[12]) {
    for (int i = 0; i < 6; i++) q[i] = qc[i];
    for (int i = 11, j = 6; i >= 6; i--, j++) q[j] = qc[i];
    q[6] = -q[6];
}

extern "C" void ckco_kdl_fk(const ckco_kin* k, const double q[12], double T[16]) {     /* kdl_class.cpp:251-278 */
    double qc[12], R[9], p[3];
    to_chain_order(q, qc);
    chain_fk(k, qc, R, p, NULL);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[4 * i + j] = R[3 * i + j]; T[4 * i + 3] = p[i]; }
    T[12] = T[13] = T[14] = 0; T[15] = 1;
}

/* KDL Rotation::GetRot() = axis * angle via GetRotAngle(axis, eps = 1e-6) */
static void kdl_get_rot(const double R[9], double rv[3]) {
    const double eps = 1e-6, eps2 = 1e-5;
    if (fabs(R[1] - R[3]) < eps && fabs(R[2] - R[6]) < eps && fabs(R[5] - R[7]) < eps) {
        if (fabs(R[1] + R[3]) < eps2 && fabs(R[2] + R[6]) < eps2 && fabs(R[5] + R[7]) < eps2 && fabs(R[0] + R[4] + R[8] - 3) < eps2) {
            rv[0] = rv[1] = rv[2] = 0;            /* identity: angle 0 */
            return;
        }
        const double angle = 3.14159265358979323846;   /* 180 degrees */
        const double xx = (R[0] + 1) / 2, yy = (R[4] + 1) / 2, zz = (R[8] + 1) / 2;
        const double xy = (R[1] + R[3]) / 4, xz = (R[2] + R[6]) / 4, yz = (R[5] + R[7]) / 4;
        double x, y, z;
        if (xx > yy && xx > zz) { x = sqrt(xx); y = xy / x; z = xz / x; }
        else if (yy > zz) { y = sqrt(yy); x = xy / y; z = yz / y; }
        else { z = sqrt(zz); x = xz / z; y = yz / z; }
        rv[0] = x * angle; rv[1] = y * angle; rv[2] = z * angle;
        return;
    }
    const double f = (R[0] + R[4] + R[8] - 1) / 2;
    double x = R[7] - R[5], y = R[2] - R[6], z = R[3] - R[1];
    const double n = sqrt(x * x + y * y + z * z);
    const double angle = atan2(n / 2, f);
    if (n > 0) { x /= n; y /= n; z /= n; }
    rv[0] = x * angle; rv[1] = y * angle; rv[2] = z * angle;
}

/* ChainIkSolverVel_pinv: dq = V S^+ U^T twist with singular values |s| < eps dropped.  The SVD is computed
 * here by one-sided Jacobi on J^T (12x6); KDL uses a Householder SVD -- same decomposition. */
static void pinv_solve(double J[6][12], const double tw[6], double eps, double dq[12]) {
    double B[12][6], W[6][6];
    for (int r = 0; r < 12; r++) for (int c = 0; c < 6; c++) B[r][c] = J[c][r];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) W[i][j] = (i == j);
    for (int sweep = 0; sweep < 150; sweep++) {
        double off = 0;
        for (int a = 0; a < 6; a++)
            for (int b = a + 1; b < 6; b++) {
                double aa = 0, bb = 0, ab = 0;
                for (int r = 0; r < 12; r++) { aa += B[r][a] * B[r][a]; bb += B[r][b] * B[r][b]; ab += B[r][a] * B[r][b]; }
                if (fabs(ab) <= 1e-300 || fabs(ab) <= 1e-17 * sqrt(aa * bb)) continue;
                off = fmax(off, fabs(ab) / sqrt(aa * bb));
                const double zeta = (bb - aa) / (2 * ab);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                const double c = 1 / sqrt(1 + t * t), s = c * t;
                for (int r = 0; r < 12; r++) { const double x = B[r][a], y = B[r][b]; B[r][a] = c * x - s * y; B[r][b] = s * x + c * y; }
                for (int r = 0; r < 6; r++) { const double x = W[r][a], y = W[r][b]; W[r][a] = c * x - s * y; W[r][b] = s * x + c * y; }
            }
        if (off < 1e-15) break;
    }
    for (int r = 0; r < 12; r++) dq[r] = 0;
    for (int c = 0; c < 6; c++) {
        double s2 = 0;
        for (int r = 0; r < 12; r++) s2 += B[r][c] * B[r][c];
        const double sig = sqrt(s2);
        if (fabs(sig) < eps) continue;
        double ut = 0;
        for (int r = 0; r < 6; r++) ut += W[r][c] * tw[r];
        const double coef = ut / s2;                /* (b_c / sig) * (1 / sig) * (w_c . tw) */
        for (int r = 0; r < 12; r++) dq[r] += B[r][c] * coef;
    }
}

/* kdl::GD (kdl_class.cpp:68-140) with ChainIkSolverPos_NR(chain, fk, vel_pinv, 10000, 1e-5)::CartToJnt */
extern "C" int ckco_gd_project(const ckco_kin* k, const double q_in[12], double q_out[12], int* converged, int* iters) {
    double qc[12];
    to_chain_order(q_in, qc);
    const double target_p[3] = { k->D, 0, 0 };       /* T_pose = Trans(D,0,0), identity rotation (kdl_class.cpp:28,85-89) */
    int ok = 0, it = 0;
    for (it = 0; it < 10000; it++) {
        double R[9], p[3], J[6][12], tw[6], dq[12];
        chain_fk(k, qc, R, p, J);
        /* delta_twist = diff(f, p_in): vel = p_in.p - f.p ; rot = f.M * GetRot(f.M^T * p_in.M), p_in.M = I */
        for (int i = 0; i < 3; i++) tw[i] = target_p[i] - p[i];
        double Rt[9], rv[3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
        kdl_get_rot(Rt, rv);
        for (int i = 0; i < 3; i++) tw[3 + i] = R[3 * i] * rv[0] + R[3 * i + 1] * rv[1] + R[3 * i + 2] * rv[2];
        pinv_solve(J, tw, 1e-5, dq);
        for (int i = 0; i < 12; i++) qc[i] += dq[i];
        int eq = 1;
        for (int i = 0; i < 6; i++) if (!(fabs(tw[i]) < 1e-5)) eq = 0;      /* Equal(delta_twist, Twist::Zero(), eps) */
        if (eq) { ok = 1; it++; break; }
    }
    if (iters) *iters = it;
    if (converged) *converged = ok;
    if (!ok) return 0;
    double q[12];
    for (int i = 0; i < 12; i++) q[i] = (fabs(qc[i]) < 1e-4) ? 0 : qc[i];          /* :110-114 */
    for (int i = 0; i < 12; i++) {                                                  /* :116-122 */
        q[i] = fmod(q[i], 2 * CKCO_PI_);
        if (q[i] > CKCO_PI_) q[i] -= 2 * CKCO_PI_;
        if (q[i] < -CKCO_PI_) q[i] += 2 * CKCO_PI_;
    }
    from_chain_order(q, q_out);                                                     /* :124-128 */
    return ckco_check_angle_limits(k, q_out);                                       /* :130 */
}

extern "C" int ckco_gd_is_valid(const ckco_kin* k, const ckco_world* w, const double q_in[12], double q_out[12]) {   /* GD.cpp:180-196 */
    if (!ckco_gd_project(k, q_in, q_out, NULL, NULL)) return 0;
    if (ckco_collision_state(w, q_out, 1, NULL, NULL)) return 0;
    return 1;
}

static int gd_rbs(const ckco_kin* k, const ckco_world* w, const double* a, const double* b, int depth, int64_t* n) {  /* GD.cpp:214-241 */
    double s = 0;
    for (int i = 0; i < 12; i++) s += pow(a[i] - b[i], 2);
    if (sqrt(s) < 0.05) return 1;
    if (depth > 150) return 0;
    double m[12], mp[12];
    for (int i = 0; i < 12; i++) m[i] = (a[i] + b[i]) / 2;
    if (n) (*n)++;
    if (!ckco_gd_is_valid(k, w, m, mp)) return 0;
    return gd_rbs(k, w, a, mp, depth + 1, n) && gd_rbs(k, w, mp, b, depth + 1, n);
}

extern "C" int ckco_gd_check_motion_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], int64_t* nchecks) {
    if (nchecks) *nchecks = 0;
    return gd_rbs(k, w, qa, qb, 0, nchecks);
}

/* check_relax_constraint StateValidityCheckerRX.cpp:105-135, Ka = (0.3/100)^2 (RX.h:230-233) */
extern "C" int ckco_rx_check(const ckco_kin* k, const double q[12], double eps, double* resid) {
    double T[16];
    ckco_kdl_fk(k, q, T);
    const double tp[3] = { k->D, 0, 0 };
    double d = 0;
    for (int i = 0; i < 3; i++) d += (T[4 * i + 3] - tp[i]) * (T[4 * i + 3] - tp[i]);
    const double roll = atan2(T[4], T[0]);
    const double pitch = atan2(-T[8], sqrt(T[9] * T[9] + T[10] * T[10]));
    const double yaw = atan2(T[9], T[10]);
    const double a = (yaw * yaw + pitch * pitch + roll * roll);
    const double epsilonD = 100, epsilonA = 0.3;
    const double Ka = (epsilonA * epsilonA) / (epsilonD * epsilonD);
    const double r = sqrt(Ka * d + a);
    if (resid) *resid = r;
    return r < eps;
}

/* ------------------------------------------------------------------------------------------------ */
/* batch driver                                                                                       */
/* ------------------------------------------------------------------------------------------------ */
extern "C" void ckco_batch_pcs_validate(const ckco_kin* k, const ckco_world* w, int64_t n, const double* q_aos,
                                        const int8_t* ac, const int8_t* ik, uint8_t* out_ok, uint8_t* out_valid,
                                        double* q_out, ckco_collide_stats* stats_sum) {
    for (int64_t i = 0; i < n; i++) {
        double q[12];
        memcpy(q, q_aos + 12 * i, sizeof(q));
        const int ok = ckco_pcs_project(k, q, ac ? ac[i] : 0, ik[i]);
        int valid = 0;
        if (ok) valid = !ckco_collision_state(w, q, stats_sum == NULL, stats_sum, NULL);
        if (out_ok) out_ok[i] = (uint8_t)ok;
        if (out_valid) out_valid[i] = (uint8_t)valid;
        if (q_out) memcpy(q_out + 12 * i, q, sizeof(q));
    }
}
