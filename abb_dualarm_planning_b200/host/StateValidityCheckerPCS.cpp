/* StateValidityCheckerPCS.cpp -- see the header.  Replaces validity_checkers/StateValidityCheckerPCS.cpp. */
#include "StateValidityCheckerPCS.h"
#include <cmath>

void StateValidityChecker::setQ() { Q = { {1, 0, 0, L}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1} }; }
void StateValidityChecker::setP() {
    const int dl = 20, n = (int)(L / dl);
    P.clear();
    for (int i = 0; i <= n; i++) P.push_back({ (double)i * dl, 0, 0 });
}
void StateValidityChecker::ckc_check(int rc) {
    if (rc == CKC_OK) return;
    std::cerr << "libckc: " << ckc_last_error(ckc_.ctx()) << " (" << rc << ")" << std::endl;
    exit(-1);
}
void StateValidityChecker::init(int env) {
    L = env == 1 ? ROD_LENGTH_ENV_I : ROD_LENGTH_ENV_II;
    setQ(); setP();
    q_IK_solution_1.assign(6, 0); q_IK_solution_2.assign(6, 0);
    initMatrix(T_fk_solution_1, 4, 4); initMatrix(T_fk_solution_2, 4, 4); initMatrix(T2_, 4, 4);
    p_fk_solution_1.assign(3, 0); p_fk_solution_2.assign(3, 0);
    initMatrix(Q_IK_solutions_1, 8, 6); initMatrix(Q_IK_solutions_2, 8, 6);
    valid_IK_solutions_indices_1.assign(8, 0); valid_IK_solutions_indices_2.assign(8, 0);
    initiate_log_parameters();
    total_runtime = 0; PlanDistance = 0; final_solved = false;
}

void StateValidityChecker::initiate_log_parameters() {
    IK_counter = 0; IK_time = 0; collisionCheck_counter = 0; collisionCheck_time = 0; isValid_counter = 0;
    nodes_in_path = 0; nodes_in_trees = 0; local_connection_time = 0; local_connection_count = 0; local_connection_success_count = 0;
    sampling_time = 0; sampling_counter.assign(2, 0);
}

/* ---- single-arm kinematics (two_robots), one device call each ---- */
void StateValidityChecker::FKsolve_rob(State q, int robot_num) {
    double T[16]; const int8_t rb = (int8_t)robot_num;
    ckc_check(ckc_fk(ckc_.ctx(), 1, q.data(), &rb, T));
    Matrix& M = robot_num == 1 ? T_fk_solution_1 : T_fk_solution_2;
    State& p = robot_num == 1 ? p_fk_solution_1 : p_fk_solution_2;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M[i][j] = T[4 * i + j];
    for (int i = 0; i < 3; i++) p[i] = T[4 * i + 3];
}
bool StateValidityChecker::IKsolve_rob(Matrix T, int robot_num, int solution_num) {
    double t[16], q[6]; const int8_t rb = (int8_t)robot_num, so = (int8_t)solution_num; uint8_t ok = 0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t[4 * i + j] = T[i][j];
    ckc_check(ckc_ik(ckc_.ctx(), 1, t, &rb, &so, q, &ok));
    if (ok) (robot_num == 1 ? q_IK_solution_1 : q_IK_solution_2).assign(q, q + 6);     /* kept on failure: reference quirk 9 */
    return ok != 0;
}
int StateValidityChecker::calc_all(const Matrix& T, int robot, Matrix& sols, State& idx) {
    double t[8 * 16], q[8 * 6]; int8_t rb[8], so[8]; uint8_t ok[8];
    for (int s = 0; s < 8; s++) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t[16 * s + 4 * i + j] = T[i][j]; rb[s] = (int8_t)robot; so[s] = (int8_t)s; }
    ckc_check(ckc_ik(ckc_.ctx(), 8, t, rb, so, q, ok));
    int count = 0;
    for (int s = 0; s < 8; s++)
        if (ok[s]) {
            sols[count].assign(q + 6 * s, q + 6 * s + 6); idx[count] = s; count++;
            (robot == 1 ? q_IK_solution_1 : q_IK_solution_2).assign(q + 6 * s, q + 6 * s + 6);      /* the last feasible branch stays in the getter */
        }
    return count;
}
int StateValidityChecker::calc_all_IK_solutions_1(Matrix T) { return calc_all(T, 1, Q_IK_solutions_1, valid_IK_solutions_indices_1); }
int StateValidityChecker::calc_all_IK_solutions_2(Matrix T) { return calc_all(T, 2, Q_IK_solutions_2, valid_IK_solutions_indices_2); }
bool StateValidityChecker::IsRobotsFeasible_R1(Matrix T, State q) {
    FKsolve_rob(q, 1);
    T2_ = MatricesMult(get_FK_solution_T1(), T);
    T2_ = MatricesMult(T2_, { { -1, 0, 0, 0 }, { 0, -1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } });
    countSolutions = calc_all_IK_solutions_2(T2_);
    return countSolutions >= 1;
}
bool StateValidityChecker::IsRobotsFeasible_R2(Matrix T, State q) {
    FKsolve_rob(q, 2);
    Matrix T1 = MatricesMult(get_FK_solution_T2(), { { -1, 0, 0, 0 }, { 0, -1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } });
    T1 = MatricesMult(T1, T);
    countSolutions = calc_all_IK_solutions_1(T1);
    return countSolutions >= 1;
}
Matrix StateValidityChecker::MatricesMult(Matrix M1, Matrix M2) {
    Matrix R(4, State(4, 0.0));
    for (unsigned i = 0; i < 4; ++i) for (unsigned j = 0; j < 4; ++j) for (int k = 0; k < 4; ++k) R[i][j] += M1[i][k] * M2[k][j];
    return R;
}
bool StateValidityChecker::InvertMatrix(Matrix M, Matrix& Minv) {
    /* general 4x4 inverse by cofactors, as the reference's (apc_class.cpp:439-575); false when the determinant is 0 */
    double m[16], inv[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[4 * i + j] = M[i][j];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0) return false;
    det = 1.0 / det;
    Minv.assign(4, State(4));
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Minv[i][j] = inv[4 * i + j] * det;
    return true;
}
void StateValidityChecker::printMatrix(Matrix M) { for (auto& r : M) { for (double v : r) std::cout << v << " "; std::cout << std::endl; } }
void StateValidityChecker::printVector(State p) { std::cout << "["; for (double v : p) std::cout << v << " "; std::cout << "]" << std::endl; }
State StateValidityChecker::rand_q_ambient() {
    const double lim[6][2] = { { -165, 165 }, { -110, 110 }, { -110, 70 }, { -160, 160 }, { -120, 120 }, { -180, 180 } };
    State q(12);
    for (int r = 0; r < 2; r++)
        for (int j = 0; j < 6; j++) {
            const double lo = j == 5 ? -PI_ : deg2rad(lim[j][0]), hi = j == 5 ? PI_ : deg2rad(lim[j][1]);
            q[6 * r + j] = (double)rand() / RAND_MAX * (hi - lo) + lo;
        }
    return q;
}

/* ---- projection ---- */
bool StateValidityChecker::calc_specific_IK_solution_R1(Matrix, State q1, int IKsol) {
    double q[12] = { 0 }, qo[12]; for (int i = 0; i < 6; i++) q[i] = q1[i];
    int8_t ac = 0, ik = (int8_t)IKsol; uint8_t ok = 0;
    if (ckc_pcs_project(ckc_.ctx(), 1, q, &ac, &ik, qo, &ok) != CKC_OK) return false;
    if (ok) q_IK_solution_2.assign(qo + 6, qo + 12);          /* on failure the previous solution stays (reference quirk 9) */
    return ok != 0;
}
bool StateValidityChecker::calc_specific_IK_solution_R2(Matrix, State q2, int IKsol) {
    double q[12] = { 0 }, qo[12]; for (int i = 0; i < 6; i++) q[6 + i] = q2[i];
    int8_t ac = 1, ik = (int8_t)IKsol; uint8_t ok = 0;
    if (ckc_pcs_project(ckc_.ctx(), 1, q, &ac, &ik, qo, &ok) != CKC_OK) return false;
    if (ok) q_IK_solution_1.assign(qo, qo + 6);
    return ok != 0;
}
bool StateValidityChecker::check_angle_limits(State q) {
    uint8_t ok = 0; ckc_check_angle_limits(ckc_.ctx(), 1, q.data(), &ok); return ok != 0;
}
int StateValidityChecker::collision_state(Matrix, State q1, State q2) {
    collisionCheck_counter++;
    clock_t b = clock();
    double q[12]; ckc_pack(q1, q2, q); uint8_t hit = 1;
    ckc_collide(ckc_.ctx(), 1, q, &hit);
    collisionCheck_time += double(clock() - b) / CLOCKS_PER_SEC;
    return hit;
}

bool StateValidityChecker::IKproject(State& q1, State& q2, int active_chain, int ik_sol) {
    IK_counter++;
    clock_t sT = clock();
    double q[12], qo[12]; ckc_pack(q1, q2, q);
    int8_t ac = (int8_t)active_chain, ik = (int8_t)ik_sol; uint8_t ok = 0;
    ckc_pcs_project(ckc_.ctx(), 1, q, &ac, &ik, qo, &ok);
    if (ok) ckc_unpack(qo, q1, q2);
    IK_time += double(clock() - sT) / CLOCKS_PER_SEC;
    return ok != 0;
}

bool StateValidityChecker::IKproject(State& q1, State& q2, State& ik, int& active_chain, State ik_nn) {
    /* PCS.cpp:94-138: try the active chain's branch, else switch the active chain; then collision */
    ik = { -1, -1 };
    IK_counter++;
    bool valid = true;
    clock_t sT = clock();
    if (!active_chain) {
        if (!calc_specific_IK_solution_R1(Q, q1, (int)ik_nn[0])) {
            if (!calc_specific_IK_solution_R2(Q, q2, (int)ik_nn[1])) valid = false;
            else { active_chain = !active_chain; q1 = get_IK_solution_q1(); ik[1] = ik_nn[1]; }
        } else { q2 = get_IK_solution_q2(); ik[0] = ik_nn[0]; }
    } else {
        if (!calc_specific_IK_solution_R2(Q, q2, (int)ik_nn[0])) {
            if (!calc_specific_IK_solution_R1(Q, q1, (int)ik_nn[1])) valid = false;
            else { active_chain = !active_chain; q2 = get_IK_solution_q2(); ik[0] = ik_nn[0]; }
        } else { q1 = get_IK_solution_q1(); ik[1] = ik_nn[1]; }
    }
    IK_time += double(clock() - sT) / CLOCKS_PER_SEC;
    if (valid && withObs && collision_state(getPMatrix(), q1, q2)) valid = false;
    return valid;
}

bool StateValidityChecker::IKproject(State& q1, State& q2, int active_chain) {
    /* PCS.cpp:165-193: the 8 branches in random order; the first that projects AND is identified on both chains wins.
     * All 8 projections and identifications are one batch each. */
    std::vector<int> IKsol = { 0, 1, 2, 3, 4, 5, 6, 7 };
    for (int i = 7; i > 0; i--) std::swap(IKsol[i], IKsol[rand() % (i + 1)]);          /* std::random_shuffle */
    double q[8 * 12], qo[8 * 12]; int8_t ac[8], ik[8], id[16]; uint8_t ok[8];
    for (int s = 0; s < 8; s++) { ckc_pack(q1, q2, q + 12 * s); ac[s] = (int8_t)active_chain; ik[s] = (int8_t)IKsol[s]; }
    ckc_check(ckc_pcs_project(ckc_.ctx(), 8, q, ac, ik, qo, ok));
    ckc_check(ckc_identify_ik(ckc_.ctx(), 8, qo, id));
    IK_counter += 8;
    for (int s = 0; s < 8; s++) {
        if (!ok[s] || id[2 * s] == -1 || id[2 * s + 1] == -1) continue;
        ckc_unpack(qo + 12 * s, q1, q2);
        return true;
    }
    return false;
}
bool StateValidityChecker::close_chain(const ob::State* state, int q1_active_ik_sol) {
    /* PCS.cpp:59-92: all IK solutions of arm 2 for the rod pose of arm 1, pick the requested one, then find the branch of arm 1
     * that reproduces q1 from it (distance < 0.1) */
    State q1(6), q2(6), ik(2);
    retrieveStateVector(state, q1, q2);
    FKsolve_rob(q1, 1);
    Matrix T2 = MatricesMult(get_FK_solution_T1(), getQ());
    T2 = MatricesMult(T2, { { -1, 0, 0, 0 }, { 0, -1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } });
    if (calc_all_IK_solutions_2(T2) == 0) return false;
    q2 = get_all_IK_solutions_2(q1_active_ik_sol);
    ik[0] = get_valid_IK_solutions_indices_2(q1_active_ik_sol);
    Matrix Tinv = getQ();
    InvertMatrix(getQ(), Tinv);
    FKsolve_rob(q2, 2);
    Matrix T1 = MatricesMult(get_FK_solution_T2(), { { -1, 0, 0, 0 }, { 0, -1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } });
    T1 = MatricesMult(T1, Tinv);
    const int cnt = calc_all_IK_solutions_1(T1);
    for (int c = 0; c < cnt; c++)
        if (normDistance(get_all_IK_solutions_1(c), q1) < 1e-1) { ik[1] = get_valid_IK_solutions_indices_1(c); return true; }
    return false;
}

State StateValidityChecker::identify_state_ik(const ob::State* state, State ik) {
    State q1(6), q2(6); retrieveStateVector(state, q1, q2);
    return identify_state_ik(q1, q2, ik);
}
State StateValidityChecker::identify_state_ik(State q1, State q2, State ik) {
    clock_t sT = clock();
    double q[12]; ckc_pack(q1, q2, q); int8_t r[2] = { -1, -1 };
    ckc_check(ckc_identify_ik(ckc_.ctx(), 1, q, r));
    iden += double(clock() - sT) / CLOCKS_PER_SEC;
    /* PCS.cpp:277, :297: a branch index that is already known is kept */
    if (ik[0] == -1) ik[0] = r[0];
    if (ik[1] == -1) ik[1] = r[1];
    return ik;
}

/* ---- validity ---- */
#ifdef CKC_FLAVOUR_HB
bool StateValidityChecker::isValid(const ob::State* state) {          /* HB.cpp:212-231: the GD form (project all 12 joints, collide) */
    isValid_counter++;
    State q(12); retrieveStateVector(state, q);
    double qo[12]; uint8_t v = 0;
    IK_counter++;
    ckc_check(ckc_gd_validate(ckc_.ctx(), 1, q.data(), qo, &v));
    if (!v) return false;
    q.assign(qo, qo + 12);
    updateStateVector(state, q);
    return true;
}
#else
bool StateValidityChecker::isValid(const ob::State* state) {
    /* :323-357 -- identify both branches, re-project each arm from the other (0.05 / 0.12 rad agreement), collide both
     * reconstructed configurations: one device pass (ckc_pcs_is_valid) instead of four round trips */
    isValid_counter++;
    State q1(6), q2(6); retrieveStateVector(state, q1, q2);
    double q[12]; unsigned char v = 0;
    for (int i = 0; i < 6; i++) { q[i] = q1[i]; q[6 + i] = q2[i]; }
    if (!withObs) {
        State ik = identify_state_ik(q1, q2);
        if (ik[0] < 0 || ik[1] < 0) return false;
        if (!calc_specific_IK_solution_R1(Q, q1, (int)ik[0])) return false;
        if (normDistance(q2, get_IK_solution_q2()) > 0.5e-1) return false;
        if (!calc_specific_IK_solution_R2(Q, q2, (int)ik[1])) return false;
        return !(normDistance(q1, get_IK_solution_q1()) > 0.12);
    }
    if (ckc_pcs_is_valid(ckc_.ctx(), 1, q, &v) != CKC_OK) { std::cerr << "ckc: " << ckc_last_error(ckc_.ctx()) << std::endl; exit(-1); }
    return v != 0;
}
#endif
bool StateValidityChecker::isValid(const ob::State* state, int active_chain, int IK_sol) {
    State q1(6), q2(6); retrieveStateVector(state, q1, q2);
    return isValidRBS(q1, q2, active_chain, IK_sol);
}
bool StateValidityChecker::isValidRBS(State& q1, State& q2, int active_chain, int IK_sol) {
    isValid_counter++; IK_counter++; collisionCheck_counter++;
    double q[12], qo[12]; ckc_pack(q1, q2, q);
    int8_t ac = (int8_t)active_chain, ik = (int8_t)IK_sol; uint8_t ok = 0, valid = 0;
    ckc_pcs_validate(ckc_.ctx(), 1, q, &ac, &ik, qo, &ok, &valid);
    if (ok) ckc_unpack(qo, q1, q2);
    return valid != 0;
}
void StateValidityChecker::isValidBatch(std::vector<State>& q, const std::vector<int>& ac, const std::vector<int>& ik, std::vector<unsigned char>& valid) {
    const size_t n = q.size();
    std::vector<double> in(12 * n), out(12 * n); std::vector<int8_t> a(n), k(n); std::vector<uint8_t> ok(n);
    valid.assign(n, 0);
    for (size_t i = 0; i < n; i++) { for (int j = 0; j < 12; j++) in[12 * i + j] = q[i][j]; a[i] = (int8_t)ac[i]; k[i] = (int8_t)ik[i]; }
    ckc_pcs_validate(ckc_.ctx(), (int64_t)n, in.data(), a.data(), k.data(), out.data(), ok.data(), valid.data());
    for (size_t i = 0; i < n; i++) if (ok[i]) q[i].assign(out.begin() + 12 * i, out.begin() + 12 * i + 12);
    isValid_counter += (int)n; IK_counter += (int)n;
}

/* ---- local connection ---- */
bool StateValidityChecker::checkMotion(const ob::State* s1, const ob::State* s2, int active_chain, int ik_sol) {
    /* PCS.cpp:389-432: fixed grid nd = d / RBS_tol, interior points m / nd visited in bisection order until the first invalid
     * one; the boolean is "all interior grid points valid", so the whole grid is one batch */
    State a1(6), a2(6), b1(6), b2(6);
    retrieveStateVector(s1, a1, a2); retrieveStateVector(s2, b1, b2);
    const State a = join_Vectors(a1, a2), b = join_Vectors(b1, b2);
    const int nd = (int)(stateDistance(s1, s2) / RBS_tol);
    if (nd < 2) return true;
    std::vector<State> pts; std::vector<int> ac(nd - 1, active_chain), ik(nd - 1, ik_sol); std::vector<unsigned char> valid;
    for (int m = 1; m < nd; m++) {
        const double t = (double)m / (double)nd;
        State q(12);
        for (int j = 0; j < 12; j++) q[j] = a[j] + (b[j] - a[j]) * t;      /* RealVectorStateSpace::interpolate */
        pts.push_back(q);
    }
    isValidBatch(pts, ac, ik, valid);
    for (unsigned char v : valid) if (!v) return false;
    return true;
}

bool StateValidityChecker::checkMotionRBS(const ob::State* s1, const ob::State* s2, int active_chain, int ik_sol) {
    State qa1(6), qa2(6), qb1(6), qb2(6);
    retrieveStateVector(s1, qa1, qa2); retrieveStateVector(s2, qb1, qb2);
    return checkMotionRBS(qa1, qa2, qb1, qb2, active_chain, ik_sol, 0, 0);
}
bool StateValidityChecker::checkMotionRBS(State qa1, State qa2, State qb1, State qb2, int active_chain, int ik_sol, int recursion_depth, int) {
    /* the whole bisection tree of the edge is evaluated on the device, one level per pass */
    double a[12], b[12]; ckc_pack(qa1, qa2, a); ckc_pack(qb1, qb2, b);
    int8_t ac = (int8_t)active_chain, ik = (int8_t)ik_sol; uint8_t ok = 0; int32_t nchecks = 0;
    (void)recursion_depth;
    ckc_pcs_check_motion_rbs(ckc_.ctx(), 1, a, b, &ac, &ik, &ok, &nchecks);
    isValid_counter += nchecks;
    return ok != 0;
}
void StateValidityChecker::checkMotionRBSBatch(const std::vector<State>& a, const std::vector<State>& b, const std::vector<int>& ac, const std::vector<int>& ik,
                                               std::vector<unsigned char>& ok) {
    const size_t n = a.size();
    std::vector<double> qa(12 * n), qb(12 * n); std::vector<int8_t> c(n), k(n); std::vector<int32_t> nc(n);
    ok.assign(n, 0);
    for (size_t i = 0; i < n; i++) { for (int j = 0; j < 12; j++) { qa[12 * i + j] = a[i][j]; qb[12 * i + j] = b[i][j]; } c[i] = (int8_t)ac[i]; k[i] = (int8_t)ik[i]; }
    ckc_pcs_check_motion_rbs(ckc_.ctx(), (int64_t)n, qa.data(), qb.data(), c.data(), k.data(), ok.data(), nc.data());
    for (size_t i = 0; i < n; i++) isValid_counter += nc[i];
}

bool StateValidityChecker::reconstructRBS(const ob::State* s1, const ob::State* s2, Matrix& Confs, int active_chain, int ik_sol) {
    State qa1(6), qa2(6), qb1(6), qb2(6);
    retrieveStateVector(s1, qa1, qa2); retrieveStateVector(s2, qb1, qb2);
    Confs.push_back(join_Vectors(qa1, qa2)); Confs.push_back(join_Vectors(qb1, qb2));
    return reconstructRBS(qa1, qa2, qb1, qb2, active_chain, ik_sol, Confs, 0, 1, 1);
}
bool StateValidityChecker::reconstructRBS(State qa1, State qa2, State qb1, State qb2, int active_chain, int ik_sol, Matrix& M, int iteration,
                                          int last_index, int firstORsecond) {
    /* PCS.cpp:545-577.  The reference inserts each valid midpoint of its recursion between its two parents, so that M ends up as
     * the bisection tree's nodes in order along the edge.  The device evaluates the whole tree (ckc_pcs_reconstruct_rbs) and returns
     * the nodes already ordered: the rows strictly between qa and qb go in where the reference's recursion would have put them --
     * at last_index for a first-half call, last_index + 1 for a second-half call. */
    double a[12], b[12]; ckc_pack(qa1, qa2, a); ckc_pack(qb1, qb2, b);
    std::vector<double> rows(12 * 4096); int64_t m = 0; uint8_t ok = 0;
    int rc = ckc_pcs_reconstruct_rbs(ckc_.ctx(), a, b, active_chain, ik_sol, rows.data(), 4096, &m, &ok);
    if (rc == CKC_ERR_CAPACITY) { rows.resize((size_t)m * 12); rc = ckc_pcs_reconstruct_rbs(ckc_.ctx(), a, b, active_chain, ik_sol, rows.data(), m, &m, &ok); }
    ckc_check(rc);
    (void)iteration;
    const int at = firstORsecond == 1 ? last_index : last_index + 1;
    for (int64_t i = 1; i + 1 < m; i++) M.insert(M.begin() + at + (int)(i - 1), State(rows.begin() + 12 * i, rows.begin() + 12 * i + 12));
    isValid_counter += (int)std::max<int64_t>(0, m - 2);
    return ok != 0;
}

/* ---- sampling: rejection sampler of PCS.cpp:196-259, candidates validated a batch at a time ---- */
State StateValidityChecker::sample_q() {
    const int B = 256;
    std::vector<double> q(12 * B), qo(12 * B); std::vector<int8_t> ik(B); std::vector<uint8_t> ok(B), hit(B), lim(B); std::vector<int8_t> id(2 * B);
    clock_t sT = clock();
    int rounds = 0;
    for (;;) {
        if (++rounds > 100000) { std::cerr << "sample_q: no valid configuration in " << (long)rounds * B << " draws" << std::endl; exit(-1); }
        for (int i = 0; i < B; i++) {
            for (int j = 0; j < 6; j++) q[12 * i + j] = ((double)rand() / (RAND_MAX)) * 2 * PI_ - PI_;
            for (int j = 6; j < 12; j++) q[12 * i + j] = 0;
            ik[i] = (int8_t)(rand() % 8);
        }
        ckc_check(ckc_pcs_project(ckc_.ctx(), B, q.data(), nullptr, ik.data(), qo.data(), ok.data()));
        ckc_check(ckc_identify_ik(ckc_.ctx(), B, qo.data(), id.data()));
        ckc_check(ckc_collide(ckc_.ctx(), B, qo.data(), hit.data()));
        ckc_check(ckc_check_angle_limits(ckc_.ctx(), B, qo.data(), lim.data()));
        IK_counter += B;
        for (int i = 0; i < B; i++) {
            if (!ok[i] || (id[2 * i] == -1 && id[2 * i + 1] == -1)) { sampling_counter[1]++; continue; }
            if (withObs && hit[i] && !lim[i]) { sampling_counter[1]++; continue; }      /* the reference's `&&` (PCS.cpp:217) */
            sampling_time += double(clock() - sT) / CLOCKS_PER_SEC; sampling_counter[0]++;
            return State(qo.begin() + 12 * i, qo.begin() + 12 * i + 12);
        }
    }
}
bool StateValidityChecker::sample_q(ob::State* st) {
    State q = sample_q(), q1(6), q2(6);
    seperate_Vector(q, q1, q2);
    updateStateVector(st, q1, q2);
    return true;
}

/* ---- small helpers, as in the reference ---- */
void StateValidityChecker::retrieveStateVector(const ob::State* state, State& q1, State& q2) {
    const ob::RealVectorStateSpace::StateType* S = state->as<ob::RealVectorStateSpace::StateType>();
    for (unsigned i = 0; i < 6; i++) { q1[i] = S->values[i]; q2[i] = S->values[i + 6]; }
}
void StateValidityChecker::retrieveStateVector(const ob::State* state, State& q) {
    const ob::RealVectorStateSpace::StateType* S = state->as<ob::RealVectorStateSpace::StateType>();
    for (unsigned i = 0; i < 12; i++) q[i] = S->values[i];
}
void StateValidityChecker::updateStateVector(const ob::State* state, State q) {
    const ob::RealVectorStateSpace::StateType* S = state->as<ob::RealVectorStateSpace::StateType>();
    for (unsigned i = 0; i < 12; i++) S->values[i] = q[i];
}
void StateValidityChecker::log_q(ob::State* s) { State q1(6), q2(6); retrieveStateVector(s, q1, q2); log_q(q1, q2); }
void StateValidityChecker::updateStateVector(const ob::State* state, State q1, State q2) {
    const ob::RealVectorStateSpace::StateType* S = state->as<ob::RealVectorStateSpace::StateType>();
    for (unsigned i = 0; i < 6; i++) { S->values[i] = q1[i]; S->values[i + 6] = q2[i]; }
}
void StateValidityChecker::printStateVector(const ob::State* state) {
    State q1(6), q2(6); retrieveStateVector(state, q1, q2);
    std::cout << "q1: ["; for (double v : q1) std::cout << v << " "; std::cout << "]\nq2: ["; for (double v : q2) std::cout << v << " "; std::cout << "]" << std::endl;
}
double StateValidityChecker::normDistance(State a1, State a2) { double s = 0; for (size_t i = 0; i < a1.size(); i++) s += pow(a1[i] - a2[i], 2); return sqrt(s); }
double StateValidityChecker::stateDistance(const ob::State* s1, const ob::State* s2) {
    State a1(6), a2(6), b1(6), b2(6); retrieveStateVector(s1, a1, a2); retrieveStateVector(s2, b1, b2);
    return normDistanceDuo(a1, a2, b1, b2);
}
double StateValidityChecker::normDistanceDuo(State qa1, State qa2, State qb1, State qb2) {
    double s = 0; for (size_t i = 0; i < qa1.size(); i++) s += pow(qa1[i] - qb1[i], 2) + pow(qa2[i] - qb2[i], 2); return sqrt(s);
}
void StateValidityChecker::midpoint(State qa1, State qa2, State qb1, State qb2, State& q1, State& q2) {
    for (int i = 0; i < 6; i++) { q1[i] = (qa1[i] + qb1[i]) / 2; q2[i] = (qa2[i] + qb2[i]) / 2; }
}
State StateValidityChecker::join_Vectors(State q1, State q2) { State q(q1); q.insert(q.end(), q2.begin(), q2.end()); return q; }
void StateValidityChecker::seperate_Vector(State q, State& q1, State& q2) { for (size_t i = 0; i < q.size() / 2; i++) { q1[i] = q[i]; q2[i] = q[i + q.size() / 2]; } }
void StateValidityChecker::log_q(State q1, State q2) {
    std::ofstream f("./paths/path.txt"); f << 1 << std::endl;
    for (double v : q1) f << v << " "; for (double v : q2) f << v << " "; f << std::endl;
}
void StateValidityChecker::LogPerf2file() {                   /* 16 lines, fixed order: PCS.cpp:628-651 */
    std::ofstream f("./paths/perf_log.txt");
    f << final_solved << std::endl << PlanDistance << std::endl << total_runtime << std::endl << get_IK_counter() << std::endl << get_IK_time() << std::endl
      << get_collisionCheck_counter() << std::endl << get_collisionCheck_time() << std::endl << get_isValid_counter() << std::endl << nodes_in_path << std::endl
      << nodes_in_trees << std::endl << local_connection_time << std::endl << local_connection_count << std::endl << local_connection_success_count << std::endl
      << sampling_time << std::endl << sampling_counter[0] << std::endl << sampling_counter[1] << std::endl;
}

#ifdef CKC_FLAVOUR_HB
/* ---- hybrid flavour: GD members (StateValidityCheckerHB.cpp:109-147 + kdl_class.cpp:68-140) ---- */
bool StateValidityChecker::GD(State q) {
    IK_counter++;
    double qo[12]; uint8_t ok = 0, conv = 0; int32_t it = 0;
    ckc_gd_project(ckc_.ctx(), 1, q.data(), qo, &ok, &conv, &it);
    if (conv) q_solution.assign(qo, qo + 12);
    return ok != 0;
}
bool StateValidityChecker::IKproject(State& q, bool includeObs) {
    double qo[12]; uint8_t v = 0, conv = 0; int32_t it = 0;
    IK_counter++;
    if (includeObs && withObs) {
        ckc_check(ckc_gd_validate(ckc_.ctx(), 1, q.data(), qo, &v));
        if (v) q.assign(qo, qo + 12);
        return v != 0;
    }
    ckc_check(ckc_gd_project(ckc_.ctx(), 1, q.data(), qo, &v, &conv, &it));
    if (v) q.assign(qo, qo + 12);
    return v != 0;
}
bool StateValidityChecker::IKproject(const ob::State* state, bool includeObs) {
    State q(12); retrieveStateVector(state, q);
    if (!IKproject(q, includeObs)) return false;
    updateStateVector(state, q);
    return true;
}
void StateValidityChecker::FK(State q) {
    double T[16];
    ckc_check(ckc_kdl_fk(ckc_.ctx(), 1, q.data(), T));
    initMatrix(T_fk, 4, 4);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T_fk[i][j] = T[4 * i + j];
}
bool StateValidityChecker::checkMotion(const ob::State* s1, const ob::State* s2) {
    /* HB.cpp:233-271: nd = |q1 - q2| / dq, points i / (nd - 1), i = 1 .. nd - 1, each validated independently -> one batch */
    State q1(12), q2(12); retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    const int nd = (int)(normDistance(q1, q2) / 0.05);
    if (nd <= 2) return true;
    std::vector<double> in(12 * (size_t)(nd - 1)), out(12 * (size_t)(nd - 1)); std::vector<uint8_t> valid((size_t)(nd - 1));
    for (int i = 1; i < nd; i++) { const double t = (double)i / (double)(nd - 1); for (int j = 0; j < 12; j++) in[12 * (size_t)(i - 1) + j] = q1[j] + (q2[j] - q1[j]) * t; }
    ckc_check(ckc_gd_validate(ckc_.ctx(), nd - 1, in.data(), out.data(), valid.data()));
    isValid_counter += nd - 1; IK_counter += nd - 1;
    for (uint8_t v : valid) if (!v) return false;
    return true;
}
bool StateValidityChecker::isValidSew(State& q) {
    isValid_counter++;
    if (!GD(q)) return false;
    q = get_GD_result();
    State q1(6), q2(6); seperate_Vector(q, q1, q2);
    return !(withObs && collision_state(P, q1, q2));
}
bool StateValidityChecker::checkMotionSew(const ob::State* s1, const ob::State* s2) {      /* literal, see StateValidityCheckerGD.cpp */
    State q(12), q1(12), q2(12);
    retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    double d = normDistance(q1, q2);
    while (d > 0.05) {
        for (size_t i = 0; i < q1.size(); i++) q[i] = q1[i] + 0.05 / d * (q2[i] - q1[i]);
        if (!isValidSew(q) || (d = normDistance(q1, q)) > RBS_tol) return false;
        q1 = q;
    }
    return true;
}
bool StateValidityChecker::reconstructSew(const ob::State* s1, const ob::State* s2, Matrix& Confs) {
    State q1(12), q2(12); retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    return reconstructSew(q1, q2, Confs);
}
bool StateValidityChecker::reconstructSew(State q1, State q2, Matrix& M) {
    State q(12);
    M.push_back(q1);
    double d = normDistance(q1, q2);
    while (d > RBS_tol) {
        for (size_t i = 0; i < q1.size(); i++) q[i] = q1[i] + 0.05 / d * (q2[i] - q1[i]);
        if (!isValidSew(q) || (d = normDistance(q1, q)) > RBS_tol) return false;
        M.push_back(q);
        q1 = q;
    }
    M.push_back(q2);
    return true;
}
double StateValidityChecker::maxDistance(State a1, State a2) { double m = 0; for (size_t i = 0; i < a1.size(); i++) m = fabs(a1[i] - a2[i]) > m ? fabs(a1[i] - a2[i]) : m; return m; }
double StateValidityChecker::MaxAngleDistance(State a1, State a2) { return maxDistance(a1, a2); }
void StateValidityChecker::Join_States(State& q, State q1, State q2) { for (size_t i = 0; i < q1.size(); i++) { q[i] = q1[i]; q[i + q1.size()] = q2[i]; } }
State StateValidityChecker::sample_q_gd() {
    const int B = 1024; std::vector<double> q(12 * B), qo(12 * B); std::vector<uint8_t> v(B);
    int rounds = 0;
    for (;;) {
        if (++rounds > 100000) { std::cerr << "sample_q: no valid configuration in " << (long)rounds * B << " draws" << std::endl; exit(-1); }
        for (int i = 0; i < 12 * B; i++) q[i] = -PI_ + (double)rand() / RAND_MAX * 2 * PI_;
        ckc_check(ckc_gd_validate(ckc_.ctx(), B, q.data(), qo.data(), v.data()));
        IK_counter += B;
        for (int i = 0; i < B; i++) { if (v[i]) { sampling_counter[0]++; return State(qo.begin() + 12 * i, qo.begin() + 12 * i + 12); } sampling_counter[1]++; }
    }
}
#endif

#ifdef CKC_FLAVOUR_RSS
/* ---- RSS flavour ---- */
bool StateValidityChecker::sampleSingular(ob::State* state) {
    State q1(6), q2(6), q2_mode(6), q2_add(6);
    const int sMode = rand() % 3;                                  /* RSS.cpp:260-275 */
    switch (sMode) {
    case 0: q2_mode = { 1, 1, 0, 1, 0, 1 }; q2_add = { 0, 0, -PI_ / 2, 0, 0, 0 }; break;
    case 1: q2_mode = { 1, 1, 1, 1, 0, 1 }; q2_add = { 0, 0, 0, 0, 0, 0 }; break;
    default: q2_mode = { 1, 1, 0, 1, 1, 1 }; q2_add = { 0, 0, -PI_ / 2, 0, 0, 0 };
    }
    for (;;) {
        for (int i = 0; i < 6; i++) {
            q1[i] = ((double)rand() / (RAND_MAX)) * 2 * PI_ - PI_;
            q2[i] = ((double)rand() / (RAND_MAX)) * 2 * PI_ - PI_;
            q2[i] = q2[i] * q2_mode[i] + q2_add[i];               /* make the passive chain's sample singular */
        }
        /* IsRobotsFeasible_R2 + calc_all_IK_solutions_1 (apc_class.cpp:249-262, :407-421): the 8 branches in one batch */
        double q[8 * 12], qo[8 * 12]; int8_t ac[8], ik[8]; uint8_t ok[8];
        for (int s = 0; s < 8; s++) { for (int j = 0; j < 6; j++) { q[12 * s + j] = 0; q[12 * s + 6 + j] = q2[j]; } ac[s] = 1; ik[s] = (int8_t)s; }
        ckc_pcs_project(ckc_.ctx(), 8, q, ac, ik, qo, ok);
        IK_counter += 8;
        int feas[8]; countSolutions = 0;
        for (int s = 0; s < 8; s++) if (ok[s]) feas[countSolutions++] = s;
        if (countSolutions == 0) continue;
        const int pick = feas[rand() % countSolutions];
        q1.assign(qo + 12 * pick, qo + 12 * pick + 6);
        updateStateVector(state, q1, q2);
        State idk = identify_state_ik(state);
        if (idk[1] == -1) continue;
        if (!collision_state(getPMatrix(), q1, q2)) return true;
    }
}
bool StateValidityChecker::checkMotionRBS(const ob::State* s1, const ob::State* s2, int sg, int active_chain, int ik_sol) {
    State qa1(6), qa2(6), qb1(6), qb2(6);
    retrieveStateVector(s1, qa1, qa2); retrieveStateVector(s2, qb1, qb2);
    if (sg == 1) return checkMotionRBS(qa1, qa2, qb1, qb2, qa1, qa2, active_chain, ik_sol, 0, 0);
    return checkMotionRBS(qa1, qa2, qb1, qb2, qb1, qb2, active_chain, ik_sol, 0, 0);
}
bool StateValidityChecker::checkMotionRBS(State qa1, State qa2, State qb1, State qb2, State qsg1, State qsg2, int active_chain, int ik_sol,
                                          int recursion_depth, int ndc) {
    /* RSS.cpp:586-613.  The singular stop rule only exists at this level: the two recursive calls of the reference go to the
     * plain 8-argument overload, i.e. two ordinary RBS edges -- submitted as one batch of two */
    State q1(6), q2(6);
    if (normDistanceDuo(qa1, qa2, qb1, qb2) < RBS_tol) return true;
    if (recursion_depth > RBS_max_depth) return false;
    midpoint(qa1, qa2, qb1, qb2, q1, q2);
    if (!isValidRBS(q1, q2, active_chain, ik_sol)) return false;
    if (normDistanceDuo(q1, q2, qsg1, qsg2) < 15 * RBS_tol) return true;
    std::vector<State> a = { join_Vectors(qa1, qa2), join_Vectors(q1, q2) }, b = { join_Vectors(q1, q2), join_Vectors(qb1, qb2) };
    std::vector<unsigned char> ok;
    checkMotionRBSBatch(a, b, { active_chain, active_chain }, { ik_sol, ik_sol }, ok);
    (void)ndc;
    return ok[0] && ok[1];
}
bool StateValidityChecker::reconstructRBS(const ob::State* s1, const ob::State* s2, Matrix& Confs, int sg, int active_chain, int ik_sol) {
    State qa1(6), qa2(6), qb1(6), qb2(6);
    retrieveStateVector(s1, qa1, qa2); retrieveStateVector(s2, qb1, qb2);
    Confs.push_back(join_Vectors(qa1, qa2)); Confs.push_back(join_Vectors(qb1, qb2));
    if (sg == 1) return reconstructRBS(qa1, qa2, qb1, qb2, qa1, qa2, active_chain, ik_sol, Confs, 0, 1, 1);
    return reconstructRBS(qa1, qa2, qb1, qb2, qb1, qb2, active_chain, ik_sol, Confs, 0, 1, 1);
}
bool StateValidityChecker::reconstructRBS(State qa1, State qa2, State qb1, State qb2, State qsg1, State qsg2, int active_chain, int ik_sol, Matrix& M,
                                          int iteration, int last_index, int firstORsecond) {
    /* RSS.cpp:702-741: the singular stop rule only exists at this level (a midpoint within 15 RBS_tol of the singular end point
     * ends the edge WITHOUT being inserted); the two halves are ordinary reconstructions */
    State q1(6), q2(6);
    iteration++;
    if (normDistanceDuo(qa1, qa2, qb1, qb2) < RBS_tol) return true;
    if (iteration > RBS_max_depth) return false;
    midpoint(qa1, qa2, qb1, qb2, q1, q2);
    if (!isValidRBS(q1, q2, active_chain, ik_sol)) return false;
    if (normDistanceDuo(q1, q2, qsg1, qsg2) < 15 * RBS_tol) return true;
    if (firstORsecond == 1) M.insert(M.begin() + last_index, join_Vectors(q1, q2));
    else M.insert(M.begin() + (++last_index), join_Vectors(q1, q2));
    int prev_size = (int)M.size();
    if (!reconstructRBS(qa1, qa2, q1, q2, active_chain, ik_sol, M, iteration, last_index, 1)) return false;
    last_index += (int)M.size() - prev_size;
    return reconstructRBS(q1, q2, qb1, qb2, active_chain, ik_sol, M, iteration, last_index, 2);
}
#endif
