/* StateValidityCheckerPCS.cpp -- see the header.  Replaces validity_checkers/StateValidityCheckerPCS.cpp. */
#include "StateValidityCheckerPCS.h"
#include <cmath>

void StateValidityChecker::init(int env) {
    L = env == 1 ? ROD_LENGTH_ENV_I : ROD_LENGTH_ENV_II;
    Q = { {1, 0, 0, L}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1} };            /* setQ, PCS.h:136-147 */
    const int dl = 20, n = (int)(L / dl);                                          /* setP, PCS.h:150-158 */
    for (int i = 0; i <= n; i++) P.push_back({ (double)i * dl, 0, 0 });
    q_IK_solution_1.assign(6, 0); q_IK_solution_2.assign(6, 0);
    initiate_log_parameters();
    total_runtime = 0; PlanDistance = 0; final_solved = false;
}

void StateValidityChecker::initiate_log_parameters() {
    IK_counter = 0; IK_time = 0; collisionCheck_counter = 0; collisionCheck_time = 0; isValid_counter = 0;
    nodes_in_path = 0; nodes_in_trees = 0; local_connection_time = 0; local_connection_count = 0; local_connection_success_count = 0;
    sampling_time = 0; sampling_counter.assign(2, 0);
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

State StateValidityChecker::identify_state_ik(const ob::State* state, State ik) {
    State q1(6), q2(6); retrieveStateVector(state, q1, q2);
    return identify_state_ik(q1, q2, ik);
}
State StateValidityChecker::identify_state_ik(State q1, State q2, State ik) {
    double q[12]; ckc_pack(q1, q2, q); int8_t r[2] = { -1, -1 };
    ckc_identify_ik(ckc_.ctx(), 1, q, r);
    /* PCS.cpp:277, :297: a branch index that is already known is kept */
    if (ik[0] == -1) ik[0] = r[0];
    if (ik[1] == -1) ik[1] = r[1];
    return ik;
}

/* ---- validity ---- */
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
    /* post-processing only (ordered midpoints are needed): the reference's recursion with single-sample device calls */
    State q1(6), q2(6);
    iteration++;
    if (normDistanceDuo(qa1, qa2, qb1, qb2) < RBS_tol) return true;
    if (iteration > RBS_max_depth) return false;
    midpoint(qa1, qa2, qb1, qb2, q1, q2);
    if (!isValidRBS(q1, q2, active_chain, ik_sol)) return false;
    if (firstORsecond == 1) M.insert(M.begin() + last_index, join_Vectors(q1, q2));
    else M.insert(M.begin() + (++last_index), join_Vectors(q1, q2));
    int prev_size = (int)M.size();
    if (!reconstructRBS(qa1, qa2, q1, q2, active_chain, ik_sol, M, iteration, last_index, 1)) return false;
    last_index += (int)M.size() - prev_size;
    if (!reconstructRBS(q1, q2, qb1, qb2, active_chain, ik_sol, M, iteration, last_index, 2)) return false;
    return true;
}

/* ---- sampling: rejection sampler of PCS.cpp:196-259, candidates validated a batch at a time ---- */
State StateValidityChecker::sample_q() {
    const int B = 256;
    std::vector<double> q(12 * B), qo(12 * B); std::vector<int8_t> ik(B); std::vector<uint8_t> ok(B), hit(B), lim(B); std::vector<int8_t> id(2 * B);
    clock_t sT = clock();
    for (;;) {
        for (int i = 0; i < B; i++) {
            for (int j = 0; j < 6; j++) q[12 * i + j] = ((double)rand() / (RAND_MAX)) * 2 * PI_ - PI_;
            for (int j = 6; j < 12; j++) q[12 * i + j] = 0;
            ik[i] = (int8_t)(rand() % 8);
        }
        ckc_pcs_project(ckc_.ctx(), B, q.data(), nullptr, ik.data(), qo.data(), ok.data());
        ckc_identify_ik(ckc_.ctx(), B, qo.data(), id.data());
        ckc_collide(ckc_.ctx(), B, qo.data(), hit.data());
        ckc_check_angle_limits(ckc_.ctx(), B, qo.data(), lim.data());
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
bool StateValidityChecker::IKproject(State& q) {
    double qo[12]; uint8_t v = 0;
    IK_counter++;
    ckc_gd_validate(ckc_.ctx(), 1, q.data(), qo, &v);
    if (v) q.assign(qo, qo + 12);
    return v != 0;
}
bool StateValidityChecker::IKproject(const ob::State* state) {
    State q1(6), q2(6); retrieveStateVector(state, q1, q2);
    State q = join_Vectors(q1, q2);
    if (!IKproject(q)) return false;
    seperate_Vector(q, q1, q2); updateStateVector(state, q1, q2);
    return true;
}
State StateValidityChecker::sample_q_gd() {
    const int B = 1024; std::vector<double> q(12 * B), qo(12 * B); std::vector<uint8_t> v(B);
    for (;;) {
        for (int i = 0; i < 12 * B; i++) q[i] = -PI_ + (double)rand() / RAND_MAX * 2 * PI_;
        ckc_gd_validate(ckc_.ctx(), B, q.data(), qo.data(), v.data());
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
#endif
