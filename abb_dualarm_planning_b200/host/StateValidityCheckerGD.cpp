/* StateValidityCheckerGD.cpp -- see the header.  Replaces validity_checkers/StateValidityChecker{GD,RX}.cpp. */
#include "StateValidityCheckerGD.h"
#include <cmath>

void StateValidityChecker::ckc_check(int rc) {
    if (rc == CKC_OK) return;
    std::cerr << "libckc: " << ckc_last_error(ckc_.ctx()) << " (" << rc << ")" << std::endl;
    exit(-1);
}
void StateValidityChecker::setQ() { Q = { {1, 0, 0, L}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1} }; }
void StateValidityChecker::setP() {
    const int dl = 20, np = (int)(L / dl);
    P.clear();
    for (int i = 0; i <= np; i++) P.push_back({ (double)i * dl, 0, 0 });
}
void StateValidityChecker::init(int env) {
    L = env == 1 ? ROD_LENGTH_ENV_I : ROD_LENGTH_ENV_II;
    setQ(); setP();
    const double D = env == 1 ? ROBOTS_DISTANCE_ENV_I : ROBOTS_DISTANCE_ENV_II;
    T_pose = { { 1, 0, 0, D }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };          /* kdl_class.cpp:28 */
    initMatrix(T_fk, 4, 4); initMatrix(T_fk_temp, 4, 4);
    q_solution.assign(12, 0); q_prev.assign(12, 0);
    initiate_log_parameters();
    total_runtime = 0; PlanDistance = 0; final_solved = false;
}
void StateValidityChecker::initiate_log_parameters() {
    IK_counter = 0; IK_time = 0; collisionCheck_counter = 0; collisionCheck_time = 0; isValid_counter = 0;
    nodes_in_path = 0; nodes_in_trees = 0; local_connection_time = 0; local_connection_count = 0; local_connection_success_count = 0;
    sampling_time = 0; sampling_counter.assign(2, 0);
}
void StateValidityChecker::FK(State q) {
    double T[16];
    ckc_check(ckc_kdl_fk(ckc_.ctx(), 1, q.data(), T));
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T_fk[i][j] = T[4 * i + j];
}
void StateValidityChecker::printMatrix(Matrix M) { for (auto& r : M) { for (double v : r) std::cout << v << " "; std::cout << std::endl; } }
void StateValidityChecker::printVector(State p) { std::cout << "["; for (double v : p) std::cout << v << " "; std::cout << "]" << std::endl; }
void StateValidityChecker::printStateVector(const ob::State* state) {
    State q(12); retrieveStateVector(state, q);
    std::cout << "q: "; printVector(q);
}
void StateValidityChecker::log_q(State q) {
    std::ofstream f("../paths/path.txt"); f << 1 << std::endl;
    for (size_t j = 0; j < q.size(); j++) f << q[j] << " ";
    f << std::endl;
}
void StateValidityChecker::log_q(State q, bool New) {
    std::ofstream f;
    if (New) { f.open("../paths/path.txt"); f << 1 << std::endl; }
    else f.open("../paths/path.txt", std::ios::app);
    for (int j = 0; j < 12; j++) f << q[j] << " ";
    f << std::endl;
}
void StateValidityChecker::LogPerf2file() {                   /* 16 lines, fixed order: GD.cpp:447-470 */
    std::ofstream f("./paths/perf_log.txt");
    f << final_solved << std::endl << PlanDistance << std::endl << total_runtime << std::endl << get_IK_counter() << std::endl << get_IK_time() << std::endl
      << get_collisionCheck_counter() << std::endl << get_collisionCheck_time() << std::endl << get_isValid_counter() << std::endl << nodes_in_path << std::endl
      << nodes_in_trees << std::endl << local_connection_time << std::endl << local_connection_count << std::endl << local_connection_success_count << std::endl
      << sampling_time << std::endl << sampling_counter[0] << std::endl << sampling_counter[1] << std::endl;
}
double StateValidityChecker::stateDistance(const ob::State* s1, const ob::State* s2) {
    State q1(12), q2(12); retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    return normDistance(q1, q2);
}
double StateValidityChecker::maxDistance(State a1, State a2) { double m = 0; for (size_t i = 0; i < a1.size(); i++) m = fabs(a1[i] - a2[i]) > m ? fabs(a1[i] - a2[i]) : m; return m; }
double StateValidityChecker::MaxAngleDistance(State a1, State a2) { return maxDistance(a1, a2); }
void StateValidityChecker::Join_States(State& q, State q1, State q2) { for (size_t i = 0; i < q1.size(); i++) { q[i] = q1[i]; q[i + q1.size()] = q2[i]; } }

/* reconstructRBS (GD.cpp:246-296, RX.cpp:263-310): the matrix of configurations the bisection visits, in order along the edge */
bool StateValidityChecker::reconstructRBS(const ob::State* s1, const ob::State* s2, Matrix& Confs) {
    State q1(n), q2(n); retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    Confs.push_back(q1); Confs.push_back(q2);
    return reconstructRBS(q1, q2, Confs, 0, 1, 1);
}

bool StateValidityChecker::GD(State q) {
    IK_counter++;
    clock_t b = clock();
    double qo[12]; uint8_t ok = 0, conv = 0; int32_t it = 0;
    ckc_gd_project(ckc_.ctx(), 1, q.data(), qo, &ok, &conv, &it);
    if (conv) q_solution.assign(qo, qo + 12);          /* q_solution is only written when the solver converged (kdl_class.cpp:108-128) */
    IK_time += double(clock() - b) / CLOCKS_PER_SEC;
    return ok != 0;
}
bool StateValidityChecker::check_angle_limits(State q) { uint8_t ok = 0; ckc_check_angle_limits(ckc_.ctx(), 1, q.data(), &ok); return ok != 0; }
int StateValidityChecker::collision_state(Matrix, State q1, State q2) {
    collisionCheck_counter++;
    clock_t b = clock();
    double q[12]; ckc_pack(q1, q2, q); uint8_t hit = 1;
    ckc_collide(ckc_.ctx(), 1, q, &hit);
    collisionCheck_time += double(clock() - b) / CLOCKS_PER_SEC;
    return hit;
}

#ifdef CKC_FLAVOUR_RX
bool StateValidityChecker::check_relax_constraint(State q) {
    IK_counter++;
    uint8_t v = 0; double r = 0;
    ckc_rx_validate(ckc_.ctx(), 1, q.data(), epsilon, &v, &r);
    return r < epsilon;
}
bool StateValidityChecker::isValidRBS(State q) {
    isValid_counter++; IK_counter++;
    uint8_t v = 0;
    ckc_rx_validate(ckc_.ctx(), 1, q.data(), epsilon, &v, nullptr);
    return v != 0;
}
bool StateValidityChecker::check_project(const ob::State* state) { State q(12); retrieveStateVector(state, q); return isValidRBS(q); }
bool StateValidityChecker::isValid(const ob::State* state) {
    State q(12); retrieveStateVector(state, q);
    if (!isValidRBS(q)) return false;
    q_prev = q; return true;
}
void StateValidityChecker::isValidBatch(std::vector<State>& q, std::vector<unsigned char>& valid) {
    const size_t m = q.size(); std::vector<double> in(12 * m); valid.assign(m, 0);
    for (size_t i = 0; i < m; i++) for (int j = 0; j < 12; j++) in[12 * i + j] = q[i][j];
    ckc_rx_validate(ckc_.ctx(), (int64_t)m, in.data(), epsilon, valid.data(), nullptr);
    isValid_counter += (int)m;
}
State StateValidityChecker::sample_q() {
    const int B = 4096; std::vector<double> q(12 * B); std::vector<uint8_t> v(B);
    int rounds = 0;
    for (;;) {
        for (int i = 0; i < 12 * B; i++) q[i] = -PI + (double)rand() / RAND_MAX * 2 * PI;
        ckc_check(ckc_rx_validate(ckc_.ctx(), B, q.data(), epsilon, v.data(), nullptr));
        for (int i = 0; i < B; i++) { if (v[i]) { sampling_counter[0]++; return State(q.begin() + 12 * i, q.begin() + 12 * i + 12); } sampling_counter[1]++; }
        if (++rounds > 100000) { std::cerr << "sample_q: no valid configuration in " << (long)rounds * B << " draws" << std::endl; exit(-1); }
    }
}
/* RX keeps every midpoint as is (no projection, RX.cpp:216-243), so the whole bisection tree of an edge is known from its end
 * points: it is generated on the host with the reference's own arithmetic ((s1 + s2) / 2 per component, the d < RBS_tol and
 * depth tests per segment) and validated in ONE batched call; the edge is valid iff every node is (the reference's depth-first
 * early return visits fewer nodes of a failing edge, the boolean is the same). */
static bool rx_tree(const State& s1, const State& s2, int depth, double tol, int max_depth, std::vector<double>& nodes, double (*dist)(const State&, const State&)) {
    if (dist(s1, s2) < tol) return true;
    if (depth > max_depth) return false;                      /* RX.cpp:225-226: this branch alone already fails the edge */
    State mid(12); for (int i = 0; i < 12; i++) mid[i] = (s1[i] + s2[i]) / 2;
    nodes.insert(nodes.end(), mid.begin(), mid.end());
    if (nodes.size() > 12u * (1u << 22)) return false;        /* 4 M nodes: an edge of > 2e5 rad, never a planner edge */
    return rx_tree(s1, mid, depth + 1, tol, max_depth, nodes, dist) & rx_tree(mid, s2, depth + 1, tol, max_depth, nodes, dist);
}
static double rx_dist(const State& a, const State& b) { double s = 0; for (size_t i = 0; i < a.size(); i++) s += pow(a[i] - b[i], 2); return sqrt(s); }

bool StateValidityChecker::checkMotionRBS(State s1, State s2, int recursion_depth, int) {
    std::vector<double> nodes;
    const bool shape_ok = rx_tree(s1, s2, recursion_depth, RBS_tol, RBS_max_depth, nodes, rx_dist);
    if (nodes.empty()) return shape_ok;
    const size_t m = nodes.size() / 12;
    std::vector<unsigned char> valid(m, 0);
    ckc_rx_validate(ckc_.ctx(), (int64_t)m, nodes.data(), epsilon, valid.data(), nullptr);
    isValid_counter += (int)m; IK_counter += (int)m;
    bool all = shape_ok;
    for (size_t i = 0; i < m; i++) all = all && valid[i] != 0;
    return all;
}
/* in-order list of the nodes of the (projection-free) bisection tree of one edge */
static bool rx_tree_ordered(const State& s1, const State& s2, int depth, double tol, int max_depth, std::vector<State>& out) {
    if (rx_dist(s1, s2) < tol) return true;
    if (depth > max_depth) return false;
    State mid(12); for (int i = 0; i < 12; i++) mid[i] = (s1[i] + s2[i]) / 2;
    if (out.size() > (1u << 20)) return false;
    const bool l = rx_tree_ordered(s1, mid, depth + 1, tol, max_depth, out);
    out.push_back(mid);
    const bool r = rx_tree_ordered(mid, s2, depth + 1, tol, max_depth, out);
    return l && r;
}
bool StateValidityChecker::reconstructRBS(State q1, State q2, Matrix& M, int iteration, int last_index, int firstORsecond) {
    std::vector<State> nodes;
    const bool shape_ok = rx_tree_ordered(q1, q2, iteration + 1, RBS_tol, RBS_max_depth, nodes);
    std::vector<unsigned char> valid;
    if (!nodes.empty()) isValidBatch(nodes, valid);
    bool all = shape_ok;
    for (unsigned char v : valid) all = all && v != 0;
    const int at = firstORsecond == 1 ? last_index : last_index + 1;
    /* the reference stops at its first invalid midpoint and leaves a partial matrix; a valid edge gives the same rows here */
    if (all) M.insert(M.begin() + at, nodes.begin(), nodes.end());
    return all;
}
bool StateValidityChecker::sample_q(State& q) {
    clock_t sT = clock();
    for (size_t i = 0; i < q.size(); i++) q[i] = -PI + (double)rand() / RAND_MAX * 2 * PI;
    const bool ok = isValidRBS(q);
    sampling_time += double(clock() - sT) / CLOCKS_PER_SEC;
    sampling_counter[ok ? 0 : 1]++;
    return ok;
}
bool StateValidityChecker::IKproject(State& q, bool includeObs) {
    if (!includeObs) return check_relax_constraint(q) && check_angle_limits(q);
    return isValidRBS(q);
}
bool StateValidityChecker::IKproject(const ob::State* state, bool includeObs) { State q(12); retrieveStateVector(state, q); return IKproject(q, includeObs); }
void StateValidityChecker::checkMotionRBSBatch(const std::vector<State>& a, const std::vector<State>& b, std::vector<unsigned char>& ok) {
    /* all trees of all edges in one call */
    std::vector<double> nodes; std::vector<size_t> first(a.size() + 1, 0); std::vector<char> shape(a.size(), 0);
    for (size_t e = 0; e < a.size(); e++) { first[e] = nodes.size() / 12; shape[e] = rx_tree(a[e], b[e], 0, RBS_tol, RBS_max_depth, nodes, rx_dist); }
    first[a.size()] = nodes.size() / 12;
    const size_t m = nodes.size() / 12;
    std::vector<unsigned char> valid(m, 0);
    if (m) ckc_rx_validate(ckc_.ctx(), (int64_t)m, nodes.data(), epsilon, valid.data(), nullptr);
    isValid_counter += (int)m; IK_counter += (int)m;
    ok.assign(a.size(), 0);
    for (size_t e = 0; e < a.size(); e++) {
        bool all = shape[e] != 0;
        for (size_t i = first[e]; i < first[e + 1]; i++) all = all && valid[i] != 0;
        ok[e] = all ? 1 : 0;
    }
}
#else
bool StateValidityChecker::IKproject(State& q, bool includeObs) {
    if (!GD(q)) return false;
    q = get_GD_result();
    State q1(6), q2(6); seperate_Vector(q, q1, q2);
    if (includeObs && withObs && collision_state(P, q1, q2)) return false;
    return true;
}
bool StateValidityChecker::IKproject(const ob::State* state, bool includeObs) {
    State q(12); retrieveStateVector(state, q);
    if (!IKproject(q, includeObs)) return false;
    updateStateVector(state, q); return true;
}
bool StateValidityChecker::isValidRBS(State& q) {
    isValid_counter++; IK_counter++;
    double qo[12]; uint8_t v = 0;
    ckc_gd_validate(ckc_.ctx(), 1, q.data(), qo, &v);
    if (v) q.assign(qo, qo + 12);
    return v != 0;
}
bool StateValidityChecker::isValid(const ob::State* state) {
    State q(12); retrieveStateVector(state, q);
    if (!isValidRBS(q)) return false;
    updateStateVector(state, q); q_prev = q; return true;
}
void StateValidityChecker::isValidBatch(std::vector<State>& q, std::vector<unsigned char>& valid) {
    const size_t m = q.size(); std::vector<double> in(12 * m), out(12 * m); valid.assign(m, 0);
    for (size_t i = 0; i < m; i++) for (int j = 0; j < 12; j++) in[12 * i + j] = q[i][j];
    ckc_gd_validate(ckc_.ctx(), (int64_t)m, in.data(), out.data(), valid.data());
    for (size_t i = 0; i < m; i++) if (valid[i]) q[i].assign(out.begin() + 12 * i, out.begin() + 12 * i + 12);
    isValid_counter += (int)m; IK_counter += (int)m;
}
State StateValidityChecker::sample_q() {                     /* GD.cpp:53-80, candidates validated a batch at a time */
    const int B = 1024; std::vector<double> q(12 * B), qo(12 * B); std::vector<uint8_t> v(B);
    clock_t sT = clock();
    int rounds = 0;
    for (;;) {
        if (++rounds > 100000) { std::cerr << "sample_q: no valid configuration in " << (long)rounds * B << " draws" << std::endl; exit(-1); }
        for (int i = 0; i < 12 * B; i++) q[i] = -PI + (double)rand() / RAND_MAX * 2 * PI;
        ckc_check(ckc_gd_validate(ckc_.ctx(), B, q.data(), qo.data(), v.data()));
        IK_counter += B;
        for (int i = 0; i < B; i++) {
            if (v[i]) { sampling_time += double(clock() - sT) / CLOCKS_PER_SEC; sampling_counter[0]++; return State(qo.begin() + 12 * i, qo.begin() + 12 * i + 12); }
            sampling_counter[1]++;
        }
    }
}
bool StateValidityChecker::checkMotionRBS(State s1, State s2, int, int) {
    uint8_t ok = 0; int32_t nc = 0;
    ckc_gd_check_motion_rbs(ckc_.ctx(), 1, s1.data(), s2.data(), &ok, &nc);
    isValid_counter += nc;
    return ok != 0;
}
bool StateValidityChecker::reconstructRBS(State q1, State q2, Matrix& M, int iteration, int last_index, int firstORsecond) {
    /* GD.cpp:258-296: each midpoint is seeded from its two projected parents, so the tree is evaluated level by level on the device
     * (ckc_gd_reconstruct_rbs) and comes back ordered along the edge; the rows between q1 and q2 go where the reference's
     * recursion inserts them (last_index for a first-half call, last_index + 1 for a second-half call) */
    std::vector<double> rows(12 * 4096); int64_t m = 0; uint8_t ok = 0;
    int rc = ckc_gd_reconstruct_rbs(ckc_.ctx(), q1.data(), q2.data(), rows.data(), 4096, &m, &ok);
    if (rc == CKC_ERR_CAPACITY) { rows.resize((size_t)m * 12); rc = ckc_gd_reconstruct_rbs(ckc_.ctx(), q1.data(), q2.data(), rows.data(), m, &m, &ok); }
    ckc_check(rc);
    (void)iteration;
    const int at = firstORsecond == 1 ? last_index : last_index + 1;
    for (int64_t i = 1; i + 1 < m; i++) M.insert(M.begin() + at + (int)(i - 1), State(rows.begin() + 12 * i, rows.begin() + 12 * i + 12));
    isValid_counter += (int)(m > 2 ? m - 2 : 0);
    return ok != 0;
}
/* sewing local connection (GD.cpp:298-376): every step starts from the previous projected point -- serial by construction,
 * one device call per step (post-processing / CBiRRT_GD_s only) */
bool StateValidityChecker::isValidSew(State& q) {
    isValid_counter++;
    if (!GD(q)) return false;
    q = get_GD_result();
    State q1(6), q2(6); seperate_Vector(q, q1, q2);
    return !(withObs && collision_state(P, q1, q2));
}
bool StateValidityChecker::checkMotionSew(const ob::State* s1, const ob::State* s2) {
    /* literal: the reference overwrites d (distance to the goal) with the length of the step just taken (GD.cpp:329), which is at
     * most RBS_tol = dq, so its loop ends after one step; and it falls off the end of the function without a return value
     * (GD.cpp:334).  Same control flow here, returning true where the reference's result is undefined. */
    State q(12), q1(12), q2(12);
    retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    double d = normDistance(q1, q2);
    while (d > dq) {
        for (size_t i = 0; i < q1.size(); i++) q[i] = q1[i] + dq / d * (q2[i] - q1[i]);
        if (!isValidSew(q) || (d = normDistance(q1, q)) > RBS_tol) return false;
        q1 = q;
    }
    return true;
}
bool StateValidityChecker::reconstructSew(const ob::State* s1, const ob::State* s2, Matrix& Confs) {
    State q1(n), q2(n); retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    return reconstructSew(q1, q2, Confs);
}
bool StateValidityChecker::reconstructSew(State q1, State q2, Matrix& M) {       /* GD.cpp:352-376, literal (same overwritten d) */
    State q(n);
    M.push_back(q1);
    double d = normDistance(q1, q2);
    while (d > RBS_tol) {
        for (size_t i = 0; i < q1.size(); i++) q[i] = q1[i] + dq / d * (q2[i] - q1[i]);
        if (!isValidSew(q) || (d = normDistance(q1, q)) > RBS_tol) return false;
        M.push_back(q);
        q1 = q;
    }
    M.push_back(q2);
    return true;
}
void StateValidityChecker::checkMotionRBSBatch(const std::vector<State>& a, const std::vector<State>& b, std::vector<unsigned char>& ok) {
    const size_t m = a.size(); std::vector<double> qa(12 * m), qb(12 * m); std::vector<int32_t> nc(m); ok.assign(m, 0);
    for (size_t i = 0; i < m; i++) for (int j = 0; j < 12; j++) { qa[12 * i + j] = a[i][j]; qb[12 * i + j] = b[i][j]; }
    ckc_gd_check_motion_rbs(ckc_.ctx(), (int64_t)m, qa.data(), qb.data(), ok.data(), nc.data());
    for (size_t i = 0; i < m; i++) isValid_counter += nc[i];
}
#endif

bool StateValidityChecker::checkMotion(const ob::State* s1, const ob::State* s2) {
    /* GD.cpp:136-175: nd = |q1 - q2| / dq, points i / (nd - 1), i = 1..nd-1, each validated independently -> one batch */
    State q1(12), q2(12); retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    const double dq = 0.05;
    const int nd = (int)(normDistance(q1, q2) / dq);
    if (nd <= 2) return true;
    std::vector<State> pts; std::vector<unsigned char> valid;
    for (int i = 1; i < nd; i++) {
        const double t = (double)i / (double)(nd - 1);
        State q(12);
        for (int j = 0; j < 12; j++) q[j] = q1[j] + (q2[j] - q1[j]) * t;
        pts.push_back(q);
    }
    isValidBatch(pts, valid);
    for (unsigned char v : valid) if (!v) return false;
    return true;
}
bool StateValidityChecker::checkMotionRBS(const ob::State* s1, const ob::State* s2) {
    State q1(12), q2(12); retrieveStateVector(s1, q1); retrieveStateVector(s2, q2);
    return checkMotionRBS(q1, q2, 0, 0);
}
void StateValidityChecker::retrieveStateVector(const ob::State* state, State& q) {
    const ob::RealVectorStateSpace::StateType* S = state->as<ob::RealVectorStateSpace::StateType>();
    for (unsigned i = 0; i < 12; i++) q[i] = S->values[i];
}
void StateValidityChecker::retrieveStateVector(const ob::State* state, State& q1, State& q2) {
    const ob::RealVectorStateSpace::StateType* S = state->as<ob::RealVectorStateSpace::StateType>();
    for (unsigned i = 0; i < 6; i++) { q1[i] = S->values[i]; q2[i] = S->values[i + 6]; }
}
void StateValidityChecker::updateStateVector(const ob::State* state, State q) {
    const ob::RealVectorStateSpace::StateType* S = state->as<ob::RealVectorStateSpace::StateType>();
    for (unsigned i = 0; i < 12; i++) S->values[i] = q[i];
}
void StateValidityChecker::seperate_Vector(State q, State& q1, State& q2) { for (size_t i = 0; i < q.size() / 2; i++) { q1[i] = q[i]; q2[i] = q[i + q.size() / 2]; } }
State StateValidityChecker::join_Vectors(State q1, State q2) { State q(q1); q.insert(q.end(), q2.begin(), q2.end()); return q; }
double StateValidityChecker::normDistance(State a1, State a2) { double s = 0; for (size_t i = 0; i < a1.size(); i++) s += pow(a1[i] - a2[i], 2); return sqrt(s); }
