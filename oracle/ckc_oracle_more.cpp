/*
 * ckc_oracle_more.cpp -- CPU ORACLE (test infrastructure, NOT product code): the checker members that round 2 added tests for.
 * Restates, with the reference's own control flow (paths relative to /root/reference):
 *   reconstructRBS   validity_checkers/StateValidityCheckerPCS.cpp:531-577, StateValidityCheckerGD.cpp:246-296
 *   checkMotion      StateValidityCheckerPCS.cpp:389-432 (fixed grid, breadth-first bisection order), StateValidityCheckerGD.cpp:136-175
 *   IKproject        StateValidityCheckerPCS.cpp:94-138 (5-argument form: neighbour's branch, fall back to the other active chain)
 * Pinned against the reference's prebuilt binary where it exports them (reconstructRBS, IKproject: oracle/refshim.cpp ops 13 / 14,
 * fixtures tests/golden/reconstruct_rbs.npz, ikproject5.npz); checkMotion needs OMPL objects inside the binary and is pinned through
 * isValidRBS, which it calls on every grid point.
 */
#include "ckc_oracle.h"
#include <math.h>
#include <string.h>
#include <queue>
#include <utility>
#include <vector>

typedef std::vector<std::vector<double> > Rows;

static double norm12(const double* a, const double* b) {
    double sum = 0;
    for (int i = 0; i < 6; i++) sum += pow(a[i] - b[i], 2) + pow(a[i + 6] - b[i + 6], 2);
    return sqrt(sum);
}

/* PCS.cpp:545-577 with the reference's insert bookkeeping (iteration, last_index, firstORsecond) */
static int pcs_reconstruct(const ckco_kin* k, const ckco_world* w, const double* a, const double* b, int ac, int ik, Rows& M, int iteration,
                           int last_index, int first_or_second) {
    iteration++;
    if (norm12(a, b) < 0.05) return 1;
    if (iteration > 150) return 0;
    double m[12];
    for (int i = 0; i < 12; i++) m[i] = (a[i] + b[i]) / 2;
    if (!ckco_pcs_is_valid_rbs(k, w, m, ac, ik)) return 0;
    std::vector<double> row(m, m + 12);
    if (first_or_second == 1) M.insert(M.begin() + last_index, row);
    else M.insert(M.begin() + (++last_index), row);
    const int prev_size = (int)M.size();
    if (!pcs_reconstruct(k, w, a, m, ac, ik, M, iteration, last_index, 1)) return 0;
    last_index += (int)M.size() - prev_size;
    if (!pcs_reconstruct(k, w, m, b, ac, ik, M, iteration, last_index, 2)) return 0;
    return 1;
}

static int flush(const Rows& M, double* out, int cap, int* n) {
    *n = (int)M.size();
    for (int i = 0; i < (int)M.size() && i < cap; i++) memcpy(out + 12 * i, M[i].data(), 96);
    return (int)M.size() <= cap;
}

/* reconstructRBS(s1, s2, Confs, ac, ik) PCS.cpp:531-543: Confs = {a, b}, then the recursion inserts between them */
extern "C" int ckco_pcs_reconstruct_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], int ac, int ik,
                                        double* rows, int cap, int* nrows) {
    Rows M;
    M.push_back(std::vector<double>(qa, qa + 12)); M.push_back(std::vector<double>(qb, qb + 12));
    const int ok = pcs_reconstruct(k, w, qa, qb, ac, ik, M, 0, 1, 1);
    flush(M, rows, cap, nrows);
    return ok;
}

static int gd_reconstruct(const ckco_kin* k, const ckco_world* w, const double* a, const double* b, Rows& M, int iteration, int last_index,
                          int first_or_second) {                                         /* GD.cpp:258-296 */
    iteration++;
    double s = 0;
    for (int i = 0; i < 12; i++) s += pow(a[i] - b[i], 2);
    if (sqrt(s) < 0.05) return 1;
    if (iteration > 150) return 0;
    double m[12], mp[12];
    for (int i = 0; i < 12; i++) m[i] = (a[i] + b[i]) / 2;
    if (!ckco_gd_is_valid(k, w, m, mp)) return 0;
    std::vector<double> row(mp, mp + 12);
    if (first_or_second == 1) M.insert(M.begin() + last_index, row);
    else M.insert(M.begin() + (++last_index), row);
    const int prev_size = (int)M.size();
    if (!gd_reconstruct(k, w, a, mp, M, iteration, last_index, 1)) return 0;
    last_index += (int)M.size() - prev_size;
    if (!gd_reconstruct(k, w, mp, b, M, iteration, last_index, 2)) return 0;
    return 1;
}

extern "C" int ckco_gd_reconstruct_rbs(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], double* rows, int cap,
                                       int* nrows) {
    Rows M;
    M.push_back(std::vector<double>(qa, qa + 12)); M.push_back(std::vector<double>(qb, qb + 12));
    const int ok = gd_reconstruct(k, w, qa, qb, M, 0, 1, 1);
    flush(M, rows, cap, nrows);
    return ok;
}

/* checkMotion(s1, s2, ac, ik) PCS.cpp:389-432: nd = d / RBS_tol grid, interior points visited in breadth-first bisection order
 * until the first invalid one; nchecks counts the isValid calls made (the reference's early exit included) */
extern "C" int ckco_pcs_check_motion(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], int ac, int ik, int* nchecks) {
    if (nchecks) *nchecks = 0;
    const double d = norm12(qa, qb);
    const int nd = (int)(d / 0.05);
    int result = 1;
    std::queue<std::pair<int, int> > pos;
    if (nd >= 2) {
        pos.push(std::make_pair(1, nd - 1));
        while (!pos.empty()) {
            const std::pair<int, int> x = pos.front();
            const int mid = (x.first + x.second) / 2;
            const double t = (double)mid / (double)nd;
            double q[12];
            for (int i = 0; i < 12; i++) q[i] = qa[i] + (qb[i] - qa[i]) * t;       /* RealVectorStateSpace::interpolate */
            if (nchecks) (*nchecks)++;
            if (!ckco_pcs_is_valid_rbs(k, w, q, ac, ik)) { result = 0; break; }
            pos.pop();
            if (x.first < mid) pos.push(std::make_pair(x.first, mid - 1));
            if (x.second > mid) pos.push(std::make_pair(mid + 1, x.second));
        }
    }
    return result;
}

/* checkMotion(s1, s2) GD.cpp:136-175: nd = |q1 - q2| / dq; for nd > 2 the points i / (nd - 1), i = 1 .. nd - 1, each through
 * isValid (GD projection of the interpolated point + collision) */
extern "C" int ckco_gd_check_motion(const ckco_kin* k, const ckco_world* w, const double qa[12], const double qb[12], int* nchecks) {
    if (nchecks) *nchecks = 0;
    double s = 0;
    for (int i = 0; i < 12; i++) s += pow(qa[i] - qb[i], 2);
    const int nd = (int)(sqrt(s) / 0.05);
    if (nd > 2)
        for (int i = 1; i < nd; i++) {
            const double t = (double)i / (double)(nd - 1);
            double q[12], qp[12];
            for (int j = 0; j < 12; j++) q[j] = qa[j] + (qb[j] - qa[j]) * t;
            if (nchecks) (*nchecks)++;
            if (!ckco_gd_is_valid(k, w, q, qp)) return 0;
        }
    return 1;
}

/* IKproject(q1, q2, ik, active_chain, ik_nn) PCS.cpp:94-138: the active chain's branch ik_nn[0] first, then the OTHER chain with
 * ik_nn[1] (the active chain flips), then collision.  q in/out, ik_out[2], *ac in/out; returns valid.
 * Note the reference's indexing: for active_chain = 1 it also projects with ik_nn[0] first and records it in ik[1] (:117-131). */
extern "C" int ckco_pcs_ikproject5(const ckco_kin* k, const ckco_world* w, double q[12], int ik_out[2], int* ac, const int ik_nn[2]) {
    ik_out[0] = ik_out[1] = -1;
    int valid = 1;
    double p[6];
    if (!*ac) {
        if (!ckco_project_r1(k, q, ik_nn[0], p)) {
            if (!ckco_project_r2(k, q + 6, ik_nn[1], p)) valid = 0;
            else { *ac = !*ac; memcpy(q, p, 48); ik_out[1] = ik_nn[1]; }
        } else { memcpy(q + 6, p, 48); ik_out[0] = ik_nn[0]; }
    } else {
        if (!ckco_project_r2(k, q + 6, ik_nn[0], p)) {
            if (!ckco_project_r1(k, q, ik_nn[1], p)) valid = 0;
            else { *ac = !*ac; memcpy(q + 6, p, 48); ik_out[0] = ik_nn[0]; }
        } else { memcpy(q, p, 48); ik_out[1] = ik_nn[1]; }
    }
    if (valid && ckco_collision_state(w, q, 1, NULL, NULL)) valid = 0;
    return valid;
}
