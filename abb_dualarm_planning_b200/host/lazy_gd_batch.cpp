/*
 * lazy_gd_batch.cpp -- three more callers of the hot path (BASELINE configs[3] names SBL_GD): the reference's
 * GD-flavour single-tree and lazy planners restated without OMPL and driven over the C ABI of include/ckc.h,
 *     rrt      planners/RRT_GD.cpp      :100-262  goal-biased RRT: project the step (kdl::GD + limits + collision), RBS edge check, add
 *     lazyrrt  planners/LazyRRT_GD.cpp  :120-262  the same tree grown WITHOUT edge checks; when the goal is reached the unchecked
 *                                                 edges of the candidate path are validated, the first invalid one drops its subtree
 *     sbl      planners/SBL_GD.cpp      :107-420  bidirectional lazy planner over a grid of a random 3-D linear projection: expand a
 *                                                 motion picked with probability ~ 1 / (motions in its cell), connect through a shared
 *                                                 cell, validate both root paths lazily
 * as ensembles of independent queries stepped in lock-step (the pending question of every live query goes into ONE batched call per
 * kind), or one query at a time (workers = 0: single-query latency).  Batching INSIDE a query: the lazy planners validate all unchecked
 * edges of a candidate path in one ckc_gd_check_motion_rbs call (the reference checks them one by one from the root and stops at the
 * first failure; the edges behind a failure are removed with the subtree either way, so the tree ends up the same).
 * What cannot be restated bit for bit is OMPL's random stream and its default projection (RealVectorRandomLinearProjectionEvaluator);
 * the control flow, the questions asked and their order per query are the reference's.  Every returned path is re-validated by the CPU
 * oracle (tests/test_gpu_planner.py).
 *
 * usage: lazy_gd_batch <rrt|lazyrrt|sbl> <env> <n_queries> <max_step> <seed> <out.txt> [workers (0 = sequential)] [max_rounds]
 * out.txt: per solved query "query k nodes m rounds r" + m rows "q[12] 0 -1 -1 -1 -1 e_dir" (e_dir: the edge to the NEXT row was
 *          checked 0 = from this row toward the next, 1 = from the next row toward this one; -1 on the last row), then "summary ...".
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <thread>
#include <vector>
#include "../../include/ckc.h"

static const double kPI_ = 3.1416;
/* run/plan_GD.cpp: bounds of the sampled state space (the same box as plan_PCS.cpp:91-114) */
static const double kLo[12] = { -2.88, -1.919, -1.919, -2.79, -2.09, -kPI_, -2.88, -1.919, -1.919, -2.79, -2.09, -kPI_ };
static const double kHi[12] = { 2.88, 1.919, 1.22, 2.79, 2.09, kPI_, 2.88, 1.919, 1.22, 2.79, 2.09, kPI_ };
static const double kStart1[12] = { 0.5236, 1.7453, -1.8326, -1.4835, 1.5708, 0, 1.004278, 0.2729, 0.9486, -1.15011, 1.81001, -1.97739 };
static const double kGoal1[12] = { 0.5236, 0.34907, 0.69813, -1.3963, 1.5708, 0, 0.7096, 1.8032, -1.7061, -1.6286, 1.9143, -2.0155 };
static const double kStart2[12] = { 1.1, 1.1, 0, 1.24, -1.5708, 0, -0.79567, 0.60136, 0.43858, -0.74986, -1.0074, -0.092294 };
static const double kGoal2[12] = { -1.1, 1.35, -0.2, -1, -1.9, 0, 0.80875, 0.72363, -0.47891, -1.0484, 0.73278, 1.7491 };

enum Planner { P_RRT, P_LAZYRRT, P_SBL };
static double uni(std::mt19937_64& g) { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }
static double dist12(const double* a, const double* b) { double s = 0; for (int j = 0; j < 12; j++) s += (a[j] - b[j]) * (a[j] - b[j]); return std::sqrt(s); }

struct Motion {
    double q[12];
    int parent = -1;
    bool valid = false, alive = true;       /* valid: the edge parent -> this has passed checkMotionRBS (lazy planners) */
    int root = 0;
    std::vector<int> children;
};

/* SBL's grid over a random linear projection of the 12-D state (SBL_GD.h:120-160; OMPL's default projection of a RealVector space) */
struct Projection {
    double M[3][12], cell[3];
    void init() {
        std::mt19937_64 g(20171001ull);
        std::normal_distribution<double> nd;
        for (int r = 0; r < 3; r++) {
            for (int j = 0; j < 12; j++) M[r][j] = nd(g);
            for (int p = 0; p < r; p++) { double d = 0; for (int j = 0; j < 12; j++) d += M[r][j] * M[p][j]; for (int j = 0; j < 12; j++) M[r][j] -= d * M[p][j]; }
            double n = 0; for (int j = 0; j < 12; j++) n += M[r][j] * M[r][j];
            n = std::sqrt(n); for (int j = 0; j < 12; j++) M[r][j] /= n;
        }
        double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
        for (int s = 0; s < 1000; s++) {            /* inferCellSizes: extent of the projected bounds / 20 */
            double q[12], p[3]; for (int j = 0; j < 12; j++) q[j] = kLo[j] + uni(g) * (kHi[j] - kLo[j]);
            project(q, p);
            for (int r = 0; r < 3; r++) { lo[r] = std::min(lo[r], p[r]); hi[r] = std::max(hi[r], p[r]); }
        }
        for (int r = 0; r < 3; r++) cell[r] = (hi[r] - lo[r]) / 20.0;
    }
    void project(const double* q, double* p) const { for (int r = 0; r < 3; r++) { p[r] = 0; for (int j = 0; j < 12; j++) p[r] += M[r][j] * q[j]; } }
    int64_t coord(const double* q) const {
        double p[3]; project(q, p);
        int64_t c = 0;
        for (int r = 0; r < 3; r++) c = c * 4096 + (int64_t)(std::floor(p[r] / cell[r]) + 2048);
        return c;
    }
};

struct Tree {
    std::vector<Motion> m;
    std::map<int64_t, std::vector<int> > grid;       /* SBL: cell -> motions */
    int size = 0;
    int nearest(const double* q) const {
        int best = -1; double bd = 1e300;
        for (size_t i = 0; i < m.size(); i++) { if (!m[i].alive) continue; const double d = dist12(m[i].q, q); if (d < bd) { bd = d; best = (int)i; } }
        return best;
    }
};

enum Wait { W_NONE, W_PROJECT, W_RBS, W_PATH };

struct Query {
    Tree t[2];
    std::mt19937_64 rng;
    bool solved = false, start_turn = true;
    int rounds = 0;
    Wait wait = W_NONE;
    /* the step in flight */
    int tree = 0, nm = -1;
    double dstate[12];
    bool goal_direct = false;
    /* lazy path validation in flight: edges (child motion ids) in root -> leaf order, per tree */
    std::vector<int> chk[2];
    int connect_self = -1, connect_other = -1;       /* SBL: the two ends of the candidate connection */
    std::vector<double> path; std::vector<int> path_dir;
    bool junction_from_start = true;                 /* SBL: the junction edge was checked from the start tree's side */
};

struct Ensemble {
    ckc_ctx* ctx = nullptr;
    Planner planner = P_RRT;
    double max_step = 0.6, goal_bias = 0.05;
    const double* start = nullptr; const double* goal = nullptr;
    Projection proj;
    std::vector<Query> qs;
    long calls = 0, items = 0;

    int add_motion(Query& Q, int tr, const double* q, int parent, bool valid) {
        Motion mo; memcpy(mo.q, q, 96); mo.parent = parent; mo.valid = valid;
        Tree& T = Q.t[tr];
        T.m.push_back(mo);
        const int id = (int)T.m.size() - 1;
        if (parent >= 0) T.m[parent].children.push_back(id);
        if (planner == P_SBL) T.grid[proj.coord(q)].push_back(id);
        T.size++;
        return id;
    }
    void remove_motion(Query& Q, int tr, int id) {          /* SBL_GD.cpp:332-384 / LazyRRT removeMotion: the motion and its subtree */
        Tree& T = Q.t[tr];
        Motion& mo = T.m[id];
        if (!mo.alive) return;
        mo.alive = false; T.size--;
        if (planner == P_SBL) {
            auto it = T.grid.find(proj.coord(mo.q));
            if (it != T.grid.end()) {
                auto& v = it->second; v.erase(std::remove(v.begin(), v.end(), id), v.end());
                if (v.empty()) T.grid.erase(it);
            }
        }
        if (mo.parent >= 0) { auto& c = T.m[mo.parent].children; c.erase(std::remove(c.begin(), c.end(), id), c.end()); }
        const std::vector<int> kids = mo.children;
        for (int k : kids) { T.m[k].parent = -1; remove_motion(Q, tr, k); }
    }
    int select_motion(Query& Q, int tr) {                   /* SBL_GD.cpp:415-419: cell ~ 1 / (motions in it), then uniform inside */
        Tree& T = Q.t[tr];
        double tot = 0;
        for (auto& c : T.grid) tot += 1.0 / c.second.size();
        double u = uni(Q.rng) * tot;
        for (auto& c : T.grid) { u -= 1.0 / c.second.size(); if (u <= 0) return c.second[(size_t)(uni(Q.rng) * c.second.size()) % c.second.size()]; }
        return T.grid.empty() ? -1 : T.grid.rbegin()->second[0];
    }
    void unchecked_edges(const Tree& T, int leaf, std::vector<int>& out) {       /* root -> leaf order */
        out.clear();
        for (int k = leaf; k >= 0; k = T.m[k].parent) if (!T.m[k].valid && T.m[k].parent >= 0) out.push_back(k);
        std::reverse(out.begin(), out.end());
    }
    void finish(Query& Q, int leaf0, int leaf1) {           /* path: start root -> leaf0 [-> leaf1 -> goal root] */
        std::vector<int> p0, p1;
        for (int k = leaf0; k >= 0; k = Q.t[0].m[k].parent) p0.push_back(k);
        std::reverse(p0.begin(), p0.end());
        for (int k = leaf1; k >= 0; k = Q.t[1].m[k].parent) p1.push_back(k);
        Q.path.clear(); Q.path_dir.clear();
        for (int k : p0) { Q.path.insert(Q.path.end(), Q.t[0].m[k].q, Q.t[0].m[k].q + 12); Q.path_dir.push_back(0); }
        for (int k : p1) { Q.path.insert(Q.path.end(), Q.t[1].m[k].q, Q.t[1].m[k].q + 12); Q.path_dir.push_back(1); }
        /* the junction row of SBL: leaf0's copy of the other tree's state was checked from the start side */
        Q.solved = true;
    }

    /* host logic of one query until it has a question for the device */
    void advance(Query& Q) {
        while (!Q.solved && Q.wait == W_NONE) {
            if (planner == P_SBL) {
                Q.tree = Q.start_turn ? 0 : 1; Q.start_turn = !Q.start_turn;
                const int ex = select_motion(Q, Q.tree);
                if (ex < 0) continue;
                const double* e = Q.t[Q.tree].m[ex].q;
                for (int j = 0; j < 12; j++) {              /* sampleUniformNear: the box around the motion, clamped to the bounds */
                    const double lo = std::max(kLo[j], e[j] - max_step), hi = std::min(kHi[j], e[j] + max_step);
                    Q.dstate[j] = lo + uni(Q.rng) * (hi - lo);
                }
                Q.nm = ex; Q.wait = W_PROJECT;
                return;
            }
            /* RRT / LazyRRT: goal-biased sample, nearest, step (RRT_GD.cpp:148-172) */
            double r[12]; bool gg = false;
            if (uni(Q.rng) < goal_bias) { memcpy(r, goal, 96); gg = true; }
            else for (int j = 0; j < 12; j++) r[j] = kLo[j] + uni(Q.rng) * (kHi[j] - kLo[j]);
            Q.nm = Q.t[0].nearest(r);
            const double* n = Q.t[0].m[Q.nm].q;
            const double d = dist12(n, r);
            bool reach = true;
            if (d > max_step) { for (int j = 0; j < 12; j++) Q.dstate[j] = n[j] + (r[j] - n[j]) * (max_step / d); reach = false; }
            else memcpy(Q.dstate, r, 96);
            Q.goal_direct = gg && reach;
            Q.tree = 0;
            Q.wait = Q.goal_direct ? (planner == P_RRT ? W_RBS : W_NONE) : W_PROJECT;
            if (Q.wait == W_NONE) after_projection(Q);       /* LazyRRT reaching the goal directly: add it, then the lazy path check */
        }
    }
    /* the step's state is on the manifold (projected, or the goal itself) */
    void after_projection(Query& Q) {
        if (planner == P_RRT) { Q.wait = W_RBS; return; }
        if (planner == P_LAZYRRT) {
            const int id = add_motion(Q, 0, Q.dstate, Q.nm, false);
            Q.wait = W_NONE;
            if (Q.goal_direct) { unchecked_edges(Q.t[0], id, Q.chk[0]); Q.chk[1].clear(); Q.connect_self = id; Q.connect_other = -1; Q.wait = W_PATH; if (Q.chk[0].empty()) path_checked(Q, nullptr); }
            return;
        }
        /* SBL: add, then checkSolution (SBL_GD.cpp:224-300): a motion of the other tree in the same cell -> candidate connection */
        const int tr = Q.tree, ot = 1 - tr;
        const int id = add_motion(Q, tr, Q.dstate, Q.nm, false);
        Q.wait = W_NONE;
        auto it = Q.t[ot].grid.find(proj.coord(Q.dstate));
        if (it == Q.t[ot].grid.end() || it->second.empty()) return;
        const int co = it->second[(size_t)(uni(Q.rng) * it->second.size()) % it->second.size()];
        const int cn = add_motion(Q, tr, Q.t[ot].m[co].q, id, false);          /* connect: a copy of the other tree's state under the new motion */
        unchecked_edges(Q.t[tr], cn, Q.chk[tr]); unchecked_edges(Q.t[ot], co, Q.chk[ot]);
        Q.connect_self = cn; Q.connect_other = co;
        Q.wait = W_PATH;
        if (Q.chk[0].empty() && Q.chk[1].empty()) path_checked(Q, nullptr);
    }
    /* results of the lazy validation: res = flags of chk[tree] ++ chk[other] in that order (nullptr: nothing to check) */
    void path_checked(Query& Q, const uint8_t* res) {
        Q.wait = W_NONE;
        const int tr = planner == P_SBL ? Q.tree : 0, ot = 1 - tr;
        const int order[2] = { tr, ot };
        size_t o = 0; bool ok = true;
        for (int s = 0; s < 2 && ok; s++) {
            const int tt = order[s];
            for (size_t i = 0; i < Q.chk[tt].size(); i++, o++) {
                if (res[o]) { Q.t[tt].m[Q.chk[tt][i]].valid = true; continue; }
                remove_motion(Q, tt, Q.chk[tt][i]);         /* the first invalid edge from the root drops its subtree (isPathValid :318-322) */
                ok = false; break;
            }
        }
        if (!ok) return;
        if (planner == P_SBL) {
            const int self_parent = Q.t[tr].m[Q.connect_self].parent;          /* the copy itself is dropped from the path: its state is connect_other's */
            if (tr == 0) finish(Q, self_parent, Q.connect_other); else finish(Q, Q.connect_other, self_parent);
            /* direction of the junction edge: it was checked from the expanding tree's side */
            Q.junction_from_start = tr == 0;
        } else finish(Q, Q.connect_self, -1);
    }

    void run(int max_rounds) {
        std::vector<int> idx; std::vector<double> q, qo, qa, qb; std::vector<uint8_t> valid, res; std::vector<int32_t> nchk;
        for (int round = 0; round < max_rounds; round++) {
            int live = 0;
            for (auto& Q : qs) { if (Q.solved) continue; live++; advance(Q); if (!Q.solved) Q.rounds++; }
            if (!live) break;
            /* ---- IKproject(state) = kdl::GD + joint limits + collision, one batch ---- */
            idx.clear();
            for (size_t k = 0; k < qs.size(); k++) if (!qs[k].solved && qs[k].wait == W_PROJECT) idx.push_back((int)k);
            if (!idx.empty()) {
                const size_t m = idx.size();
                q.resize(m * 12); qo.resize(m * 12); valid.resize(m);
                for (size_t i = 0; i < m; i++) memcpy(&q[12 * i], qs[idx[i]].dstate, 96);
                check(ckc_gd_validate(ctx, (int64_t)m, q.data(), qo.data(), valid.data()), m);
                for (size_t i = 0; i < m; i++) {
                    Query& Q = qs[idx[i]];
                    if (!valid[i]) { Q.wait = W_NONE; continue; }
                    memcpy(Q.dstate, &qo[12 * i], 96);
                    after_projection(Q);
                }
            }
            /* ---- edge checks: RRT's single edge and the lazy planners' path edges, all queries in one call ---- */
            qa.clear(); qb.clear();
            std::vector<std::pair<int, int> > owner;          /* (query, number of edges) */
            for (size_t k = 0; k < qs.size(); k++) {
                Query& Q = qs[k];
                if (Q.solved) continue;
                if (Q.wait == W_RBS) {
                    const double* n = Q.t[0].m[Q.nm].q;
                    qa.insert(qa.end(), n, n + 12); qb.insert(qb.end(), Q.dstate, Q.dstate + 12); owner.push_back({ (int)k, 1 });
                } else if (Q.wait == W_PATH) {
                    const int tr = planner == P_SBL ? Q.tree : 0; const int order[2] = { tr, 1 - tr };
                    int cnt = 0;
                    for (int s = 0; s < 2; s++) for (int id : Q.chk[order[s]]) {
                        const Motion& mo = Q.t[order[s]].m[id];
                        const double* p = Q.t[order[s]].m[mo.parent].q;
                        qa.insert(qa.end(), p, p + 12); qb.insert(qb.end(), mo.q, mo.q + 12); cnt++;      /* checkMotionRBS(parent, motion) */
                    }
                    owner.push_back({ (int)k, cnt });
                }
            }
            if (!owner.empty()) {
                const size_t m = qa.size() / 12;
                res.resize(m); nchk.resize(m);
                if (m) check(ckc_gd_check_motion_rbs(ctx, (int64_t)m, qa.data(), qb.data(), res.data(), nchk.data()), m);
                size_t o = 0;
                for (auto& ow : owner) {
                    Query& Q = qs[ow.first];
                    if (Q.wait == W_RBS) {
                        Q.wait = W_NONE;
                        if (res[o]) {
                            const int id = add_motion(Q, 0, Q.dstate, Q.nm, true);
                            if (Q.goal_direct) finish(Q, id, -1);                 /* goal->isSatisfied: the goal state itself was added */
                        }
                    } else path_checked(Q, res.data() + o);
                    o += (size_t)ow.second;
                }
            }
        }
    }
    void check(int rc, size_t m) {
        if (rc != CKC_OK) { fprintf(stderr, "ckc: %s\n", ckc_last_error(ctx)); exit(-1); }
        calls++; items += (long)m;
    }
};

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s <rrt|lazyrrt|sbl> <env> <n_queries> <max_step> <seed> <out.txt> [workers (0 = sequential)] [max_rounds]\n", argv[0]); return 2; }
    const Planner pl = !strcmp(argv[1], "rrt") ? P_RRT : (!strcmp(argv[1], "lazyrrt") ? P_LAZYRRT : P_SBL);
    const int env = atoi(argv[2]), nq = atoi(argv[3]); const double max_step = atof(argv[4]); const uint64_t seed = strtoull(argv[5], nullptr, 10);
    const bool sequential = argc > 7 && atoi(argv[7]) == 0;
    const int workers = std::max(1, std::min(argc > 7 ? atoi(argv[7]) : 1, std::max(nq, 1)));
    const int max_rounds = argc > 8 ? atoi(argv[8]) : 200000;
    const char* mesh = getenv("CKC_MESH_PATH"); const char* dev = getenv("CKC_DEVICE");
    const double* start = env == 2 ? kStart2 : kStart1; const double* goal = env == 2 ? kGoal2 : kGoal1;
    std::vector<Ensemble> E(workers);
    for (int w = 0; w < workers; w++) {
        E[w].planner = pl; E[w].max_step = max_step; E[w].start = start; E[w].goal = goal; E[w].proj.init();
        if (ckc_create(&E[w].ctx, env, mesh ? mesh : "./simulator", dev ? atoi(dev) : 0) != CKC_OK) { fprintf(stderr, "ckc_create failed\n"); return 1; }
        double o[12]; uint8_t v; int32_t nc;                          /* untimed warm-up (buffer allocation) */
        ckc_gd_validate(E[w].ctx, 1, start, o, &v); ckc_gd_check_motion_rbs(E[w].ctx, 1, start, goal, &v, &nc);
    }
    std::vector<std::vector<int> > qid(workers);
    for (int k = 0; k < nq; k++) {
        const int w = k % workers;
        E[w].qs.emplace_back(); qid[w].push_back(k);
        Query& Q = E[w].qs.back();
        Q.rng.seed(seed * 1000003ull + (uint64_t)k);
        E[w].add_motion(Q, 0, start, -1, true);
        if (pl == P_SBL) E[w].add_motion(Q, 1, goal, -1, true);
    }
    std::vector<double> lat;
    const auto t0 = std::chrono::steady_clock::now();
    if (sequential) {
        std::vector<Query> all_q; all_q.swap(E[0].qs);
        std::vector<Query> done;
        for (auto& Q : all_q) {
            E[0].qs.clear(); E[0].qs.push_back(std::move(Q));
            const auto q0 = std::chrono::steady_clock::now();
            E[0].run(max_rounds);
            lat.push_back(1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - q0).count());
            done.push_back(std::move(E[0].qs[0]));
        }
        E[0].qs.swap(done);
    } else if (workers == 1) E[0].run(max_rounds);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < workers; w++) th.emplace_back([&E, w, max_rounds]() { E[w].run(max_rounds); });
        for (auto& t : th) t.join();
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<const Query*> all(nq);
    for (int w = 0; w < workers; w++) for (size_t i = 0; i < qid[w].size(); i++) all[qid[w][i]] = &E[w].qs[i];
    FILE* fp = fopen(argv[6], "w");
    if (!fp) { fprintf(stderr, "cannot write %s\n", argv[6]); return 1; }
    int solved = 0; long rounds = 0, path_nodes = 0, tree_nodes = 0, calls = 0, items = 0;
    for (int k = 0; k < nq; k++) {
        const Query& Q = *all[k];
        tree_nodes += Q.t[0].size + Q.t[1].size;
        if (!Q.solved) continue;
        solved++; rounds += Q.rounds;
        const size_t m = Q.path.size() / 12; path_nodes += (long)m;
        fprintf(fp, "query %d nodes %zu rounds %d\n", k, m, Q.rounds);
        for (size_t i = 0; i < m; i++) {
            for (int j = 0; j < 12; j++) fprintf(fp, "%.17g ", Q.path[12 * i + j]);
            int e_dir = -1;
            if (i + 1 < m) {
                if (Q.path_dir[i] == 0 && Q.path_dir[i + 1] == 0) e_dir = 0;             /* parent -> child in the start tree */
                else if (Q.path_dir[i] == 1 && Q.path_dir[i + 1] == 1) e_dir = 1;        /* child <- parent in the goal tree */
                else e_dir = Q.junction_from_start ? 0 : 1;                            /* the junction: checked from the expanding tree's side */
            }
            fprintf(fp, "0 -1 -1 -1 -1 %d\n", e_dir);
        }
    }
    for (int w = 0; w < workers; w++) { calls += E[w].calls; items += E[w].items; }
    double lmean = 0, lmed = 0, lmax = 0;
    if (!lat.empty()) { std::vector<double> sl(lat); std::sort(sl.begin(), sl.end()); for (double v : sl) lmean += v; lmean /= sl.size(); lmed = sl[sl.size() / 2]; lmax = sl.back(); }
    fprintf(fp, "summary queries %d solved %d seconds %.6f ms_per_query %.4f workers %d gpu_calls %ld items %ld mean_rounds %.1f mean_path_nodes %.1f mean_tree_nodes %.1f max_step %.3f env %d flavour gd planner %s sequential %d latency_mean_ms %.4f latency_median_ms %.4f latency_max_ms %.4f\n",
            nq, solved, secs, 1e3 * secs / std::max(nq, 1), sequential ? 0 : workers, calls, items, solved ? (double)rounds / solved : 0.0, solved ? (double)path_nodes / solved : 0.0,
            (double)tree_nodes / std::max(nq, 1), max_step, env, argv[1], sequential ? 1 : 0, lmean, lmed, lmax);
    fclose(fp);
    printf("%s: queries %d solved %d in %.3f s (%.3f ms per query), %ld batched calls, %.1f items per call\n", argv[1], nq, solved, secs, 1e3 * secs / std::max(nq, 1), calls,
           calls ? (double)items / calls : 0.0);
    for (int w = 0; w < workers; w++) ckc_destroy(E[w].ctx);
    return 0;
}
