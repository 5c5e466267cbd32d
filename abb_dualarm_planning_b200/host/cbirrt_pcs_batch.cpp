/*
 * cbirrt_pcs_batch.cpp -- the caller side of the hot path (BASELINE configs[3]): an ensemble of
 * independent CBiRRT-PCS planning queries stepped in lock-step, so that every validity question the planners ask in one
 * step becomes ONE batched call of include/ckc.h.
 *
 * The planner is the reference's planners/CBiRRT_PCS.cpp (a constrained RRT-Connect) restated without OMPL: per query
 * two trees, growTree(mode 1 = extend toward a uniform sample, 100 steps; mode 2 = connect the other tree to the new node,
 * 500 steps) (:160-305), the solve loop that alternates the trees (:371-486), the projection of every step onto the
 * closure constraint with the neighbour's IK branch and fall-back to the other active chain (:191-246), identify_state_ik
 * on the projected state (:251), the RBS local connection with either active chain (:275-279), the duplicate-free path
 * assembly (:424-460).  Reproduced quirk: activeDistance with active chain 1 measures |q2(target)| (:127-139).
 * What differs from the reference is only WHEN the questions are asked: the reference asks them one at a time, this
 * driver collects the pending question of every live query and asks them together:
 *     step A  ckc_pcs_validate            project with (active chain, neighbour's branch) + joint limits + collision
 *     step A' ckc_pcs_validate            the fall-back projection with the other chain for the samples A could not close
 *     step B  ckc_identify_ik             both IK branches of the projected states
 *     step C  ckc_pcs_check_motion_rbs    local connection with active chain 0 where the branches agree
 *     step C' ckc_pcs_check_motion_rbs    ... with active chain 1 for the edges C rejected / could not try
 * One query (n = 1) therefore behaves like the reference's serial planner; n = 500 runs the reference's benchmark
 * protocol (500 repetitions of the same query, run/plan_PCS.cpp:300-330) in the time of a few.
 *
 * usage: cbirrt_pcs_batch <env> <n_queries> <max_step> <seed> <out.txt> [workers] [pcs|gd|hb] [max_rounds]
 * flavour gd = planners/CBiRRT_GD.cpp: 12-D distance, IKproject(dstate) = kdl::GD + limits + collision (ckc_gd_validate),
 * checkMotionRBS(s1, s2) (ckc_gd_check_motion_rbs); rows then carry -1 for the branch / edge columns.
 * out.txt: per solved query "query k nodes m rounds r" followed by m rows
 *          q[12] a_chain ik0 ik1 e_ac e_ik e_dir     (edge to the NEXT row: RBS arguments that validated it; e_dir 1 = checked
 *          from the next row toward this one; -1 -1 -1 on the last row), then one summary line starting with "summary".
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
#include "../../include/ckc.h"

static const double kPI_ = 3.1416;
/* run/plan_PCS.cpp:91-114: bounds of the sampled state space */
static const double kLo[12] = { -2.88, -1.919, -1.919, -2.79, -2.09, -kPI_, -2.88, -1.919, -1.919, -2.79, -2.09, -kPI_ };
static const double kHi[12] = { 2.88, 1.919, 1.22, 2.79, 2.09, kPI_, 2.88, 1.919, 1.22, 2.79, 2.09, kPI_ };
/* run/plan_PCS.cpp:268-276: the published start / goal pairs */
static const double kStart1[12] = { 0.5236, 1.7453, -1.8326, -1.4835, 1.5708, 0, 1.004278, 0.2729, 0.9486, -1.15011, 1.81001, -1.97739 };
static const double kGoal1[12] = { 0.5236, 0.34907, 0.69813, -1.3963, 1.5708, 0, 0.7096, 1.8032, -1.7061, -1.6286, 1.9143, -2.0155 };
static const double kStart2[12] = { 1.1, 1.1, 0, 1.24, -1.5708, 0, -0.79567, 0.60136, 0.43858, -0.74986, -1.0074, -0.092294 };
static const double kGoal2[12] = { -1.1, 1.35, -0.2, -1, -1.9, 0, 0.80875, 0.72363, -0.47891, -1.0484, 0.73278, 1.7491 };

struct Node {
    double q[12];
    int ik[2];
    int parent;
    int a_chain;
    int e_ac, e_ik;          /* RBS arguments that validated the edge parent -> this node */
};
struct Tree {
    std::vector<Node> n;
    int nearest(const double* q) const {            /* si_->distance: 12-D Euclid, linear scan */
        int best = 0; double bd = 1e300;
        for (size_t i = 0; i < n.size(); i++) {
            double s = 0;
            for (int j = 0; j < 12; j++) { const double d = n[i].q[j] - q[j]; s += d * d; }
            if (s < bd) { bd = s; best = (int)i; }
        }
        return best;
    }
};

enum Wait { W_NONE, W_PROJECT, W_PROJECT_FALLBACK, W_IDENTIFY, W_RBS0, W_RBS1, W_SPEC };

/* one step of a growTree window (CBiRRT-PCS): the interpolated state, after step A the projected one */
struct SpecStep {
    double ds[12];
    int ik[2];
    bool reach = false, direct = false, ok = false, valid = false;
    int rbs = -1;            /* -1: no RBS pass validated the edge yet; 0 / 1: the active chain that did */
};

struct Query {
    Tree t[2];                       /* 0 = start tree, 1 = goal tree */
    std::mt19937_64 rng;
    bool start_turn = true, solved = false, failed = false;
    int rounds = 0;
    /* solve-loop state */
    int stage = 0;                   /* 0: begin an iteration, 1: growTree(extend) running, 2: growTree(connect) running */
    int tree = 0, other = 1, added = -1;
    /* growTree activation record */
    int g_tree = 0, mode = 1, count = 0, nm = 0, active_chain = 0, xmotion = -1;
    double target[12]; int target_ik[2];
    bool reach = false, reached = false;
    /* the step in flight */
    Wait wait = W_NONE;
    double dstate[12]; int ik[2];
    std::vector<SpecStep> win; bool win_fallback = false, win_cut = false;      /* the window in flight (CBiRRT-PCS) */
    std::vector<int> path_tree, path_node;
    Node junction; bool junction_in_start_tree = false;      /* the copy of the connection state that :430-433 drops */
};

static double uni(std::mt19937_64& g) { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }

/* CBiRRT_PCS.cpp:122-140 including its active-chain-1 quirk (qa stays zero, qb = q2 of the target) */
static double active_distance(int active_chain, const double* a, const double* b) {
    double s = 0;
    if (!active_chain) for (int i = 0; i < 6; i++) s += (a[i] - b[i]) * (a[i] - b[i]);
    else for (int i = 0; i < 6; i++) s += b[6 + i] * b[6 + i];
    return std::sqrt(s);
}

struct Ensemble {
    ckc_ctx* ctx = nullptr;
    double max_step = 1.0;
    bool gd = false;                 /* CBiRRT-GD (planners/CBiRRT_GD.cpp) instead of CBiRRT-PCS */
    bool hb = false;                 /* CBiRRT-HB (planners/CBiRRT_HB.cpp): kdl::GD projection, IK-branch identification, APC (PCS) local connection */
    int window = 64;                 /* growTree steps asked per round and query (CBiRRT-PCS); 1 = the reference's step-by-step order */
    std::vector<Query> qs;
    long calls = 0, items = 0;
    double t_kind[3] = { 0, 0, 0 }; long n_kind[3] = { 0, 0, 0 }; long i_kind[3] = { 0, 0, 0 };      /* validate / identify / rbs */
    std::chrono::steady_clock::time_point t_mark;
    void mark() { t_mark = std::chrono::steady_clock::now(); }

    void begin_grow(Query& Q, int tree, int nm, const double* target, const int* tik, int mode, int count) {
        Q.g_tree = tree; Q.nm = nm; Q.mode = mode; Q.count = gd ? count + 1 : count; Q.reached = false;      /* GD: while (count >= 0), CBiRRT_GD.cpp:166 */
        memcpy(Q.target, target, sizeof(Q.target)); Q.target_ik[0] = tik[0]; Q.target_ik[1] = tik[1];
        Q.active_chain = (int)(Q.rng() & 1);                                   /* :168 */
    }

    /* growTree returned (:206,:221,:238,:265,:298-303): advance the solve loop; leaves wait == W_NONE */
    void end_grow(Query& Q) {
        Q.wait = W_NONE;
        if (Q.stage == 1) {
            Q.added = Q.nm;                                                    /* :411 reached_motion of the extend step */
            Q.xmotion = -1;
            const Node tn = Q.t[Q.tree].n[Q.added];
            const int nm2 = Q.t[Q.other].nearest(tn.q);                        /* :416 */
            Q.stage = 2;
            begin_grow(Q, Q.other, nm2, tn.q, tn.ik, 2, 500);                  /* :417 */
        } else {
            if (Q.reached) finish(Q);
            Q.stage = 0;
        }
    }

    /* :424-460: one step back on one side (the connection state exists in both trees), then the two root paths */
    void finish(Query& Q) {
        int sm = Q.other == 0 ? Q.xmotion : Q.added, gm = Q.other == 0 ? Q.added : Q.xmotion;
        /* the dropped copy's own edge record (parent -> copy) is the check that validated the junction edge of the path */
        if (Q.t[0].n[sm].parent >= 0) { Q.junction = Q.t[0].n[sm]; Q.junction_in_start_tree = true; sm = Q.t[0].n[sm].parent; }
        else { Q.junction = Q.t[1].n[gm]; Q.junction_in_start_tree = false; gm = Q.t[1].n[gm].parent; }
        std::vector<int> p1, p2;
        for (int k = sm; k >= 0; k = Q.t[0].n[k].parent) p1.push_back(k);
        for (int k = gm; k >= 0; k = Q.t[1].n[k].parent) p2.push_back(k);
        Q.path_tree.clear(); Q.path_node.clear();
        for (int i = (int)p1.size() - 1; i >= 0; i--) { Q.path_tree.push_back(0); Q.path_node.push_back(p1[i]); }
        for (size_t i = 0; i < p2.size(); i++) { Q.path_tree.push_back(1); Q.path_node.push_back(p2[i]); }
        Q.solved = true;
    }

    /* one growTree loop head (:173-190, :253-262): either posts the step's first question or ends the growTree call */
    void grow_head(Query& Q) {
        if (Q.count == 0) { end_grow(Q); return; }
        Q.count--;
        const Node& nmn = Q.t[Q.g_tree].n[Q.nm];
        double d;
        if (gd) {                                   /* CBiRRT_GD.cpp:171: distanceFunction = 12-D Euclid */
            d = 0;
            for (int j = 0; j < 12; j++) d += (nmn.q[j] - Q.target[j]) * (nmn.q[j] - Q.target[j]);
            d = std::sqrt(d);
        } else d = active_distance(Q.active_chain, nmn.q, Q.target);
        if (d > max_step) {
            const double f = max_step / d;
            for (int j = 0; j < 12; j++) Q.dstate[j] = nmn.q[j] + (Q.target[j] - nmn.q[j]) * f;      /* RealVectorStateSpace::interpolate */
            Q.reach = false;
        } else { memcpy(Q.dstate, Q.target, sizeof(Q.dstate)); Q.reach = true; }
        if (Q.mode == 1 || !Q.reach) { Q.ik[0] = Q.ik[1] = -1; Q.wait = W_PROJECT; return; }
        if (gd) { Q.wait = W_RBS0; return; }       /* CBiRRT_GD.cpp:183-206: no branch bookkeeping */
        /* the state belongs to the opposite tree and already satisfies the constraint (CBiRRT_HB.cpp:205-212) */
        Q.ik[0] = Q.target_ik[0]; Q.ik[1] = Q.target_ik[1];
        if (nmn.ik[0] != Q.ik[0] && nmn.ik[1] != Q.ik[1]) { end_grow(Q); return; }
        Q.wait = W_RBS0;
    }

    /* host logic of one query until it has a question for the device (or is solved) */
    void advance(Query& Q) {
        while (!Q.solved && Q.wait == W_NONE) {
            if (Q.stage == 0) begin_iteration(Q);
            if (gd || hb) grow_head(Q);
            else if (Q.count == 0) end_grow(Q);
            else build_window(Q);
        }
    }

    /* start an iteration of the solve loop (:371-413) */
    void begin_iteration(Query& Q) {
        Q.tree = Q.start_turn ? 0 : 1; Q.start_turn = !Q.start_turn; Q.other = 1 - Q.tree;
        double r[12];
        for (int j = 0; j < 12; j++) r[j] = kLo[j] + uni(Q.rng) * (kHi[j] - kLo[j]);       /* sampler_->sampleUniform */
        const int tik[2] = { -1, -1 };
        Q.stage = 1;
        begin_grow(Q, Q.tree, Q.t[Q.tree].nearest(r), r, tik, 1, 100);                        /* :409-410 */
    }

    void run(int max_rounds) {
        std::vector<int> idx; std::vector<double> q, qo, qa, qb; std::vector<int8_t> ac, ik, id; std::vector<uint8_t> ok, valid, res; std::vector<int32_t> nchk;
        for (int round = 0; round < max_rounds; round++) {
            /* host part: every live query advances to its next question */
            int live = 0;
            for (auto& Q : qs) {
                if (Q.solved || Q.failed) continue;
                live++;
                advance(Q);
                if (Q.solved) continue;
                Q.rounds++;
            }
            if (!live) break;
            if (gd) { run_gd_steps(idx, q, qo, qa, qb, valid, res, nchk); continue; }
            if (hb) { run_hb_steps(idx, q, qo, qa, qb, ac, ik, id, valid, res, nchk); continue; }
            run_pcs_steps(idx, q, qo, qa, qb, ac, ik, id, ok, valid, res, nchk);
        }
    }

    /* ---- CBiRRT-PCS: one round = the WINDOW of every live query ------------------------------------------------------------
     * Inside one growTree call the reference walks from nmotion toward the target in steps of max_step, and each step is
     * project -> identify -> RBS, one question at a time (:173-298).  The questions of consecutive steps look dependent (step k + 1
     * starts from the node step k added) but are not: the projection only REPLACES the passive arm (apc_class.cpp:356-389), so the
     * active arm's joints of node k are those of the interpolated state, the next interpolation factor (activeDistance, :122-140)
     * and the next active joints follow from them alone, and the projection arguments (active chain, that chain's IK branch) are
     * inherited unchanged.  The whole run of steps is therefore known up front: a window of up to `window` steps is projected +
     * collision-checked in ONE ckc_pcs_validate, identified in one ckc_identify_ik and its edges are checked in one or two
     * ckc_pcs_check_motion_rbs calls; the steps are then committed in order up to the first one the reference would have stopped
     * at (projection failure -> that step is re-asked exactly, with the fall-back chain; collision or failed edge -> growTree
     * returns).  Same tree, same path, same random stream as the step-by-step planner (window = 1 IS that planner); only
     * speculative work behind a failing step is wasted.  This is the batching INSIDE one query that SURVEY 8(f) row 1 asks for. */
    void build_window(Query& Q) {
        const Node& nmn = Q.t[Q.g_tree].n[Q.nm];
        Q.win.clear(); Q.win_fallback = false; Q.win_cut = false;
        double cur[12]; memcpy(cur, nmn.q, 96);
        int cnt = Q.count;
        while (cnt > 0 && (int)Q.win.size() < window) {
            cnt--;
            SpecStep st;
            const double d = active_distance(Q.active_chain, cur, Q.target);
            if (d > max_step) {
                const double f = max_step / d;
                for (int j = 0; j < 12; j++) st.ds[j] = cur[j] + (Q.target[j] - cur[j]) * f;      /* RealVectorStateSpace::interpolate */
                st.reach = false;
            } else { memcpy(st.ds, Q.target, 96); st.reach = true; }
            st.direct = Q.mode == 2 && st.reach;       /* the state belongs to the opposite tree and already satisfies the constraint */
            st.ik[0] = st.ik[1] = -1;
            if (st.direct) { st.ik[0] = Q.target_ik[0]; st.ik[1] = Q.target_ik[1]; st.ok = st.valid = true; }
            Q.win.push_back(st);
            if (st.reach) break;
            memcpy(cur, st.ds, 96);                    /* active joints exact; the passive ones are replaced by the projection anyway */
        }
        Q.wait = W_SPEC;
    }

    void run_pcs_steps(std::vector<int>& idx, std::vector<double>& q, std::vector<double>& qo, std::vector<double>& qa, std::vector<double>& qb,
                       std::vector<int8_t>& ac, std::vector<int8_t>& ik, std::vector<int8_t>& id, std::vector<uint8_t>& ok, std::vector<uint8_t>& valid,
                       std::vector<uint8_t>& res, std::vector<int32_t>& nchk) {
        std::vector<std::pair<int, int> > ref;          /* (query, step) of every batched item */
        /* ---- step A / A': projection + collision of the windows; A' = the fall-back chain for a window whose FIRST step did not close ---- */
        for (int pass = 0; pass < 2; pass++) {
            ref.clear();
            for (size_t k = 0; k < qs.size(); k++) {
                Query& Q = qs[k];
                if (Q.solved || Q.wait != W_SPEC || Q.win_fallback != (pass == 1)) continue;
                for (size_t i = 0; i < Q.win.size(); i++) if (!Q.win[i].direct) ref.push_back(std::make_pair((int)k, (int)i));
            }
            if (ref.empty()) continue;
            const size_t m = ref.size();
            q.resize(m * 12); qo.resize(m * 12); ac.resize(m); ik.resize(m); ok.resize(m); valid.resize(m);
            for (size_t i = 0; i < m; i++) {
                Query& Q = qs[ref[i].first];
                memcpy(&q[12 * i], Q.win[ref[i].second].ds, 96);
                ac[i] = (int8_t)Q.active_chain; ik[i] = (int8_t)Q.t[Q.g_tree].n[Q.nm].ik[Q.active_chain];
            }
            mark();
            check(ckc_pcs_validate(ctx, (int64_t)m, q.data(), ac.data(), ik.data(), qo.data(), ok.data(), valid.data()), m, 0);
            for (size_t i = 0; i < m; i++) {
                Query& Q = qs[ref[i].first]; SpecStep& st = Q.win[ref[i].second];
                st.ok = ok[i] != 0; st.valid = valid[i] != 0;
                if (st.ok) { memcpy(st.ds, &qo[12 * i], 96); st.ik[Q.active_chain] = Q.t[Q.g_tree].n[Q.nm].ik[Q.active_chain]; }
            }
            for (size_t k = 0; k < qs.size(); k++) {
                Query& Q = qs[k];
                if (Q.solved || Q.wait != W_SPEC || Q.win_fallback != (pass == 1)) continue;
                for (size_t i = 0; i < Q.win.size(); i++) {
                    const SpecStep& st = Q.win[i];
                    if (st.direct) continue;
                    if (!st.ok) {
                        if (i > 0) Q.win.resize(i);                       /* commit the steps before it; this one is re-asked exactly next round */
                        else if (pass == 0) { Q.active_chain = 1 - Q.active_chain; Q.win.resize(1); Q.win_fallback = true; }      /* :196-203 / :214-221 */
                        else { Q.count--; end_grow(Q); }                  /* neither chain closes the step */
                        break;
                    }
                    if (!st.valid) {                                      /* :236-240 collision: growTree returns after the steps before it */
                        Q.win.resize(i); Q.win_cut = true;
                        if (i == 0) { Q.count--; end_grow(Q); }
                        break;
                    }
                }
            }
        }
        /* ---- step B: identify_state_ik of every projected state (:251) ---- */
        ref.clear();
        for (size_t k = 0; k < qs.size(); k++) {
            Query& Q = qs[k];
            if (Q.solved || Q.wait != W_SPEC) continue;
            for (size_t i = 0; i < Q.win.size(); i++) if (!Q.win[i].direct) ref.push_back(std::make_pair((int)k, (int)i));
        }
        if (!ref.empty()) {
            const size_t m = ref.size();
            q.resize(m * 12); id.resize(m * 2);
            for (size_t i = 0; i < m; i++) memcpy(&q[12 * i], qs[ref[i].first].win[ref[i].second].ds, 96);
            mark();
            check(ckc_identify_ik(ctx, (int64_t)m, q.data(), id.data()), m, 1);
            for (size_t i = 0; i < m; i++) {
                SpecStep& st = qs[ref[i].first].win[ref[i].second];
                for (int c = 0; c < 2; c++) if (st.ik[c] == -1) st.ik[c] = id[2 * i + c];
            }
        }
        /* ---- step C / C': RBS local connection of every window edge, active chain 0 where the branches agree, else / then chain 1 (:275-279) ---- */
        for (int pass = 0; pass < 2; pass++) {
            ref.clear();
            for (size_t k = 0; k < qs.size(); k++) {
                Query& Q = qs[k];
                if (Q.solved || Q.wait != W_SPEC) continue;
                for (size_t i = 0; i < Q.win.size(); i++) {
                    const SpecStep& st = Q.win[i];
                    const int* pik = i == 0 ? Q.t[Q.g_tree].n[Q.nm].ik : Q.win[i - 1].ik;
                    if (st.rbs < 0 && pik[pass] == st.ik[pass]) ref.push_back(std::make_pair((int)k, (int)i));
                }
            }
            if (ref.empty()) continue;
            const size_t m = ref.size();
            qa.resize(m * 12); qb.resize(m * 12); ac.resize(m); ik.resize(m); res.resize(m); nchk.resize(m);
            for (size_t i = 0; i < m; i++) {
                Query& Q = qs[ref[i].first]; const int w = ref[i].second;
                const double* pq = w == 0 ? Q.t[Q.g_tree].n[Q.nm].q : Q.win[w - 1].ds;
                const int* pik = w == 0 ? Q.t[Q.g_tree].n[Q.nm].ik : Q.win[w - 1].ik;
                memcpy(&qa[12 * i], pq, 96); memcpy(&qb[12 * i], Q.win[w].ds, 96);
                ac[i] = (int8_t)pass; ik[i] = (int8_t)pik[pass];
            }
            mark();
            check(ckc_pcs_check_motion_rbs(ctx, (int64_t)m, qa.data(), qb.data(), ac.data(), ik.data(), res.data(), nchk.data()), m, 2);
            for (size_t i = 0; i < m; i++) if (res[i]) qs[ref[i].first].win[ref[i].second].rbs = pass;
        }
        /* ---- commit the windows in step order (:283-303) ---- */
        for (auto& Q : qs) {
            if (Q.solved || Q.wait != W_SPEC) continue;
            Q.wait = W_NONE;
            bool ended = false;
            for (size_t i = 0; i < Q.win.size() && !ended; i++) {
                const SpecStep& st = Q.win[i];
                Q.count--;
                if (st.rbs < 0) { Q.xmotion = Q.nm; end_grow(Q); ended = true; break; }
                const Node& nmn = Q.t[Q.g_tree].n[Q.nm];
                Node nn;
                memcpy(nn.q, st.ds, 96); nn.ik[0] = st.ik[0]; nn.ik[1] = st.ik[1];
                nn.parent = Q.nm; nn.e_ac = st.rbs; nn.e_ik = nmn.ik[st.rbs];
                nn.a_chain = st.direct ? (nmn.ik[0] == st.ik[0] ? 0 : 1) : Q.active_chain;          /* :262-266 picks the chain of the direct step */
                Q.t[Q.g_tree].n.push_back(nn);
                Q.nm = (int)Q.t[Q.g_tree].n.size() - 1; Q.xmotion = Q.nm;
                if (st.reach) { Q.reached = true; end_grow(Q); ended = true; }                      /* :295-298 */
            }
            if (!ended && Q.win_cut) { Q.count--; end_grow(Q); }
        }
    }
    /* CBiRRT-GD: IKproject(dstate) = kdl::GD + joint limits + collision (StateValidityCheckerGD.cpp:96-110) in one batched
     * ckc_gd_validate, then checkMotionRBS(nmotion, dstate) (:214-241) in one batched ckc_gd_check_motion_rbs */
    void run_gd_steps(std::vector<int>& idx, std::vector<double>& q, std::vector<double>& qo, std::vector<double>& qa, std::vector<double>& qb,
                      std::vector<uint8_t>& valid, std::vector<uint8_t>& res, std::vector<int32_t>& nchk) {
        idx.clear();
        for (size_t k = 0; k < qs.size(); k++) if (!qs[k].solved && qs[k].wait == W_PROJECT) idx.push_back((int)k);
        if (!idx.empty()) {
            const size_t m = idx.size();
            q.resize(m * 12); qo.resize(m * 12); valid.resize(m);
            for (size_t i = 0; i < m; i++) memcpy(&q[12 * i], qs[idx[i]].dstate, 96);
            mark();
            check(ckc_gd_validate(ctx, (int64_t)m, q.data(), qo.data(), valid.data()), m, 0);
            for (size_t i = 0; i < m; i++) {
                Query& Q = qs[idx[i]];
                if (!valid[i]) { end_grow(Q); continue; }                                     /* CBiRRT_GD.cpp:187-191 */
                memcpy(Q.dstate, &qo[12 * i], 96);
                Q.wait = W_RBS0;
            }
        }
        idx.clear();
        for (size_t k = 0; k < qs.size(); k++) if (!qs[k].solved && qs[k].wait == W_RBS0) idx.push_back((int)k);
        if (idx.empty()) return;
        const size_t m = idx.size();
        qa.resize(m * 12); qb.resize(m * 12); res.resize(m); nchk.resize(m);
        for (size_t i = 0; i < m; i++) {
            Query& Q = qs[idx[i]];
            memcpy(&qa[12 * i], Q.t[Q.g_tree].n[Q.nm].q, 96); memcpy(&qb[12 * i], Q.dstate, 96);
        }
        mark();
        check(ckc_gd_check_motion_rbs(ctx, (int64_t)m, qa.data(), qb.data(), res.data(), nchk.data()), m, 2);
        for (size_t i = 0; i < m; i++) {
            Query& Q = qs[idx[i]];
            if (!res[i]) { Q.xmotion = Q.nm; end_grow(Q); continue; }                         /* :229-232 */
            Node nn;
            memcpy(nn.q, Q.dstate, 96); nn.ik[0] = nn.ik[1] = -1; nn.parent = Q.nm; nn.a_chain = 0; nn.e_ac = nn.e_ik = -1;
            Q.t[Q.g_tree].n.push_back(nn);
            Q.nm = (int)Q.t[Q.g_tree].n.size() - 1; Q.xmotion = Q.nm;
            Q.wait = W_NONE;
            if (Q.reach) { Q.reached = true; end_grow(Q); }
        }
    }

    /* CBiRRT-HB (planners/CBiRRT_HB.cpp:160-246): the step is projected with kdl::GD + limits + collision (IKproject(dstate),
     * StateValidityCheckerHB.cpp:109-122 -> ckc_gd_validate), BOTH IK branches of the projected state are identified from scratch
     * (:199 -> ckc_identify_ik), growTree returns when neither branch matches the neighbour's (:211), and the edge is checked with the
     * APC bisection, active chain 0 where that branch agrees, then chain 1 (:217-221 -> ckc_pcs_check_motion_rbs).  The Newton
     * projection moves all twelve joints, so the steps of a growTree call cannot be asked as one window: one step per round. */
    void run_hb_steps(std::vector<int>& idx, std::vector<double>& q, std::vector<double>& qo, std::vector<double>& qa, std::vector<double>& qb,
                      std::vector<int8_t>& ac, std::vector<int8_t>& ik, std::vector<int8_t>& id, std::vector<uint8_t>& valid, std::vector<uint8_t>& res,
                      std::vector<int32_t>& nchk) {
        idx.clear();
        for (size_t k = 0; k < qs.size(); k++) if (!qs[k].solved && qs[k].wait == W_PROJECT) idx.push_back((int)k);
        if (!idx.empty()) {
            const size_t m = idx.size();
            q.resize(m * 12); qo.resize(m * 12); valid.resize(m); id.resize(m * 2);
            for (size_t i = 0; i < m; i++) memcpy(&q[12 * i], qs[idx[i]].dstate, 96);
            mark();
            check(ckc_gd_validate(ctx, (int64_t)m, q.data(), qo.data(), valid.data()), m, 0);
            mark();
            check(ckc_identify_ik(ctx, (int64_t)m, qo.data(), id.data()), m, 1);      /* rows of rejected samples are ignored */
            for (size_t i = 0; i < m; i++) {
                Query& Q = qs[idx[i]];
                if (!valid[i]) { end_grow(Q); continue; }                                     /* :192-193 */
                memcpy(Q.dstate, &qo[12 * i], 96);
                Q.ik[0] = id[2 * i]; Q.ik[1] = id[2 * i + 1];                                 /* :199 */
                const Node& nmn = Q.t[Q.g_tree].n[Q.nm];
                if (nmn.ik[0] != Q.ik[0] && nmn.ik[1] != Q.ik[1]) { end_grow(Q); continue; } /* :211-212 */
                Q.wait = W_RBS0;
            }
        }
        for (int pass = 0; pass < 2; pass++) {
            idx.clear();
            for (size_t k = 0; k < qs.size(); k++) {
                Query& Q = qs[k];
                if (Q.solved || Q.wait != (pass == 0 ? W_RBS0 : W_RBS1)) continue;
                const Node& nmn = Q.t[Q.g_tree].n[Q.nm];
                if (nmn.ik[pass] == Q.ik[pass]) idx.push_back((int)k);
                else if (pass == 0) Q.wait = W_RBS1;
                else { Q.xmotion = Q.nm; end_grow(Q); }
            }
            if (idx.empty()) continue;
            const size_t m = idx.size();
            qa.resize(m * 12); qb.resize(m * 12); ac.resize(m); ik.resize(m); res.resize(m); nchk.resize(m);
            for (size_t i = 0; i < m; i++) {
                Query& Q = qs[idx[i]];
                const Node& nmn = Q.t[Q.g_tree].n[Q.nm];
                memcpy(&qa[12 * i], nmn.q, 96); memcpy(&qb[12 * i], Q.dstate, 96);
                ac[i] = (int8_t)pass; ik[i] = (int8_t)nmn.ik[pass];
            }
            mark();
            check(ckc_pcs_check_motion_rbs(ctx, (int64_t)m, qa.data(), qb.data(), ac.data(), ik.data(), res.data(), nchk.data()), m, 2);
            for (size_t i = 0; i < m; i++) {
                Query& Q = qs[idx[i]];
                if (!res[i]) {
                    if (pass == 0) Q.wait = W_RBS1;
                    else { Q.xmotion = Q.nm; end_grow(Q); }                                   /* :243-246 */
                    continue;
                }
                Node nn;                                                                      /* :225-236 */
                memcpy(nn.q, Q.dstate, 96); nn.ik[0] = Q.ik[0]; nn.ik[1] = Q.ik[1];
                nn.parent = Q.nm; nn.a_chain = pass; nn.e_ac = pass; nn.e_ik = Q.t[Q.g_tree].n[Q.nm].ik[pass];
                Q.t[Q.g_tree].n.push_back(nn);
                Q.nm = (int)Q.t[Q.g_tree].n.size() - 1; Q.xmotion = Q.nm;
                Q.wait = W_NONE;
                if (Q.reach) { Q.reached = true; end_grow(Q); }                               /* :238-241 */
            }
        }
    }

    void check(int rc, size_t m, int kind) {
        if (rc != CKC_OK) { fprintf(stderr, "ckc: %s\n", ckc_last_error(ctx)); exit(-1); }
        t_kind[kind] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_mark).count();
        n_kind[kind]++; i_kind[kind] += (long)m;
        calls++; items += (long)m;
    }
};

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s <env> <n_queries> <max_step> <seed> <out.txt> [workers] [pcs|gd|hb] [max_rounds]\n", argv[0]); return 2; }
    const int env = atoi(argv[1]), nq = atoi(argv[2]); const double max_step = atof(argv[3]); const uint64_t seed = strtoull(argv[4], nullptr, 10);
    /* workers = 0: SEQUENTIAL mode -- the queries run one after another, each alone on the device (single-query latency, what the
     * reference's 500-repetition benchmark measures); the summary then also carries latency_mean / median / max */
    const bool sequential = argc > 6 && atoi(argv[6]) == 0;
    const int workers = std::max(1, std::min(argc > 6 ? atoi(argv[6]) : 1, std::max(nq, 1)));
    const bool gd = argc > 7 && strcmp(argv[7], "gd") == 0;
    const bool hb = argc > 7 && strcmp(argv[7], "hb") == 0;
    const int max_rounds = argc > 8 ? atoi(argv[8]) : 20000;
    const char* mesh = getenv("CKC_MESH_PATH"); const char* dev = getenv("CKC_DEVICE");
    const double* start = env == 2 ? kStart2 : kStart1; const double* goal = env == 2 ? kGoal2 : kGoal1;
    /* one libckc context per worker thread (a context serves one host thread at a time, include/ckc.h); the workers'
     * batched calls overlap on the device, which hides the per-call latency of the thin RBS levels */
    std::vector<Ensemble> E(workers);
    int8_t id[4]; uint8_t hit[2];
    for (int w = 0; w < workers; w++) {
        E[w].max_step = max_step; E[w].gd = gd; E[w].hb = hb;
        if (getenv("CKC_PLAN_WINDOW")) E[w].window = std::max(1, atoi(getenv("CKC_PLAN_WINDOW")));
        if (ckc_create(&E[w].ctx, env, mesh ? mesh : "./simulator", dev ? atoi(dev) : 0) != CKC_OK) { fprintf(stderr, "ckc_create failed\n"); return 1; }
    }
    /* :325-343, :380-396: start and goal must be identifiable on both chains and collision free */
    double sg[24]; memcpy(sg, start, 96); memcpy(sg + 12, goal, 96);
    if (ckc_identify_ik(E[0].ctx, 2, sg, id) != CKC_OK || ckc_collide(E[0].ctx, 2, sg, hit) != CKC_OK) { fprintf(stderr, "ckc: %s\n", ckc_last_error(E[0].ctx)); return 1; }
    if (!gd && (id[0] < 0 || id[1] < 0 || id[2] < 0 || id[3] < 0 || hit[0] || hit[1])) { fprintf(stderr, "start / goal state not feasible\n"); return 1; }
    if (gd) id[0] = id[1] = id[2] = id[3] = -1;                          /* CBiRRT_GD.cpp:252-256: no feasibility test, no branches */
    std::vector<std::vector<int> > qid(workers);
    for (int k = 0; k < nq; k++) {
        const int w = k % workers;
        E[w].qs.emplace_back(); qid[w].push_back(k);
        Query& Q = E[w].qs.back();
        Q.rng.seed(seed * 1000003ull + (uint64_t)k);                    /* a query's random stream does not depend on the ensemble */
        for (int t = 0; t < 2; t++) {
            Node r; memcpy(r.q, t == 0 ? start : goal, 96); r.ik[0] = id[2 * t]; r.ik[1] = id[2 * t + 1]; r.parent = -1; r.a_chain = 0; r.e_ac = r.e_ik = -1;
            Q.t[t].n.push_back(r);
        }
    }
    /* one untimed warm-up call of each entry point (buffer allocation) */
    for (int w = 0; w < workers; w++) {
        double wq[12]; memcpy(wq, start, 96); int8_t a = 0, b = id[0]; double o[12]; uint8_t k1, k2; int32_t nc;
        if (gd) { ckc_gd_validate(E[w].ctx, 1, wq, o, &k1); ckc_gd_check_motion_rbs(E[w].ctx, 1, start, start, &k1, &nc); }
        else { b = b < 0 ? 0 : b; ckc_pcs_validate(E[w].ctx, 1, wq, &a, &b, o, &k1, &k2); ckc_pcs_check_motion_rbs(E[w].ctx, 1, start, start, &a, &b, &k1, &nc); }
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
    }
    else if (workers == 1) E[0].run(max_rounds);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < workers; w++) th.emplace_back([&E, w, max_rounds]() { E[w].run(max_rounds); });
        for (auto& t : th) t.join();
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::vector<const Query*> all(nq);
    for (int w = 0; w < workers; w++) for (size_t i = 0; i < qid[w].size(); i++) all[qid[w][i]] = &E[w].qs[i];
    FILE* fp = fopen(argv[5], "w");
    if (!fp) { fprintf(stderr, "cannot write %s\n", argv[5]); return 1; }
    int solved = 0; long nodes_trees = 0, rounds = 0, path_nodes = 0, calls = 0, items = 0;
    for (int k = 0; k < nq; k++) {
        const Query& Q = *all[k];
        nodes_trees += (long)(Q.t[0].n.size() + Q.t[1].n.size());
        if (!Q.solved) continue;
        solved++; rounds += Q.rounds; path_nodes += (long)Q.path_node.size();
        fprintf(fp, "query %d nodes %zu rounds %d\n", k, Q.path_node.size(), Q.rounds);
        for (size_t i = 0; i < Q.path_node.size(); i++) {
            const Node& n = Q.t[Q.path_tree[i]].n[Q.path_node[i]];
            for (int j = 0; j < 12; j++) fprintf(fp, "%.17g ", n.q[j]);
            int e_ac = -1, e_ik = -1, e_dir = -1;
            if (i + 1 < Q.path_node.size()) {
                const int tn = Q.path_tree[i + 1];
                const Node& nx = Q.t[tn].n[Q.path_node[i + 1]];
                if (Q.path_tree[i] == 0 && tn == 0) { e_ac = nx.e_ac; e_ik = nx.e_ik; e_dir = 0; }            /* parent -> child in the start tree */
                else if (Q.path_tree[i] == 1 && tn == 1) { e_ac = n.e_ac; e_ik = n.e_ik; e_dir = 1; }        /* child -> parent in the goal tree */
                else {                                                                                       /* the junction */
                    e_ac = Q.junction.e_ac; e_ik = Q.junction.e_ik;
                    e_dir = Q.junction_in_start_tree ? 0 : 1;
                }
            }
            fprintf(fp, "%d %d %d %d %d %d\n", n.a_chain, n.ik[0], n.ik[1], e_ac, e_ik, e_dir);
        }
    }
    static const char* kn[3] = { "validate", "identify", "rbs" };
    for (int k = 0; k < 3; k++) {
        long n = 0, it = 0; double t = 0;
        for (int w = 0; w < workers; w++) { n += E[w].n_kind[k]; it += E[w].i_kind[k]; t += E[w].t_kind[k]; }
        calls += n; items += it;
        printf("  %-8s %6ld calls %8ld items  %.3f s summed over workers (%.1f us per call)\n", kn[k], n, it, t, n ? 1e6 * t / n : 0.0);
    }
    double lmean = 0, lmed = 0, lmax = 0;
    if (!lat.empty()) {
        std::vector<double> sl(lat); std::sort(sl.begin(), sl.end());
        for (double v : sl) lmean += v;
        lmean /= sl.size(); lmed = sl[sl.size() / 2]; lmax = sl.back();
    }
    fprintf(fp, "summary queries %d solved %d seconds %.6f ms_per_query %.4f workers %d gpu_calls %ld items %ld mean_rounds %.1f mean_path_nodes %.1f mean_tree_nodes %.1f max_step %.3f env %d flavour %s sequential %d latency_mean_ms %.4f latency_median_ms %.4f latency_max_ms %.4f\n",
            nq, solved, secs, 1e3 * secs / std::max(nq, 1), sequential ? 0 : workers, calls, items, solved ? (double)rounds / solved : 0.0, solved ? (double)path_nodes / solved : 0.0,
            (double)nodes_trees / std::max(nq, 1), max_step, env, gd ? "gd" : (hb ? "hb" : "pcs"), sequential ? 1 : 0, lmean, lmed, lmax);
    fclose(fp);
    printf("queries %d solved %d in %.3f s (%.3f ms per query), %d workers, %ld batched calls, %.1f items per call\n", nq, solved, secs, 1e3 * secs / std::max(nq, 1),
           workers, calls, calls ? (double)items / calls : 0.0);
    for (int w = 0; w < workers; w++) ckc_destroy(E[w].ctx);
    return 0;
}
