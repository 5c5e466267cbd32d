/*
 * prm_pcs_batch.cpp -- the caller side of the hot path, roadmap flavour ("PRM addMilestone"): the
 * reference's planners/PRM_PCS.cpp (OMPL PRM with the PCS checker) restated without OMPL, with its two inner loops turned
 * into batches of the C ABI:
 *   growRoadmap  (:305-331)  sample_q until a valid state is found        -> ckc_pcs_validate on ~256 candidates per wanted milestone
 *                                                                             + ckc_identify_ik on the survivors
 *   addMilestone (:525-568)  for each of the k = 10 nearest milestones      -> ONE ckc_pcs_check_motion_rbs over the candidate
 *                            (KStrategy, :69,:128): identify_state_ik of       edges of all M new milestones with active chain 0,
 *                            both ends, checkMotionRBS with chain 0 if the      a second one with chain 1 for the edges the first
 *                            branches agree, else / then chain 1                rejected or could not try
 * Milestone i of a round sees the milestones before it (also those of its own round), as in the sequential original; the
 * start and goal are milestones 0 and 1; the query is solved when they share a component (union-find), the path is the
 * cheapest one by 12-D length (the reference's A* over motionCost).
 * Deviation, on purpose: sample_q of the reference accepts a sample unless it is colliding AND outside the joint limits
 * (StateValidityCheckerPCS.cpp:217-220), i.e. it admits colliding milestones; here a milestone must be IK feasible and
 * collision free (the drop-in class keeps the reference's rule).
 *
 * usage: prm_pcs_batch <env> <milestones_per_round> <seed> <out.txt> [max_rounds]
 * out.txt: "query 0 nodes m rounds r" + m rows  q[12] 0 ik0 ik1 e_ac e_ik e_dir  (as cbirrt_pcs_batch), then a summary line.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <random>
#include <vector>
#include "../../include/ckc.h"

static const double kPI_ = 3.1416;
static const double kStart1[12] = { 0.5236, 1.7453, -1.8326, -1.4835, 1.5708, 0, 1.004278, 0.2729, 0.9486, -1.15011, 1.81001, -1.97739 };   /* run/plan_PCS.cpp:268-276 */
static const double kGoal1[12] = { 0.5236, 0.34907, 0.69813, -1.3963, 1.5708, 0, 0.7096, 1.8032, -1.7061, -1.6286, 1.9143, -2.0155 };
static const double kStart2[12] = { 1.1, 1.1, 0, 1.24, -1.5708, 0, -0.79567, 0.60136, 0.43858, -0.74986, -1.0074, -0.092294 };
static const double kGoal2[12] = { -1.1, 1.35, -0.2, -1, -1.9, 0, 0.80875, 0.72363, -0.47891, -1.0484, 0.73278, 1.7491 };

struct Vertex { double q[12]; int ik[2]; };
struct Edge { int to; double w; int ac, ik, from_first; };      /* from_first: RBS was run from the smaller to the larger vertex index */

static double uni(std::mt19937_64& g) { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }
static double dist12(const double* a, const double* b) { double s = 0; for (int j = 0; j < 12; j++) s += (a[j] - b[j]) * (a[j] - b[j]); return std::sqrt(s); }
static int find(std::vector<int>& p, int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }

#define CK(call) do { if ((call) != CKC_OK) { fprintf(stderr, "ckc: %s\n", ckc_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s <env> <milestones_per_round> <seed> <out.txt> [max_rounds]\n", argv[0]); return 2; }
    const int env = atoi(argv[1]), M = std::max(1, atoi(argv[2])); const uint64_t seed = strtoull(argv[3], nullptr, 10);
    const int max_rounds = argc > 5 ? atoi(argv[5]) : 200, K = 10;
    const char* mesh = getenv("CKC_MESH_PATH"); const char* dev = getenv("CKC_DEVICE");
    ckc_ctx* ctx = nullptr;
    if (ckc_create(&ctx, env, mesh ? mesh : "./simulator", dev ? atoi(dev) : 0) != CKC_OK) { fprintf(stderr, "ckc_create failed\n"); return 1; }
    const double* start = env == 2 ? kStart2 : kStart1; const double* goal = env == 2 ? kGoal2 : kGoal1;
    std::vector<Vertex> V(2);
    memcpy(V[0].q, start, 96); memcpy(V[1].q, goal, 96);
    {
        double sg[24]; memcpy(sg, start, 96); memcpy(sg + 12, goal, 96);
        int8_t id[4]; uint8_t hit[2];
        CK(ckc_identify_ik(ctx, 2, sg, id)); CK(ckc_collide(ctx, 2, sg, hit));
        if (id[0] < 0 || id[1] < 0 || id[2] < 0 || id[3] < 0 || hit[0] || hit[1]) { fprintf(stderr, "start / goal state not feasible\n"); return 1; }
        V[0].ik[0] = id[0]; V[0].ik[1] = id[1]; V[1].ik[0] = id[2]; V[1].ik[1] = id[3];
    }
    std::vector<std::vector<Edge> > adj(2);
    std::vector<int> comp = { 0, 1 };
    std::mt19937_64 rng(seed * 7919ull + 13);
    long n_candidates = 0, n_edges_tried = 0, n_edges_valid = 0, calls = 0;
    int rounds = 0;
    bool solved = false;
    std::vector<double> q, qo, qa, qb; std::vector<int8_t> ac, ik, id; std::vector<uint8_t> ok, valid, res; std::vector<int32_t> nchk;
    {   /* one untimed call of the size the first round will use (device buffer allocation) */
        const size_t C = std::min<size_t>(1u << 20, 256 * (size_t)M + 4096);
        q.assign(C * 12, 0.0); qo.resize(C * 12); ac.assign(C, 0); ik.assign(C, 0); ok.resize(C); valid.resize(C);
        CK(ckc_pcs_validate(ctx, (int64_t)C, q.data(), ac.data(), ik.data(), qo.data(), ok.data(), valid.data()));
        uint8_t r1; int32_t n1; int8_t a0 = 0, k0 = (int8_t)V[0].ik[0];
        CK(ckc_pcs_check_motion_rbs(ctx, 1, start, start, &a0, &k0, &r1, &n1));
    }
    const auto t0 = std::chrono::steady_clock::now();
    while (!solved && rounds < max_rounds) {
        rounds++;
        /* ---- growRoadmap: M new valid milestones ---- */
        const size_t base = V.size();
        while (V.size() < base + (size_t)M) {
            /* ~0.5 % of the candidates are IK feasible, collision free and identifiable: ask for enough of them at once */
            const size_t C = std::min<size_t>(1u << 20, 256 * (base + (size_t)M - V.size()) + 4096);
            q.assign(C * 12, 0.0); qo.resize(C * 12); ac.assign(C, 0); ik.resize(C); ok.resize(C); valid.resize(C);
            for (size_t i = 0; i < C; i++) {
                for (int j = 0; j < 6; j++) q[12 * i + j] = uni(rng) * 2 * kPI_ - kPI_;            /* :203-204 */
                ik[i] = (int8_t)(rng() & 7);                                                       /* :206 */
            }
            CK(ckc_pcs_validate(ctx, (int64_t)C, q.data(), ac.data(), ik.data(), qo.data(), ok.data(), valid.data())); calls++;
            n_candidates += (long)C;
            std::vector<double> keep;
            for (size_t i = 0; i < C; i++) if (valid[i]) keep.insert(keep.end(), &qo[12 * i], &qo[12 * i] + 12);
            const size_t nk = keep.size() / 12;
            if (!nk) continue;
            id.resize(nk * 2);
            CK(ckc_identify_ik(ctx, (int64_t)nk, keep.data(), id.data())); calls++;
            for (size_t i = 0; i < nk && V.size() < base + (size_t)M; i++) {
                if (id[2 * i] < 0 && id[2 * i + 1] < 0) continue;                                  /* :211-214 */
                Vertex v; memcpy(v.q, &keep[12 * i], 96); v.ik[0] = id[2 * i]; v.ik[1] = id[2 * i + 1];
                V.push_back(v);
            }
        }
        adj.resize(V.size());
        for (size_t i = base; i < V.size(); i++) comp.push_back((int)i);
        /* ---- addMilestone: candidate edges to the K nearest earlier milestones ---- */
        struct Cand { int n, m; int state; };        /* state 0: try chain 0, 1: try chain 1, 2: connected, 3: rejected */
        std::vector<Cand> cand;
        std::vector<std::pair<double, int> > d;
        for (size_t m = base; m < V.size(); m++) {
            d.clear();
            for (size_t n = 0; n < m; n++) d.emplace_back(dist12(V[n].q, V[m].q), (int)n);
            const size_t k = std::min<size_t>(K, d.size());
            std::partial_sort(d.begin(), d.begin() + k, d.end());
            for (size_t i = 0; i < k; i++) cand.push_back({ d[i].second, (int)m, 0 });
        }
        std::vector<int> edge_ac(cand.size(), -1), edge_ik(cand.size(), -1);
        for (int pass = 0; pass < 2; pass++) {
            std::vector<int> idx;
            for (size_t c = 0; c < cand.size(); c++) {
                if (cand[c].state != pass) continue;
                const Vertex& a = V[cand[c].n]; const Vertex& b = V[cand[c].m];
                if (a.ik[pass] == b.ik[pass] && a.ik[pass] != -1) idx.push_back((int)c);           /* :546-549 */
                else cand[c].state = pass == 0 ? 1 : 3;
            }
            if (idx.empty()) continue;
            const size_t m = idx.size();
            qa.resize(m * 12); qb.resize(m * 12); ac.resize(m); ik.resize(m); res.resize(m); nchk.resize(m);
            for (size_t i = 0; i < m; i++) {
                const Cand& c = cand[idx[i]];
                memcpy(&qa[12 * i], V[c.n].q, 96); memcpy(&qb[12 * i], V[c.m].q, 96);
                ac[i] = (int8_t)pass; ik[i] = (int8_t)V[c.n].ik[pass];
            }
            CK(ckc_pcs_check_motion_rbs(ctx, (int64_t)m, qa.data(), qb.data(), ac.data(), ik.data(), res.data(), nchk.data())); calls++;
            n_edges_tried += (long)m;
            for (size_t i = 0; i < m; i++) {
                Cand& c = cand[idx[i]];
                if (res[i]) { c.state = 2; edge_ac[idx[i]] = pass; edge_ik[idx[i]] = V[c.n].ik[pass]; }
                else c.state = pass == 0 ? 1 : 3;
            }
        }
        for (size_t c = 0; c < cand.size(); c++) {
            if (cand[c].state != 2) continue;
            n_edges_valid++;
            const int a = cand[c].n, b = cand[c].m;
            const double w = dist12(V[a].q, V[b].q);
            adj[a].push_back({ b, w, edge_ac[c], edge_ik[c], 1 });
            adj[b].push_back({ a, w, edge_ac[c], edge_ik[c], 0 });
            const int ra = find(comp, a), rb = find(comp, b);
            if (ra != rb) comp[ra] = rb;                                                           /* uniteComponents :561 */
        }
        solved = find(comp, 0) == find(comp, 1);
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    FILE* fp = fopen(argv[4], "w");
    if (!fp) { fprintf(stderr, "cannot write %s\n", argv[4]); return 1; }
    size_t path_nodes = 0;
    if (solved) {
        /* cheapest path by 12-D length */
        const int n = (int)V.size();
        std::vector<double> dist(n, 1e300); std::vector<int> prev(n, -1), prev_e(n, -1);
        typedef std::pair<double, int> QE;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE> > pq;
        dist[0] = 0; pq.push({ 0.0, 0 });
        while (!pq.empty()) {
            const QE t = pq.top(); pq.pop();
            if (t.first > dist[t.second]) continue;
            for (size_t e = 0; e < adj[t.second].size(); e++) {
                const Edge& ed = adj[t.second][e];
                if (t.first + ed.w < dist[ed.to]) { dist[ed.to] = t.first + ed.w; prev[ed.to] = t.second; prev_e[ed.to] = (int)e; pq.push({ dist[ed.to], ed.to }); }
            }
        }
        std::vector<int> path;
        for (int v = 1; v >= 0; v = prev[v]) path.push_back(v);
        std::reverse(path.begin(), path.end());
        path_nodes = path.size();
        fprintf(fp, "query 0 nodes %zu rounds %d\n", path.size(), rounds);
        for (size_t i = 0; i < path.size(); i++) {
            const Vertex& v = V[path[i]];
            for (int j = 0; j < 12; j++) fprintf(fp, "%.17g ", v.q[j]);
            int e_ac = -1, e_ik = -1, e_dir = -1;
            if (i + 1 < path.size()) {
                const Edge& ed = adj[path[i]][prev_e[path[i + 1]]];
                e_ac = ed.ac; e_ik = ed.ik;
                /* the RBS ran from the smaller index (older milestone) to the larger one */
                e_dir = path[i] < path[i + 1] ? 0 : 1;
            }
            fprintf(fp, "0 %d %d %d %d %d\n", v.ik[0], v.ik[1], e_ac, e_ik, e_dir);
        }
    }
    fprintf(fp, "summary queries 1 solved %d seconds %.6f milestones %zu candidates %ld edges_tried %ld edges_valid %ld rounds %d gpu_calls %ld path_nodes %zu env %d\n",
            solved ? 1 : 0, secs, V.size(), n_candidates, n_edges_tried, n_edges_valid, rounds, calls, path_nodes, env);
    fclose(fp);
    printf("PRM-PCS env %d: %s in %.3f s, %zu milestones (%ld candidates), %ld of %ld edges valid, %d rounds, %ld batched calls, path of %zu milestones\n",
           env, solved ? "solved" : "NOT solved", secs, V.size(), n_candidates, n_edges_valid, n_edges_tried, rounds, calls, path_nodes);
    ckc_destroy(ctx);
    return solved ? 0 : 3;
}
