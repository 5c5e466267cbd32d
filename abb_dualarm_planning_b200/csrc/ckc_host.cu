/*
 * ckc_host.cu -- host side of libckc: context, mesh loading (.tris directory or packed .ckcmesh), OBB hierarchy
 * construction, device upload, batching/chunking and the extern "C" entry points of include/ckc.h.
 * Product code (no oracle includes).  There is deliberately no CPU compute path: every entry point runs the kernels
 * of ckc_kernels.cu or fails.
 */
#include "../../include/ckc.h"
#include "ckc_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#define CKC_VERSION_STR "ckc-b200 0.1 (sm_100a)"

/* ---------------------------------------------------------------------------------------------------------------- */
struct host_mesh {
    std::vector<double> tris;        /* ntris * 9 */
    std::vector<ckc_node> nodes;
    int ntris() const { return (int)(tris.size() / 9); }
};

struct dev_buf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

#define CKC_NUM_LANES 16          /* lanes that exist */
#define CKC_FAN_LANES 4           /* lanes the device-resident entry points fan out over */
struct ckc_lane {
    cudaStream_t st = nullptr;
    cudaStream_t st_hi = nullptr;       /* high-priority stream for the collision stage of this lane's chunk */
    cudaEvent_t ev_done = nullptr, ev_p = nullptr, ev_c = nullptr, ev_up = nullptr, ev_busy = nullptr;
    bool busy_set = false;              /* ev_busy has been recorded: the lane's scratch is in use by a device-resident call until it fires */
    dev_buf qin, qout, ac, ik, ok, valid, list, ovf, bigstack, flags, aux, aux2, small, gdstate, pairstat, smallio;
    int64_t collide_launches = 0;
    /* this lane's own first-hit statistics and processing order of the pair table: [0, MAX_PAIRS) hits, [MAX_PAIRS, 2 MAX_PAIRS)
     * order.  They are read and rewritten only by kernels on this lane's stream -- a table shared by the lanes would be
     * re-sorted on one stream while another lane's k_collide seeds its stacks from it (a torn permutation = a missed pair) */
    unsigned int* pair_hits() const { return pairstat.as<unsigned int>(); }
    int* pair_order() const { return pairstat.as<int>() + CKC_MAX_PAIRS; }
    int32_t* list_count() const { return small.as<int32_t>(); }
    int32_t* work_counter() const { return small.as<int32_t>() + 2; }      /* 2 ints */
    int32_t* ovf_count() const { return small.as<int32_t>() + 4; }
    int32_t* rbs_counts() const { return small.as<int32_t>() + 8; }        /* 2 ints */
    unsigned long long* gd_counter() const { return (unsigned long long*)(small.as<int32_t>() + 12); }
    void release() { dev_buf* b[] = { &qin, &qout, &ac, &ik, &ok, &valid, &list, &ovf, &bigstack, &flags, &aux, &aux2, &small, &gdstate, &pairstat, &smallio }; for (dev_buf* x : b) x->release(); }
};

struct ckc_ctx {
    /* multi-device context (ckc_create_multi): owns one ordinary context per device and nothing else; every host entry point
     * splits its batch into contiguous ranges, one per child, each driven by its own host thread */
    std::vector<ckc_ctx*> children;
    int env = 1, device = 0, num_sms = 148;
    std::string err;
    ckc_kin_dev kin;
    ckc_world_dev world;
    host_mesh meshes[CKC_NUM_MESHES];
    cudaStream_t stream = nullptr;
    int64_t chunk = 1 << 20;            /* samples per internal batch */
    /* persistent device storage */
    dev_buf d_nodes, d_tris, d_ctr, d_pairtab, d_static, d_pairhits, d_pairorder, d_bodies;
    dev_buf d_cidx, d_cq, d_ccount;                     /* compact outputs of the *_compact entry points (per chunk regions) */
    void* h_stage = nullptr; size_t h_stage_cap = 0;    /* pinned staging of the compact outputs */
    void* h_small = nullptr;                            /* pinned staging of the small-batch path (project_small) */
    dev_buf rbs_seg[2], rbs_cand, rbs_pool, rbs_eok, rbs_nck, rbs_hit, rbs_eac, rbs_eik, rbs_nt, rbs_in;      /* RBS frontier storage, kept between calls */
    /* per-chunk work buffers: CKC_NUM_LANES independent sets, each with its own stream, so that the host entry points
     * can overlap the H2D copy of chunk k+1, the kernels of chunk k and the D2H copy of chunk k-1 */
    ckc_lane lanes[CKC_NUM_LANES];
    int64_t pipe_chunk = 1 << 18;       /* largest pipelined chunk of the host entry points (samples); a call is cut into ~4 chunks
                                           of 2^16 .. pipe_chunk samples (measured: 2^18 beats 2^17 by 10 % on S-GD, 2^16 loses 20 %) */
    bool gd_fan = false;                /* fan GD batches out over the lanes on the _dev paths (CKC_GD_FAN) */
    bool pcs_fan = false;               /* the same for PCS batches (CKC_PCS_FAN).  Round 1 fanned them out (1.24e9 -> 1.9e9 configs/s); since the collision
                                           kernel finishes its heavy configurations inside the launch, one launch per batch is faster (2.08e9 -> 2.33e9), and the
                                           overlap between batches comes from the CALLER issuing on two streams (3.06e9).  A fanned-out call uses all lanes and
                                           the context's fork event: only one such call may be in flight per context */
    int pipe_taper = 0;                 /* the last chunk of a pipelined host call is halved this many times (CKC_PIPE_TAPER) */
    int pipe_taper_front = 0;           /* the first chunk halved this many times (CKC_PIPE_TAPER_FRONT) */
    int64_t pipe_min_chunk = 1 << 14;
    int pipe_nchunk = 4;                /* chunks a pipelined host call is cut into (CKC_PIPE_NCHUNK), bounded by pipe_chunk and 2^16 samples */
    int host_lanes = 4;                 /* lanes the host pipeline cycles through: a 1e6-sample call is 4 chunks, one per lane, so all
                                           H2D copies are in flight from the start; 8 lanes measured equal or slower (CKC_HOST_LANES) */
    bool overlap = true;                /* device-resident entry points fan out over the lanes */
    cudaEvent_t ev_fork = nullptr;
    cudaStream_t st_up = nullptr;       /* the host pipeline's uploads, in chunk order on ONE stream (ordered_upload) */
    bool ordered_upload = false;        /* CKC_ORDERED_UPLOAD: with the H2D copies of all chunks enqueued on their lanes' streams the copy engines
                                           share the link between them and every chunk arrives late; on one stream chunk 0 arrives first */
    bool timeline_copies = false;       /* CKC_TIMELINE: the stage timeline also carries the uploads (as stage 1) */
    int64_t launches = 0;
    unsigned dev_rot = 0;               /* lane rotation of the device-resident entry points */
    ckc_stats host_stats;
    /* optional per-stage CUDA-event timing */
    bool stage_timing = false;
    struct stage_ev { int stage; cudaEvent_t e0, e1; };
    std::vector<stage_ev> pending;
    double stage_ms[4] = { 0, 0, 0, 0 };
    int64_t stage_n[4] = { 0, 0, 0, 0 };
};

/* RAII helper: records an event pair around a stage when timing is enabled */
struct stage_scope {
    ckc_ctx* c; cudaStream_t st; int stage; cudaEvent_t e0 = nullptr, e1 = nullptr;
    stage_scope(ckc_ctx* c_, cudaStream_t st_, int stage_) : c(c_), st(st_), stage(stage_) {
        if (!c->stage_timing) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
        cudaEventRecord(e0, st);
    }
    ~stage_scope() {
        if (!e0) return;
        cudaEventRecord(e1, st);
        c->pending.push_back({ stage, e0, e1 });
    }
};

static int fail(ckc_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, CKC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

/* A failure in the middle of a pipelined call must not return while earlier chunks still copy into the caller's buffers
 * on other lanes: drain every lane first (errors of the drain itself are ignored, the first failure is what gets reported). */
static void drain_lanes(ckc_ctx* ctx);
#define CUS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { drain_lanes(ctx); return fail(ctx, CKC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)

/* ---------------------------------------------------------------------------------------------------------------- */
/* mesh input                                                                                                         */
/* ---------------------------------------------------------------------------------------------------------------- */
static const char* const k_mesh_file[CKC_NUM_MESHES] = {   /* the files collisionDetection::load_models opens (:502-923) */
    "Base_r.tris", "Link1_r.tris", "Link2_r.tris", "Link3_r.tris", "Link4_r2.tris", "Link5_r.tris", "Link6.tris",
    "EE_r.tris", "table.tris", "obs.tris", "cone.tris", nullptr };

static bool read_tris_file(const std::string& path, std::vector<double>& out) {
    FILE* fp = fopen(path.c_str(), "r");
    if (!fp) return false;
    int n = 0;
    bool ok = fscanf(fp, "%d", &n) == 1 && n >= 0 && n <= (1 << 20);      /* header sanity: no unbounded resize */
    if (ok) {
        out.resize((size_t)n * 9);
        for (size_t i = 0; ok && i < out.size(); i++) ok = fscanf(fp, "%lf", &out[i]) == 1;
    }
    fclose(fp);
    return ok;
}

static bool read_pack(const std::string& path, host_mesh* meshes) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    char magic[8]; int32_t nm = 0;
    bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "CKCMESH1", 8) == 0 && fread(&nm, 4, 1, fp) == 1;
    for (int m = 0; ok && m < nm; m++) {
        int32_t id = -1, nt = 0;
        ok = fread(&id, 4, 1, fp) == 1 && fread(&nt, 4, 1, fp) == 1 && id >= 0 && id < CKC_NUM_MESHES && nt >= 0 && nt <= (1 << 20);
        if (!ok) break;
        meshes[id].tris.resize((size_t)nt * 9);
        ok = fread(meshes[id].tris.data(), 8, (size_t)nt * 9, fp) == (size_t)nt * 9;
    }
    fclose(fp);
    return ok;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* OBB hierarchy: top-down, PCA-oriented boxes, median split along the longest box axis, one triangle per leaf.        */
/* Any conservative hierarchy yields the reference's boolean (PQP's own RSS tree only prunes).                        */
/* ---------------------------------------------------------------------------------------------------------------- */
static void sym_eig3(double A[3][3], double V[3][3]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = i == j;
    for (int sweep = 0; sweep < 64; sweep++) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off < 1e-28) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; k++) { const double x = A[k][p], y = A[k][q]; A[k][p] = c * x - s * y; A[k][q] = s * x + c * y; }
                for (int k = 0; k < 3; k++) { const double x = A[p][k], y = A[q][k]; A[p][k] = c * x - s * y; A[q][k] = s * x + c * y; }
                for (int k = 0; k < 3; k++) { const double x = V[k][p], y = V[k][q]; V[k][p] = c * x - s * y; V[k][q] = s * x + c * y; }
            }
    }
}

struct builder {
    host_mesh& m;
    bool fit_refine, split_best, size_area, split_sah, fit_edges; int nfrac;
    explicit builder(host_mesh& mm) : m(mm) {
        /* bit 0: in-plane refit, 1: best-of-three split axis, 2: descend by surface area, 3: binned split positions, 4: edge-aligned candidates */
        const char* e = getenv("CKC_BVH");
        const int f = e ? atoi(e) : 31;
        fit_refine = (f & 1) != 0; split_best = (f & 2) != 0; size_area = (f & 4) != 0; split_sah = (f & 8) != 0; fit_edges = (f & 16) != 0; nfrac = (f & 32) ? 5 : 3;
    }
    static double area(const ckc_node& n) { return (double)n.h[0] * n.h[1] + (double)n.h[2] * ((double)n.h[0] + n.h[1]); }
    const double* vert(int t, int v) const { return &m.tris[9 * (size_t)t + 3 * v]; }

    void fit(const std::vector<int>& idx, ckc_node& n, double axes[3][3]) {
        double mean[3] = { 0, 0, 0 };
        for (int t : idx) for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) mean[k] += vert(t, v)[k];
        for (int k = 0; k < 3; k++) mean[k] /= 3.0 * idx.size();
        double C[3][3] = { { 0 } };
        for (int t : idx) for (int v = 0; v < 3; v++) {
            double d[3]; for (int k = 0; k < 3; k++) d[k] = vert(t, v)[k] - mean[k];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[i][j] += d[i] * d[j];
        }
        double V[3][3]; sym_eig3(C, V);
        /* order the eigenvectors by decreasing eigenvalue, Gram-Schmidt, right-handed */
        int ord[3] = { 0, 1, 2 };
        std::sort(ord, ord + 3, [&](int a, int b) { return C[a][a] > C[b][b]; });
        double a0[3] = { V[0][ord[0]], V[1][ord[0]], V[2][ord[0]] }, a1[3] = { V[0][ord[1]], V[1][ord[1]], V[2][ord[1]] }, a2[3];
        auto nrm = [](double* a) { double l = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); if (l > 0) { a[0] /= l; a[1] /= l; a[2] /= l; } return l; };
        if (nrm(a0) < 1e-12) { a0[0] = 1; a0[1] = 0; a0[2] = 0; }
        double d = a0[0] * a1[0] + a0[1] * a1[1] + a0[2] * a1[2];
        for (int k = 0; k < 3; k++) a1[k] -= d * a0[k];
        if (nrm(a1) < 1e-9) {
            double e[3] = { fabs(a0[0]) < 0.9 ? 1.0 : 0.0, fabs(a0[0]) < 0.9 ? 0.0 : 1.0, 0 };
            a1[0] = a0[1] * e[2] - a0[2] * e[1]; a1[1] = a0[2] * e[0] - a0[0] * e[2]; a1[2] = a0[0] * e[1] - a0[1] * e[0];
            nrm(a1);
        }
        /* PCA fixes the plane of least spread well but not the rotation inside it: scan the in-plane angle for the smallest
         * rectangle (for a single triangle that is the one standing on its longest edge, half the area a skewed box can have) */
        if (fit_refine) {
            double a2p[3] = { a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0] };
            const double* pca[3] = { a0, a1, a2p };
            double best = 1e300, bu[3] = { a0[0], a0[1], a0[2] }, bv[3] = { a1[0], a1[1], a1[2] };
            const int steps = 48;
            /* fixed axis = the PCA axis of least spread; fit_edges: each of the three PCA axes in turn */
            for (int fix = 2; fix >= (fit_edges ? 0 : 2); fix--) {
                const double* e0 = pca[(fix + 1) % 3]; const double* e1 = pca[(fix + 2) % 3]; const double* n2 = pca[fix];
                for (int k = 0; k < steps; k++) {
                    const double th = (M_PI / 2) * k / steps, c = cos(th), sn = sin(th);
                    double u[3], v[3];
                    for (int i = 0; i < 3; i++) { u[i] = c * e0[i] + sn * e1[i]; v[i] = -sn * e0[i] + c * e1[i]; }
                    double lu = 1e300, hu = -1e300, lv = 1e300, hv = -1e300, ln = 1e300, hn = -1e300;
                    for (int t : idx) for (int vv = 0; vv < 3; vv++) {
                        const double* pt = vert(t, vv);
                        const double pu = pt[0] * u[0] + pt[1] * u[1] + pt[2] * u[2], pv = pt[0] * v[0] + pt[1] * v[1] + pt[2] * v[2];
                        const double pn = pt[0] * n2[0] + pt[1] * n2[1] + pt[2] * n2[2];
                        lu = std::min(lu, pu); hu = std::max(hu, pu); lv = std::min(lv, pv); hv = std::max(hv, pv); ln = std::min(ln, pn); hn = std::max(hn, pn);
                    }
                    /* surface area of the box (the traversal cost of a bounding volume scales with its area, not its volume) */
                    const double du = hu - lu, dv = hv - lv, dn = hn - ln;
                    const double cost = du * dv + dn * (du + dv);
                    if (cost < best * (1 - 1e-9)) { best = cost; for (int i = 0; i < 3; i++) { bu[i] = u[i]; bv[i] = v[i]; } }
                }
            }
            for (int i = 0; i < 3; i++) { a0[i] = bu[i]; a1[i] = bv[i]; }
        }
        /* the kernel rebuilds a2 = a0 x a1 from the FP32 copies: round first so that host extents match what it sees */
        float f0[3], f1[3];
        for (int k = 0; k < 3; k++) { f0[k] = (float)a0[k]; f1[k] = (float)a1[k]; a0[k] = f0[k]; a1[k] = f1[k]; }
        a2[0] = a0[1] * a1[2] - a0[2] * a1[1]; a2[1] = a0[2] * a1[0] - a0[0] * a1[2]; a2[2] = a0[0] * a1[1] - a0[1] * a1[0];
        const double* ax[3] = { a0, a1, a2 };
        double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
        for (int t : idx) for (int v = 0; v < 3; v++)
            for (int i = 0; i < 3; i++) {
                const double* p = vert(t, v);
                const double pr = p[0] * ax[i][0] + p[1] * ax[i][1] + p[2] * ax[i][2];
                lo[i] = std::min(lo[i], pr); hi[i] = std::max(hi[i], pr);
            }
        /* the FP32 axes are not exactly orthonormal (1e-7): pad the extents by a relative + absolute slack so that the
         * box still contains every vertex when evaluated in FP32 */
        double cworld[3] = { 0, 0, 0 };
        for (int i = 0; i < 3; i++) {
            const double mid = 0.5 * (lo[i] + hi[i]);
            for (int k = 0; k < 3; k++) cworld[k] += mid * ax[i][k];
        }
        double ext = 0;
        for (int i = 0; i < 3; i++) ext = std::max(ext, 0.5 * (hi[i] - lo[i]));
        double cmag = std::max(std::max(fabs(cworld[0]), fabs(cworld[1])), fabs(cworld[2]));
        const double slack = 1e-3 + 4e-6 * (ext + cmag);
        for (int i = 0; i < 3; i++) { n.h[i] = (float)(0.5 * (hi[i] - lo[i]) + slack); n.c[i] = (float)cworld[i]; }
        for (int k = 0; k < 3; k++) { n.a0[k] = f0[k]; n.a1[k] = f1[k]; }
        n.size2 = size_area ? (float)area(n) : n.h[0] * n.h[0] + n.h[1] * n.h[1] + n.h[2] * n.h[2];
        n.pad0 = n.pad1 = 0;
        for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) axes[i][k] = ax[i][k];
    }

    void build(std::vector<int>& idx, int self) {
        double axes[3][3];
        ckc_node n; memset(&n, 0, sizeof(n));
        fit(idx, n, axes);
        if (idx.size() == 1) { n.child = ~idx[0]; m.nodes[self] = n; return; }
        /* median split by triangle centroid along a box axis: the longest one, or (split_best) the one of the three whose two
         * child boxes have the smallest total surface area */
        int axis = 0; if (n.h[1] > n.h[axis]) axis = 1; if (n.h[2] > n.h[axis]) axis = 2;
        std::vector<int> L, R;
        auto split = [&](int ax, std::vector<int>& l, std::vector<int>& r, double frac = 0.5) {
            const double* a = axes[ax];
            std::vector<std::pair<double, int>> key(idx.size());
            for (size_t i = 0; i < idx.size(); i++) {
                double c = 0;
                for (int v = 0; v < 3; v++) { const double* p = vert(idx[i], v); c += p[0] * a[0] + p[1] * a[1] + p[2] * a[2]; }
                key[i] = { c, idx[i] };
            }
            std::sort(key.begin(), key.end());
            const size_t half = std::max<size_t>(1, std::min<size_t>(idx.size() - 1, (size_t)(idx.size() * frac)));
            l.clear(); r.clear();
            for (size_t i = 0; i < key.size(); i++) (i < half ? l : r).push_back(key[i].second);
        };
        if (split_best && idx.size() > 2) {
            double best = 1e300;
            const double fr[5] = { 0.5, 0.375, 0.625, 0.25, 0.75 };
            for (int ax = 0; ax < 3; ax++)
                for (int f = 0; f < (split_sah && idx.size() >= 8 ? nfrac : 1); f++) {
                    std::vector<int> l, r; split(ax, l, r, fr[f]);
                    ckc_node nl, nr; double tmp[3][3];
                    memset(&nl, 0, sizeof(nl)); memset(&nr, 0, sizeof(nr));
                    fit(l, nl, tmp); fit(r, nr, tmp);
                    const double cost = split_sah ? area(nl) * l.size() + area(nr) * r.size() : area(nl) + area(nr);
                    if (cost < best) { best = cost; L.swap(l); R.swap(r); }
                }
        } else split(axis, L, R);
        const int child = (int)m.nodes.size();
        m.nodes.resize(m.nodes.size() + 2);          /* siblings are adjacent: right = left + 1 */
        n.child = child;
        m.nodes[self] = n;
        build(L, child);
        build(R, child + 1);
    }
};

static void build_hierarchy(host_mesh& m) {
    m.nodes.clear();
    if (m.ntris() == 0) return;
    std::vector<int> idx(m.ntris());
    for (int i = 0; i < m.ntris(); i++) idx[i] = i;
    m.nodes.resize(1);
    builder b(m);
    b.build(idx, 0);
}

/* the rod: setP() (StateValidityCheckerPCS.h:150-158) + collision_state :50-60 -- points (20 i, 0, 0), i = 0..L/20, taken
 * three at a time as (degenerate) triangles */
static void make_rod(host_mesh& m, double L) {
    const int dl = 20, n = (int)(L / dl);
    const int ntri = (n + 1) / 3;
    m.tris.clear();
    for (int t = 0; t < ntri; t++)
        for (int v = 0; v < 3; v++) { m.tris.push_back((double)((3 * t + v) * dl)); m.tris.push_back(0); m.tris.push_back(0); }
}

/* pair table of collision_state (collisionDetection.cpp:267-355, env 1 :372-420, env 2 :443-472) */
static void build_pair_table(ckc_world_dev& w, int env) {
    const uint8_t lm[8] = { CKC_M_BASE, CKC_M_LINK1, CKC_M_LINK2, CKC_M_LINK3, CKC_M_LINK4, CKC_M_LINK5, CKC_M_LINK6, CKC_M_EE };
    int n = 0;
    auto add = [&](int ma, int fa, int mb, int fb) { w.pairs[n].mesh_a = (uint8_t)ma; w.pairs[n].frame_a = (uint8_t)fa; w.pairs[n].mesh_b = (uint8_t)mb; w.pairs[n].frame_b = (uint8_t)fb; n++; };
    for (int r = 0; r < 2; r++) {                       /* self collision of each arm */
        const int o = 8 * r;
        for (int b = 4; b <= 6; b++) for (int a = 0; a <= 2; a++) add(lm[a], o + a, lm[b], o + b);
        add(CKC_M_BASE, o, CKC_M_LINK3, o + 3);
        for (int a = 0; a <= 2; a++) add(lm[a], o + a, CKC_M_EE, o + 7);
    }
    for (int a = 3; a <= 6; a++) for (int b = 3; b <= 6; b++) add(lm[a], a, lm[b], 8 + b);     /* arm-arm */
    for (int a = 3; a <= 6; a++) add(lm[a], a, CKC_M_EE, 15);
    for (int b = 3; b <= 6; b++) add(CKC_M_EE, 7, lm[b], 8 + b);
    for (int r = 0; r < 2; r++) for (int b = 0; b <= 6; b++) add(CKC_M_ROD, 7, lm[b], 8 * r + b);   /* rod-arm */
    for (int r = 0; r < 2; r++) for (int b = 2; b <= 7; b++) add(CKC_M_TABLE, 0, lm[b], 8 * r + b); /* table-arm */
    add(CKC_M_TABLE, 0, CKC_M_ROD, 7);
    const int nobs = env == 1 ? 3 : 2, om = env == 1 ? CKC_M_OBS : CKC_M_CONE;
    for (int k = 0; k < nobs; k++) {
        const int fo = 18 + k;
        add(om, fo, CKC_M_ROD, 7);
        add(om, fo, CKC_M_LINK2, k == 0 ? 16 : 2);       /* res[78] poses link2 with (R22, T2_t) */
        for (int b = 3; b <= 7; b++) add(om, fo, lm[b], b);
        add(om, fo, CKC_M_LINK2, k == 0 ? 10 : 17);      /* res[97], res[110] pose link22 with (R2, T2_t2) */
        for (int b = 3; b <= 7; b++) add(om, fo, lm[b], 8 + b);
    }
    w.npairs = n;
}

static void rotz(double M[9], double t) { const double c = cos(t), s = sin(t); M[0] = c; M[1] = -s; M[2] = 0; M[3] = s; M[4] = c; M[5] = 0; M[6] = 0; M[7] = 0; M[8] = 1; }
static void mul33(double C[9], const double A[9], const double B[9]) {
    double R[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]);
    memcpy(C, R, sizeof(R));
}

static void init_constants(ckc_ctx* c) {
    ckc_kin_dev& k = c->kin;
    const int env = c->env;
    k.b = 166; k.l1 = 124; k.l2 = 270; k.l3a = 70; k.l3b = 149.6; k.l4 = 152.4;
    k.l5 = 72 - 13. / 2; k.lee = 60 + 13. / 2;                                    /* apc_class.cpp:11-12 */
    k.D = env == 1 ? 900. : 1200.; k.L = env == 1 ? 300. : 500.;                 /* StateValidityCheckerPCS.h:23-26 */
    const double deg[6][2] = { { -165, 165 }, { -110, 110 }, { -110, 70 }, { -160, 160 }, { -120, 120 }, { -400, 400 } };
    for (int j = 0; j < 6; j++) { k.lim_lo[j] = deg[j][0] * 3.1416 / 180.0; k.lim_hi[j] = deg[j][1] * 3.1416 / 180.0; }   /* deg2rad uses PI_ */
    k.base_x[0] = -k.D / 2; k.base_y[0] = 0; k.base_c[0] = cos(0.0); k.base_s[0] = sin(0.0);
    k.base_x[1] = k.D / 2; k.base_y[1] = 0; k.base_c[1] = cos(3.1416); k.base_s[1] = sin(3.1416);                       /* PCS.h:37 */
    k.l34 = sqrt((k.l3a * k.l3a + (k.l3b + k.l4) * (k.l3b + k.l4)));
    k.alpha = atan2(k.l3b + k.l4, k.l3a);
    /* KDL chain (kdl_class.cpp:11-12,32-45): integer-division link lengths */
    const double kl5 = 72 - 13 / 2, klee = 60 + 13 / 2;
    const double tips[13][3] = { { 0, 0, k.b }, { 0, 0, k.l1 }, { 0, 0, k.l2 }, { k.l3b, 0, k.l3a }, { k.l4, 0, 0 }, { kl5, 0, 0 }, { 2 * klee + k.L, 0, 0 },
                                 { kl5, 0, 0 }, { k.l4, 0, 0 }, { k.l3b, 0, -k.l3a }, { 0, 0, -k.l2 }, { 0, 0, -k.l1 }, { 0, 0, -k.b } };
    const int axes[13] = { -1, 2, 1, 1, 0, 1, 0, 0, 1, 0, 1, 1, 2 };
    memcpy(k.kdl_tip, tips, sizeof(tips)); memcpy(k.kdl_axis, axes, sizeof(axes));
    k.env_index = env - 1;

    ckc_world_dev& w = c->world;
    w.env = env; w.D = k.D; w.tol = 8.0; w.cull_tol = 8.0f + 0.05f;
    /* robot 2 base: M02 = Rz(3.14159265/2 + offsetRot), R02 = M02 * M02 (collisionDetection.cpp:177-178) */
    double M[9]; rotz(M, 3.14159265 / 2 + 0.0); mul33(w.base2_rot, M, M);
    /* T0 = (-offsetX/2, offsetY/2, offsetZ/2) then MxV(T0, R0, T0) IN PLACE (:99-103, :203-207): components computed
     * later read the already overwritten earlier ones */
    double I[9]; rotz(M, 0.0); mul33(I, M, M);
    const double* Rb[2] = { I, w.base2_rot };
    for (int r = 0; r < 2; r++) {
        double T[3] = { -k.D / 2, 0.0 / 2, 0.0 / 2 };
        const double* R = Rb[r];
        T[0] = (R[0] * T[0] + R[1] * T[1] + R[2] * T[2]);
        T[1] = (R[3] * T[0] + R[4] * T[1] + R[5] * T[2]);
        T[2] = (R[6] * T[0] + R[7] * T[1] + R[8] * T[2]);
        memcpy(w.base_t[r], T, sizeof(T));
    }
    /* static obstacle poses (:363-406 env 1; :431-456 env 2) */
    for (int o = 0; o < 3; o++) { memset(w.static_frames[o], 0, sizeof(w.static_frames[o])); w.static_frames[o][0] = w.static_frames[o][4] = w.static_frames[o][8] = 1; }
    if (env == 1) {
        w.static_frames[1][9] = 250; w.static_frames[1][10] = 250;
        w.static_frames[2][9] = -250; w.static_frames[2][10] = -250;
    } else {
        const double cx = cos(3.14), sx = sin(3.14);                              /* MRotX(Mobs, 3.14) */
        double* F = w.static_frames[1];
        F[0] = 1; F[1] = 0; F[2] = 0; F[3] = 0; F[4] = cx; F[5] = -sx; F[6] = 0; F[7] = sx; F[8] = cx; F[11] = 790;
    }
    build_pair_table(w, env);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* lifetime                                                                                                           */
/* ---------------------------------------------------------------------------------------------------------------- */
extern "C" const char* ckc_version(void) { return CKC_VERSION_STR; }
extern "C" const char* ckc_last_error(const ckc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int ckc_num_pairs(const ckc_ctx* ctx) { return ctx ? ctx->world.npairs : CKC_ERR_ARG; }

static thread_local std::string g_create_error;
extern "C" const char* ckc_create_error(void) { return g_create_error.c_str(); }
extern "C" int ckc_create(ckc_ctx** out, int env, const char* mesh_path, int device);

extern "C" int ckc_create(ckc_ctx** out, int env, const char* mesh_path, int device) {
    if (!out || (env != 1 && env != 2) || !mesh_path) { g_create_error = "bad argument"; return CKC_ERR_ARG; }
    *out = nullptr;
    ckc_ctx* ctx = new ckc_ctx();
    ctx->env = env; ctx->device = device;
    memset(&ctx->host_stats, 0, sizeof(ctx->host_stats));
    auto bail = [&](int code, const std::string& msg) { g_create_error = msg; delete ctx; return code; };

    struct stat st;
    if (stat(mesh_path, &st) != 0) return bail(CKC_ERR_IO, std::string("cannot stat ") + mesh_path);
    if (S_ISDIR(st.st_mode)) {
        for (int m = 0; m < CKC_NUM_MESHES; m++) {
            if (!k_mesh_file[m]) continue;
            if (m == CKC_M_OBS && env != 1) continue;
            if (m == CKC_M_CONE && env != 2) continue;
            if (!read_tris_file(std::string(mesh_path) + "/" + k_mesh_file[m], ctx->meshes[m].tris))
                return bail(CKC_ERR_IO, std::string("Couldn't open ") + k_mesh_file[m]);
        }
    } else if (!read_pack(mesh_path, ctx->meshes)) return bail(CKC_ERR_IO, std::string("bad mesh pack ") + mesh_path);
    init_constants(ctx);
    make_rod(ctx->meshes[CKC_M_ROD], ctx->kin.L);
    for (int m = 0; m < CKC_NUM_MESHES; m++) build_hierarchy(ctx->meshes[m]);
    /* the kernels pack node and triangle indices into 11 / 10 bits of a task word */
    static_assert(CKC_MAX_MESH_NODES == (1 << 11) && CKC_MAX_MESH_TRIS == (1 << 10) && CKC_MAX_PAIRS <= (1 << 10), "task word bit fields");
    for (int m = 0; m < CKC_NUM_MESHES; m++)
        if (ctx->meshes[m].ntris() > CKC_MAX_MESH_TRIS || (int)ctx->meshes[m].nodes.size() > CKC_MAX_MESH_NODES)
            return bail(CKC_ERR_ARG, std::string("mesh ") + (k_mesh_file[m] ? k_mesh_file[m] : "rod") + " has more than 1024 triangles (task word bit fields)");
    for (int p = 0; p < ctx->world.npairs; p++)
        if (ctx->meshes[ctx->world.pairs[p].mesh_a].ntris() == 0 || ctx->meshes[ctx->world.pairs[p].mesh_b].ntris() == 0)
            return bail(CKC_ERR_IO, "a mesh used by the pair table is empty");

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return bail(CKC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return bail(CKC_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    ctx->num_sms = prop.multiProcessorCount;
    for (int l = 0; l < CKC_NUM_LANES; l++) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        e = cudaStreamCreateWithPriority(&ctx->lanes[l].st, cudaStreamNonBlocking, prio_lo);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->lanes[l].st_hi, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lanes[l].ev_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lanes[l].ev_p, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lanes[l].ev_c, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lanes[l].ev_up, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lanes[l].ev_busy, cudaEventDisableTiming);
        if (e == cudaSuccess && l == 0) e = cudaStreamCreateWithFlags(&ctx->st_up, cudaStreamNonBlocking);
        if (e == cudaSuccess && l == 0) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = ctx->lanes[l].small.reserve(256);
        if (e == cudaSuccess) e = cudaMemset(ctx->lanes[l].small.p, 0, 256);
        if (e != cudaSuccess) { std::string msg = std::string("lane setup: ") + cudaGetErrorString(e); ckc_destroy(ctx); g_create_error = msg; return CKC_ERR_CUDA; }
    }
    ctx->stream = ctx->lanes[0].st;

    /* concatenate and upload */
    std::vector<ckc_node> nodes; std::vector<double> tris;
    for (int m = 0; m < CKC_NUM_MESHES; m++) {
        ctx->world.node_off[m] = (int)nodes.size(); ctx->world.tri_off[m] = (int)(tris.size() / 9);
        nodes.insert(nodes.end(), ctx->meshes[m].nodes.begin(), ctx->meshes[m].nodes.end());
        tris.insert(tris.end(), ctx->meshes[m].tris.begin(), ctx->meshes[m].tris.end());
    }
    bool okc = ctx->d_nodes.reserve(nodes.size() * sizeof(ckc_node)) == cudaSuccess && ctx->d_tris.reserve(tris.size() * 8) == cudaSuccess &&
               ctx->d_ctr.reserve(sizeof(ckc_counters_dev)) == cudaSuccess;
    if (okc) okc = cudaMemcpy(ctx->d_nodes.p, nodes.data(), nodes.size() * sizeof(ckc_node), cudaMemcpyHostToDevice) == cudaSuccess &&
                   cudaMemcpy(ctx->d_tris.p, tris.data(), tris.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
                   cudaMemset(ctx->d_ctr.p, 0, sizeof(ckc_counters_dev)) == cudaSuccess;
    if (!okc) { std::string msg = std::string("device upload: ") + cudaGetErrorString(cudaGetLastError()); ckc_destroy(ctx); g_create_error = msg; return CKC_ERR_CUDA; }
    ctx->world.nodes = ctx->d_nodes.as<ckc_node>();
    ctx->world.tris = ctx->d_tris.as<double>();
    {   /* pair table and static poses as global-memory tables */
        std::vector<ckc_pair_dev> pt(ctx->world.npairs);
        std::vector<std::pair<int, int> > body_key;
        for (int p = 0; p < ctx->world.npairs; p++) {
            const ckc_pair& pr = ctx->world.pairs[p];
            pt[p].node_a = ctx->world.node_off[pr.mesh_a]; pt[p].node_b = ctx->world.node_off[pr.mesh_b];
            pt[p].tri_a = ctx->world.tri_off[pr.mesh_a]; pt[p].tri_b = ctx->world.tri_off[pr.mesh_b];
            pt[p].frame_a = pr.frame_a; pt[p].frame_b = pr.frame_b; pt[p].pad1 = 0;
            /* root pre-test of k_collide: two spheres around the root boxes; radius sum + tolerance + FP32 slack, as float bits */
            auto radius = [&](int mesh) {          /* smallest sphere about the root box centre that holds every vertex (<= |half extents|) */
                const host_mesh& hm = ctx->meshes[mesh]; const ckc_node& r = hm.nodes[0];
                double r2 = 0;
                for (size_t v = 0; v + 2 < hm.tris.size(); v += 3) {
                    const double dx = hm.tris[v] - r.c[0], dy = hm.tris[v + 1] - r.c[1], dz = hm.tris[v + 2] - r.c[2];
                    r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
                }
                return sqrt(r2);
            };
            const float rsum = (float)((radius(pr.mesh_a) + radius(pr.mesh_b) + ctx->world.cull_tol) * (1 + 1e-5) + 2e-3);
            memcpy(&pt[p].pad0, &rsum, 4);
            /* the two bodies of the entry: distinct (frame, mesh) combinations, numbered in order of appearance */
            int bid[2];
            for (int side = 0; side < 2; side++) {
                const int fr = side == 0 ? pr.frame_a : pr.frame_b, me = side == 0 ? pr.mesh_a : pr.mesh_b;
                int found = -1;
                for (size_t b = 0; b < body_key.size(); b++) if (body_key[b].first == fr && body_key[b].second == me) { found = (int)b; break; }
                if (found < 0) { found = (int)body_key.size(); body_key.push_back(std::make_pair(fr, me)); }
                bid[side] = found;
            }
            pt[p].pad1 = bid[0] | (bid[1] << 8);
        }
        if (body_key.size() > 32) { ckc_destroy(ctx); g_create_error = "pair table names more than 32 (frame, mesh) bodies"; return CKC_ERR_ARG; }
        {
            std::vector<float> bt(4 * 32, 0.0f);
            for (size_t b = 0; b < body_key.size(); b++) {
                const ckc_node& r = ctx->meshes[body_key[b].second].nodes[0];
                bt[4 * b] = r.c[0]; bt[4 * b + 1] = r.c[1]; bt[4 * b + 2] = r.c[2];
                const int fr = body_key[b].first; memcpy(&bt[4 * b + 3], &fr, 4);
            }
            if (ctx->d_bodies.reserve(bt.size() * 4) != cudaSuccess || cudaMemcpy(ctx->d_bodies.p, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
                std::string msg = std::string("device upload: ") + cudaGetErrorString(cudaGetLastError()); ckc_destroy(ctx); g_create_error = msg; return CKC_ERR_CUDA;
            }
            ctx->world.body_tab = ctx->d_bodies.as<float4>(); ctx->world.nbodies = (int)body_key.size();
        }
        /* obstacle poses, then the base frames of the two robots: robot 1 R0 = Rz(0)^2 = I, robot 2 R02 = Rz(3.14159265/2)^2 */
        double stat[5][12];
        memcpy(stat, ctx->world.static_frames, sizeof(ctx->world.static_frames));
        for (int r = 0; r < 2; r++) {
            for (int e = 0; e < 9; e++) stat[3 + r][e] = r == 0 ? (e % 4 == 0 ? 1.0 : 0.0) : ctx->world.base2_rot[e];
            for (int e = 0; e < 3; e++) stat[3 + r][9 + e] = ctx->world.base_t[r][e];
        }
        bool okt = ctx->d_pairtab.reserve(pt.size() * sizeof(ckc_pair_dev)) == cudaSuccess && ctx->d_static.reserve(sizeof(stat)) == cudaSuccess &&
                   cudaMemcpy(ctx->d_pairtab.p, pt.data(), pt.size() * sizeof(ckc_pair_dev), cudaMemcpyHostToDevice) == cudaSuccess &&
                   cudaMemcpy(ctx->d_static.p, stat, sizeof(stat), cudaMemcpyHostToDevice) == cudaSuccess;
        if (!okt) { std::string msg = std::string("device upload: ") + cudaGetErrorString(cudaGetLastError()); ckc_destroy(ctx); g_create_error = msg; return CKC_ERR_CUDA; }
        ctx->world.pair_tab = ctx->d_pairtab.as<ckc_pair_dev>();
        ctx->world.static_frames_dev = ctx->d_static.as<double>();
        /* processing-order prior: obstacles, table, arm-arm, rod, self collision; refined on line from the hit statistics */
        std::vector<unsigned int> prior(CKC_MAX_PAIRS, 0); std::vector<int> ord(CKC_MAX_PAIRS, 0);
        for (int p = 0; p < ctx->world.npairs; p++) prior[p] = p >= 77 ? 4000 : (p >= 64 ? 2500 : (p >= 26 && p < 50 ? 1500 : (p >= 50 ? 800 : 400)));
        std::vector<int> idx(ctx->world.npairs);
        for (int p = 0; p < ctx->world.npairs; p++) idx[p] = p;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return prior[a] > prior[b]; });
        for (int p = 0; p < ctx->world.npairs; p++) ord[p] = idx[p];
        okt = ctx->d_pairhits.reserve(CKC_MAX_PAIRS * 4) == cudaSuccess && ctx->d_pairorder.reserve(CKC_MAX_PAIRS * 4) == cudaSuccess &&
              cudaMemcpy(ctx->d_pairhits.p, prior.data(), CKC_MAX_PAIRS * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(ctx->d_pairorder.p, ord.data(), CKC_MAX_PAIRS * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        if (!okt) { std::string msg = std::string("device upload: ") + cudaGetErrorString(cudaGetLastError()); ckc_destroy(ctx); g_create_error = msg; return CKC_ERR_CUDA; }
        ctx->world.pair_hits = ctx->d_pairhits.as<unsigned int>();
        ctx->world.pair_order = ctx->d_pairorder.as<int>();
        for (int l = 0; l < CKC_NUM_LANES && okt; l++) {
            ckc_lane& L = ctx->lanes[l];
            okt = L.pairstat.reserve(2 * CKC_MAX_PAIRS * 4) == cudaSuccess &&
                  cudaMemcpy(L.pair_hits(), prior.data(), CKC_MAX_PAIRS * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(L.pair_order(), ord.data(), CKC_MAX_PAIRS * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        }
        if (!okt) { std::string msg = std::string("device upload: ") + cudaGetErrorString(cudaGetLastError()); ckc_destroy(ctx); g_create_error = msg; return CKC_ERR_CUDA; }
    }
    const char* ch = getenv("CKC_CHUNK");
    if (ch && atoll(ch) >= 1024) ctx->chunk = atoll(ch);
    ch = getenv("CKC_OVERLAP");
    if (ch) ctx->overlap = atoi(ch) != 0;
    ch = getenv("CKC_PIPE_CHUNK");
    if (ch && atoll(ch) >= 1024) ctx->pipe_chunk = atoll(ch);
    ch = getenv("CKC_PIPE_TAPER_FRONT");
    if (ch && atoi(ch) >= 0 && atoi(ch) <= 8) ctx->pipe_taper_front = atoi(ch);
    ch = getenv("CKC_PIPE_NCHUNK");
    if (ch && atoi(ch) >= 1 && atoi(ch) <= 64) ctx->pipe_nchunk = atoi(ch);
    ch = getenv("CKC_ORDERED_UPLOAD");
    if (ch) ctx->ordered_upload = atoi(ch) != 0;
    ctx->timeline_copies = getenv("CKC_TIMELINE") != nullptr;
    ch = getenv("CKC_PIPE_TAPER");
    if (ch && atoi(ch) >= 0 && atoi(ch) <= 8) ctx->pipe_taper = atoi(ch);
    const char* pf = getenv("CKC_PCS_FAN");
    if (pf) ctx->pcs_fan = atoi(pf) != 0;
    const char* gf = getenv("CKC_GD_FAN");
    if (gf) ctx->gd_fan = atoi(gf) != 0;
    const char* hl = getenv("CKC_HOST_LANES");
    if (hl && atoi(hl) >= 1 && atoi(hl) <= CKC_NUM_LANES) ctx->host_lanes = atoi(hl);
    *out = ctx;
    return CKC_OK;
}

extern "C" void ckc_destroy(ckc_ctx* ctx) {
    if (!ctx) return;
    if (!ctx->children.empty()) { for (ckc_ctx* c : ctx->children) ckc_destroy(c); delete ctx; return; }
    cudaSetDevice(ctx->device);
    dev_buf* bufs[] = { &ctx->d_bodies, &ctx->d_pairtab, &ctx->d_static, &ctx->d_nodes, &ctx->d_tris, &ctx->d_ctr, &ctx->d_pairhits, &ctx->d_pairorder, &ctx->rbs_seg[0], &ctx->rbs_seg[1],
                        &ctx->rbs_cand, &ctx->rbs_pool, &ctx->rbs_eok, &ctx->rbs_nck, &ctx->rbs_hit, &ctx->rbs_eac, &ctx->rbs_eik, &ctx->rbs_nt, &ctx->rbs_in };
    for (dev_buf* b : bufs) b->release();
    for (int l = 0; l < CKC_NUM_LANES; l++) { ctx->lanes[l].release(); if (ctx->lanes[l].ev_done) cudaEventDestroy(ctx->lanes[l].ev_done); if (ctx->lanes[l].ev_p) cudaEventDestroy(ctx->lanes[l].ev_p); if (ctx->lanes[l].ev_c) cudaEventDestroy(ctx->lanes[l].ev_c); if (ctx->lanes[l].ev_up) cudaEventDestroy(ctx->lanes[l].ev_up); if (ctx->lanes[l].ev_busy) cudaEventDestroy(ctx->lanes[l].ev_busy);
        if (ctx->lanes[l].st_hi) cudaStreamDestroy(ctx->lanes[l].st_hi); if (ctx->lanes[l].st) cudaStreamDestroy(ctx->lanes[l].st); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->st_up) cudaStreamDestroy(ctx->st_up);
    ctx->d_cidx.release(); ctx->d_cq.release(); ctx->d_ccount.release();
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    delete ctx;
}

static void drain_lanes(ckc_ctx* ctx) {
    if (!ctx) return;
    if (ctx->st_up) cudaStreamSynchronize(ctx->st_up);
    for (int l = 0; l < CKC_NUM_LANES; l++) if (ctx->lanes[l].st) cudaStreamSynchronize(ctx->lanes[l].st);
}

/* multi-device dispatch: range [off, off + cnt) of an n-unit batch goes to child g; one host thread per child */
static int multi_dispatch(ckc_ctx* ctx, int64_t n, const std::function<int(ckc_ctx*, int64_t, int64_t)>& fn) {
    const int G = (int)ctx->children.size();
    std::vector<int> rc(G, CKC_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        const int64_t lo = n * g / G, hi = n * (g + 1) / G;
        if (hi <= lo) continue;
        th.emplace_back([&, g, lo, hi] { rc[g] = fn(ctx->children[g], lo, hi - lo); });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < G; g++)
        if (rc[g] != CKC_OK) { ctx->err = "device " + std::to_string(ctx->children[g]->device) + ": " + ctx->children[g]->err; return rc[g]; }
    return CKC_OK;
}
#define MULTI(n, expr) do { if (ctx && !ctx->children.empty()) { if ((n) < 0) return fail(ctx, CKC_ERR_ARG, "bad argument"); \
    return multi_dispatch(ctx, (n), [&](ckc_ctx* c, int64_t off, int64_t cnt) -> int { (void)off; (void)cnt; return (expr); }); } } while (0)
#define OFF(p, stride) ((p) ? (p) + off * (stride) : nullptr)

extern "C" int ckc_create_multi(ckc_ctx** out, int env, const char* mesh_path, const int* devices, int ndev) {
    if (!out || !devices || ndev < 1 || ndev > 64) { g_create_error = "bad argument"; return CKC_ERR_ARG; }
    *out = nullptr;
    ckc_ctx* m = new ckc_ctx();
    m->env = env; m->device = devices[0];
    memset(&m->host_stats, 0, sizeof(m->host_stats));
    for (int g = 0; g < ndev; g++) {
        ckc_ctx* c = nullptr;
        const int rc = ckc_create(&c, env, mesh_path, devices[g]);
        if (rc != CKC_OK) { for (ckc_ctx* x : m->children) ckc_destroy(x); delete m; return rc; }
        m->children.push_back(c);
    }
    m->world.npairs = m->children[0]->world.npairs;
    *out = m;
    return CKC_OK;
}
extern "C" int ckc_num_devices(const ckc_ctx* ctx) { return !ctx ? CKC_ERR_ARG : (ctx->children.empty() ? 1 : (int)ctx->children.size()); }

extern "C" int ckc_synchronize(ckc_ctx* ctx) {
    MULTI(ctx->children.size(), ckc_synchronize(c));
    if (!ctx) return CKC_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    for (int l = 0; l < CKC_NUM_LANES; l++) CU(cudaStreamSynchronize(ctx->lanes[l].st));
    return CKC_OK;
}

extern "C" int ckc_host_alloc(void** p, uint64_t bytes) { return cudaMallocHost(p, bytes) == cudaSuccess ? CKC_OK : CKC_ERR_CUDA; }
extern "C" int ckc_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? CKC_OK : CKC_ERR_CUDA; }

extern "C" int ckc_get_stats(ckc_ctx* ctx, ckc_stats* out) {
    if (!ctx || !out) return CKC_ERR_ARG;
    if (!ctx->children.empty()) {          /* sums over the devices */
        memset(out, 0, sizeof(*out));
        for (ckc_ctx* c : ctx->children) {
            ckc_stats s; const int rc = ckc_get_stats(c, &s); if (rc) return rc;
            int64_t* a = (int64_t*)out; const int64_t* b = (const int64_t*)&s;
            for (size_t i = 0; i < sizeof(ckc_stats) / sizeof(int64_t); i++) a[i] += b[i];
        }
        return CKC_OK;
    }
    CU(cudaSetDevice(ctx->device));
    ckc_counters_dev c;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(&c, ctx->d_ctr.p, sizeof(c), cudaMemcpyDeviceToHost));
    *out = ctx->host_stats;
    out->ik_feasible = (int64_t)c.ik_feasible; out->collision_calls = (int64_t)c.collision_calls; out->collisions = (int64_t)c.collisions;
    out->bv_tests = (int64_t)c.bv_tests; out->tri_tests = (int64_t)c.tri_tests; out->queue_overflows = (int64_t)(c.queue_overflows & 0xffffffffull);
    out->gd_iterations = (int64_t)c.gd_iterations; out->is_valid_calls = ctx->host_stats.is_valid_calls + (int64_t)c.is_valid_calls;
    out->kernel_launches = ctx->launches;
    return CKC_OK;
}

extern "C" int ckc_reset_stats(ckc_ctx* ctx) {
    if (!ctx) return CKC_ERR_ARG;
    if (!ctx->children.empty()) { for (ckc_ctx* c : ctx->children) { const int rc = ckc_reset_stats(c); if (rc) return rc; } return CKC_OK; }
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemset(ctx->d_ctr.p, 0, sizeof(ckc_counters_dev)));
    memset(&ctx->host_stats, 0, sizeof(ctx->host_stats));
    ctx->launches = 0;
    return CKC_OK;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* shared pipeline pieces                                                                                             */
/* ---------------------------------------------------------------------------------------------------------------- */
static const ckc_layout AOS = { 12, 1 };
static ckc_layout soa(int64_t n) { ckc_layout l = { 1, n }; return l; }

/* K3 (link frames + collision) on the configurations listed in d_list (count on the device), scratch of lane L, stream st */
static int run_collision_stage(ckc_ctx* ctx, ckc_lane& L, int64_t m_max, const double* d_q, ckc_layout l, const int32_t* d_list, const int32_t* d_m,
                               uint8_t* d_hit, uint8_t* d_valid, uint8_t* d_pair_flags, cudaStream_t st, bool allow_hop = true) {
    if (m_max <= 0) return CKC_OK;
    CU(L.ovf.reserve((size_t)m_max * 4));
    CU(L.bigstack.reserve((size_t)CKC_HEAVY_MAX_BLOCKS * 32768 * 4));      /* heavy routine: one 32 768-entry stack per block (<= 320 blocks); diagnostic form: 128 warps */
    /* on the lane's own stream the collision stage moves to the lane's HIGH-PRIORITY stream: its blocks then win the SM slots
     * that free up while the persistent projection kernels of the other lanes are still queued */
    const bool hop = allow_hop && ctx->overlap && st == L.st && L.st_hi;      /* RBS levels run alone: no hop, four stream operations less per level */
    cudaStream_t lane_st = st;
    if (hop) { CU(cudaEventRecord(L.ev_p, st)); CU(cudaStreamWaitEvent(L.st_hi, L.ev_p, 0)); st = L.st_hi; }
    ckc_world_dev w = ctx->world;                 /* by-value launch argument: this lane's pair statistics */
    w.pair_hits = L.pair_hits(); w.pair_order = L.pair_order();
    ckc_launch_cfg cfg = { ctx->num_sms };
    /* refresh the processing order from the hit statistics now and then (a 1-block kernel on the same stream) */
    if ((L.collide_launches++ & 7) == 7) { ckc_launch_update_pair_order(w, L.pair_order(), st); ctx->launches++; }
    { stage_scope ts(ctx, st, 2);
      ckc_launch_collide(w, cfg, d_q, l, d_list, d_m, m_max, d_hit, d_valid, d_pair_flags, L.work_counter(),
                         L.ovf.as<int32_t>(), L.ovf_count(), L.bigstack.as<uint32_t>(), ctx->d_ctr.as<ckc_counters_dev>(), st); }
    ctx->launches += 1;
    CU(cudaGetLastError());
    if (hop) { CU(cudaEventRecord(L.ev_c, st)); CU(cudaStreamWaitEvent(lane_st, L.ev_c, 0)); }
    return CKC_OK;
}

/* project (+ collide) one device-resident chunk */
static int run_pcs_chunk(ckc_ctx* ctx, ckc_lane& L, int64_t n, const double* d_q, ckc_layout li, const int8_t* d_ac, const int8_t* d_ik, uint64_t seed,
                         uint64_t i0, int seeded, double* d_qout, ckc_layout lo, uint8_t* d_ok, uint8_t* d_valid, bool collide, cudaStream_t st) {
    CU(L.list.reserve((size_t)n * 4));
    CU(cudaMemsetAsync(L.list_count(), 0, 4, st));
    { stage_scope ts(ctx, st, 0);
      ckc_launch_pcs_project(ctx->kin, n, d_q, li, d_ac, d_ik, seed, i0, seeded, d_qout, lo, d_ok, d_valid, collide ? L.list.as<int32_t>() : nullptr,
                             L.list_count(), ctx->d_ctr.as<ckc_counters_dev>(), st); }
    ctx->launches += 1;
    ctx->host_stats.ik_calls += n;
    CU(cudaGetLastError());
    if (collide) return run_collision_stage(ctx, L, n, d_qout, lo, L.list.as<int32_t>(), L.list_count(), nullptr, d_valid, nullptr, st);
    return CKC_OK;
}

static int run_gd_chunk(ckc_ctx* ctx, ckc_lane& L, int64_t n, const double* d_q, ckc_layout li, uint64_t seed, uint64_t i0, int seeded, double* d_qout,
                        ckc_layout lo, uint8_t* d_ok, uint8_t* d_conv, int32_t* d_iters, uint8_t* d_valid, bool collide, cudaStream_t st) {
    CU(L.list.reserve((size_t)n * 4));
    CU(L.gdstate.reserve((size_t)n * 4));
    CU(cudaMemsetAsync(L.list_count(), 0, 4, st));
    { stage_scope ts(ctx, st, 0);
      ckc_launch_gd_project(ctx->kin, ctx->num_sms, n, d_q, li, seed, i0, seeded, d_qout, lo, d_ok, d_conv, d_iters, d_valid,
                            collide ? L.list.as<int32_t>() : nullptr, L.list_count(), L.gdstate.as<int32_t>(), L.gd_counter(), ctx->d_ctr.as<ckc_counters_dev>(), st); }
    ctx->launches += 2;
    ctx->host_stats.ik_calls += n;
    CU(cudaGetLastError());
    if (collide) return run_collision_stage(ctx, L, n, d_qout, lo, L.list.as<int32_t>(), L.list_count(), nullptr, d_valid, nullptr, st);
    return CKC_OK;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* host entry points of the projection flavours: software pipeline over the lanes                                     */
/*   chunk k runs entirely on lane k % host_lanes (H2D -> kernels -> D2H in stream order); the lanes' streams overlap */
/*   each other, so copies in both directions and kernels of neighbouring chunks run concurrently (pinned host memory). */
/* ---------------------------------------------------------------------------------------------------------------- */
struct compact_out { int64_t* n_valid; int32_t* idx; double* q_valid; int64_t cap; };

/* Small batches (a planner's single questions: 1 .. 2048 samples): latency, not bandwidth.  Inputs are packed into ONE pinned staging
 * block and uploaded with one copy, all outputs come back with one copy, one lane, one stream synchronisation, no priority-stream hop:
 * 5 runtime calls around the kernels instead of ~25 (88 -> ~35 us per single-sample ckc_pcs_validate). */
#define CKC_SMALL_MAX 2048
static int project_small(ckc_ctx* ctx, int flavour, int64_t n, const double* q, const int8_t* ac, const int8_t* ik, double* q_out,
                         uint8_t* ok, uint8_t* conv, int32_t* iters, uint8_t* valid, bool collide) {
    ckc_lane& L = ctx->lanes[0];
    cudaStream_t st = L.st;
    const size_t in_bytes = ((size_t)n * 98 + 255) & ~(size_t)255, out_bytes = (size_t)n * 103;
    if (!ctx->h_small) CU(cudaMallocHost(&ctx->h_small, ((size_t)CKC_SMALL_MAX * 98 + 256) + (size_t)CKC_SMALL_MAX * 103));
    CU(L.smallio.reserve(((size_t)CKC_SMALL_MAX * 98 + 256) + (size_t)CKC_SMALL_MAX * 103));
    char* h = (char*)ctx->h_small; char* d = (char*)L.smallio.p;
    memcpy(h, q, (size_t)n * 96);
    if (ac) memcpy(h + (size_t)n * 96, ac, (size_t)n); else memset(h + (size_t)n * 96, 0, (size_t)n);
    if (ik) memcpy(h + (size_t)n * 97, ik, (size_t)n);
    CU(cudaMemcpyAsync(d, h, (size_t)n * 98, cudaMemcpyHostToDevice, st));
    double* d_qout = (double*)(d + in_bytes); int32_t* d_it = (int32_t*)(d + in_bytes + (size_t)n * 96);
    uint8_t* d_ok = (uint8_t*)(d + in_bytes + (size_t)n * 100); uint8_t* d_valid = d_ok + n; uint8_t* d_conv = d_valid + n;
    const bool saved = ctx->overlap;
    ctx->overlap = false;                 /* no hop to the high-priority stream: nothing to overlap with */
    int rc;
    if (flavour == 0) rc = run_pcs_chunk(ctx, L, n, (const double*)d, AOS, (const int8_t*)(d + (size_t)n * 96), (const int8_t*)(d + (size_t)n * 97), 0, 0, 0, d_qout, AOS,
                                         d_ok, d_valid, collide, st);
    else { CU(L.gdstate.reserve((size_t)CKC_SMALL_MAX * 4)); rc = run_gd_chunk(ctx, L, n, (const double*)d, AOS, 0, 0, 0, d_qout, AOS, d_ok, d_conv, d_it, d_valid, collide, st); }
    ctx->overlap = saved;
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CU(cudaMemcpyAsync(h + in_bytes, d + in_bytes, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const char* o = h + in_bytes;
    if (q_out) memcpy(q_out, o, (size_t)n * 96);
    if (iters) memcpy(iters, o + (size_t)n * 96, (size_t)n * 4);
    if (ok) memcpy(ok, o + (size_t)n * 100, (size_t)n);
    if (valid) memcpy(valid, o + (size_t)n * 101, (size_t)n);
    if (conv) memcpy(conv, o + (size_t)n * 102, (size_t)n);
    return CKC_OK;
}

static int project_host(ckc_ctx* ctx, int flavour /*0 PCS, 1 GD*/, int64_t n, const double* q, const int8_t* ac, const int8_t* ik, double* q_out,
                        uint8_t* ok, uint8_t* conv, int32_t* iters, uint8_t* valid, bool collide, const compact_out* co = nullptr) {
    if (!ctx || n < 0 || n > 0x7fffffff || (n > 0 && (!q || (flavour == 0 && !ik)))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if (n > 0 && n <= CKC_SMALL_MAX && !co) return project_small(ctx, flavour, n, q, ac, ik, q_out, ok, conv, iters, valid, collide);
    const int64_t C = std::max<int64_t>(1, std::min<int64_t>(ctx->pipe_chunk, std::max<int64_t>(std::min<int64_t>(n, 1 << 16), (n + ctx->pipe_nchunk - 1) / ctx->pipe_nchunk)));
    /* compact mode: each chunk appends its valid (index, configuration) pairs to its own device region; a fixed-size prefix
     * (5 % of the chunk + 1024 entries) of every region is copied back inside the pipeline, the rare overflow afterwards */
    /* chunk schedule: uniform chunks of C samples, the LAST one halved pipe_taper times (c -> c/2, c/4, c/4 ...).  The uploads of a call
     * are serialised on the one H2D link and are what bounds it; everything that still has to run after the last byte has arrived (the
     * last chunk's kernels and read-back) is pure drain, so the last chunk is made small while the early ones stay large (fewer kernel tails) */
    std::vector<int64_t> c_off, c_len;
    for (int64_t off = 0; off < n; off += C) { c_off.push_back(off); c_len.push_back(std::min(C, n - off)); }
    for (int t = 0; t < ctx->pipe_taper && !c_len.empty() && c_len.back() >= 2 * ctx->pipe_min_chunk; t++) {
        const int64_t c = c_len.back(), h = c / 2;
        c_len.back() = h;
        c_off.push_back(c_off.back() + h); c_len.push_back(c - h);
    }
    /* the FIRST chunk halved pipe_taper_front times (c/4, c/4, c/2, c, c ...): the kernels of a call cannot start before the first chunk
     * has arrived, and where the kernels (not the link) pace the pipeline that wait is added to the call one to one */
    for (int t = 0; t < ctx->pipe_taper_front && !c_len.empty() && c_len.front() >= 2 * ctx->pipe_min_chunk; t++) {
        const int64_t c = c_len.front(), h = c / 2;
        c_len.front() = c - h;
        c_off.front() += h;
        c_off.insert(c_off.begin(), c_off.front() - h); c_len.insert(c_len.begin(), h);
        if (t + 1 < ctx->pipe_taper_front) { /* keep halving the new first piece */ }
    }
    const int64_t nchunks = (int64_t)c_len.size();
    const int64_t ccopy = C / 20 + 1024;
    int32_t* h_cnt = nullptr; int32_t* h_idx = nullptr; double* h_q = nullptr;
    if (co && n > 0) {
        CU(ctx->d_cidx.reserve((size_t)n * 4)); CU(ctx->d_cq.reserve((size_t)n * 96)); CU(ctx->d_ccount.reserve((size_t)nchunks * 4));
        const size_t need = (size_t)nchunks * (16 + (size_t)ccopy * 4 + (size_t)ccopy * 96);
        if (need > ctx->h_stage_cap) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
            CU(cudaMallocHost(&ctx->h_stage, need));
            ctx->h_stage_cap = need;
        }
        h_cnt = (int32_t*)ctx->h_stage;
        h_q = (double*)((char*)ctx->h_stage + (size_t)nchunks * 16);
        h_idx = (int32_t*)((char*)h_q + (size_t)nchunks * ccopy * 96);
        CU(cudaMemsetAsync(ctx->d_ccount.p, 0, (size_t)nchunks * 4, ctx->lanes[0].st));
        CU(cudaStreamSynchronize(ctx->lanes[0].st));
    }
    for (int k = 0; k < (int)nchunks; k++) {
        const int64_t off = c_off[k], c = c_len[k];
        ckc_lane& L = ctx->lanes[k % ctx->host_lanes];
        cudaStream_t st = L.st;
        CUS(L.qin.reserve((size_t)C * 96)); CUS(L.qout.reserve((size_t)C * 96)); CUS(L.ok.reserve((size_t)C)); CUS(L.valid.reserve((size_t)C));
        if (flavour == 0) { CUS(L.ac.reserve((size_t)C)); CUS(L.ik.reserve((size_t)C)); }
        {
            cudaStream_t up = ctx->ordered_upload ? ctx->st_up : st;
            if (ctx->ordered_upload && k >= ctx->host_lanes) CUS(cudaStreamWaitEvent(up, L.ev_done, 0));      /* the lane's previous chunk has left its buffers */
            const bool saved_t = ctx->stage_timing; ctx->stage_timing = saved_t && ctx->timeline_copies;
            { stage_scope ts(ctx, up, 1);
              CUS(cudaMemcpyAsync(L.qin.p, q + off * 12, (size_t)c * 96, cudaMemcpyHostToDevice, up));
              if (flavour == 0) {
                  if (ac) CUS(cudaMemcpyAsync(L.ac.p, ac + off, (size_t)c, cudaMemcpyHostToDevice, up));
                  CUS(cudaMemcpyAsync(L.ik.p, ik + off, (size_t)c, cudaMemcpyHostToDevice, up));
              } }
            ctx->stage_timing = saved_t;
            if (ctx->ordered_upload) { CUS(cudaEventRecord(L.ev_up, up)); CUS(cudaStreamWaitEvent(st, L.ev_up, 0)); }
        }
        int rc;
        if (flavour == 0) {
            rc = run_pcs_chunk(ctx, L, c, L.qin.as<double>(), AOS, ac ? L.ac.as<int8_t>() : nullptr, L.ik.as<int8_t>(), 0, 0, 0, L.qout.as<double>(), AOS,
                               L.ok.as<uint8_t>(), L.valid.as<uint8_t>(), collide, st);
        } else {
            CUS(L.aux.reserve((size_t)C)); CUS(L.aux2.reserve((size_t)C * 4));
            rc = run_gd_chunk(ctx, L, c, L.qin.as<double>(), AOS, 0, 0, 0, L.qout.as<double>(), AOS, L.ok.as<uint8_t>(), L.aux.as<uint8_t>(),
                              L.aux2.as<int32_t>(), L.valid.as<uint8_t>(), collide, st);
        }
        if (rc) { drain_lanes(ctx); return rc; }
        if (q_out) CUS(cudaMemcpyAsync(q_out + off * 12, L.qout.p, (size_t)c * 96, cudaMemcpyDeviceToHost, st));
        if (ok) CUS(cudaMemcpyAsync(ok + off, L.ok.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        if (conv) CUS(cudaMemcpyAsync(conv + off, L.aux.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        if (iters) CUS(cudaMemcpyAsync(iters + off, L.aux2.p, (size_t)c * 4, cudaMemcpyDeviceToHost, st));
        if (valid) CUS(cudaMemcpyAsync(valid + off, L.valid.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        if (co) {
            int32_t* dci = ctx->d_cidx.as<int32_t>() + off; double* dcq = ctx->d_cq.as<double>() + off * 12; int32_t* dcc = ctx->d_ccount.as<int32_t>() + k;
            ckc_launch_compact_valid(c, (int32_t)off, L.valid.as<uint8_t>(), L.qout.as<double>(), AOS, dci, dcq, dcc, st);
            ctx->launches++;
            CUS(cudaGetLastError());
            const int64_t cc = std::min(ccopy, c);
            CUS(cudaMemcpyAsync(h_cnt + k, dcc, 4, cudaMemcpyDeviceToHost, st));
            CUS(cudaMemcpyAsync(h_idx + (size_t)k * ccopy, dci, (size_t)cc * 4, cudaMemcpyDeviceToHost, st));
            CUS(cudaMemcpyAsync(h_q + (size_t)k * ccopy * 12, dcq, (size_t)cc * 96, cudaMemcpyDeviceToHost, st));
        }
        if (ctx->ordered_upload) CUS(cudaEventRecord(L.ev_done, st));
    }
    for (int l = 0; l < CKC_NUM_LANES; l++) CU(cudaStreamSynchronize(ctx->lanes[l].st));
    if (ctx->ordered_upload) CU(cudaStreamSynchronize(ctx->st_up));
    if (co) {
        int64_t total = 0;
        for (int64_t kk = 0; kk < nchunks; kk++) {
            const int64_t cnt = h_cnt[kk], off = c_off[kk];
            const int64_t room = std::max<int64_t>(0, std::min(cnt, co->cap - total));
            const int64_t staged = std::min(room, ccopy);
            if (staged > 0) {
                memcpy(co->idx + total, h_idx + (size_t)kk * ccopy, (size_t)staged * 4);
                memcpy(co->q_valid + total * 12, h_q + (size_t)kk * ccopy * 12, (size_t)staged * 96);
            }
            if (room > staged) {          /* more valid samples in this chunk than the staged prefix: fetch the rest directly */
                CU(cudaMemcpy(co->idx + total + staged, ctx->d_cidx.as<int32_t>() + off + staged, (size_t)(room - staged) * 4, cudaMemcpyDeviceToHost));
                CU(cudaMemcpy(co->q_valid + (total + staged) * 12, ctx->d_cq.as<double>() + (off + staged) * 12, (size_t)(room - staged) * 96, cudaMemcpyDeviceToHost));
            }
            total += cnt;
        }
        *co->n_valid = total;
        if (total > co->cap) return fail(ctx, CKC_ERR_CAPACITY, "compact output capacity too small");
    }
    return CKC_OK;
}

extern "C" int ckc_pcs_project(ckc_ctx* ctx, int64_t n, const double* q, const int8_t* ac, const int8_t* ik, double* q_out, uint8_t* ok) {
    MULTI(n, ckc_pcs_project(c, cnt, OFF(q, 12), OFF(ac, 1), OFF(ik, 1), OFF(q_out, 12), OFF(ok, 1)));
    return project_host(ctx, 0, n, q, ac, ik, q_out, ok, nullptr, nullptr, nullptr, false);
}
extern "C" int ckc_pcs_validate(ckc_ctx* ctx, int64_t n, const double* q, const int8_t* ac, const int8_t* ik, double* q_out, uint8_t* ok, uint8_t* valid) {
    MULTI(n, ckc_pcs_validate(c, cnt, OFF(q, 12), OFF(ac, 1), OFF(ik, 1), OFF(q_out, 12), OFF(ok, 1), OFF(valid, 1)));
    if (!valid) return fail(ctx, CKC_ERR_ARG, "valid must not be NULL");
    if (ctx) ctx->host_stats.is_valid_calls += n;
    return project_host(ctx, 0, n, q, ac, ik, q_out, ok, nullptr, nullptr, valid, true);
}
extern "C" int ckc_gd_project(ckc_ctx* ctx, int64_t n, const double* q, double* q_out, uint8_t* ok, uint8_t* converged, int32_t* iters) {
    MULTI(n, ckc_gd_project(c, cnt, OFF(q, 12), OFF(q_out, 12), OFF(ok, 1), OFF(converged, 1), OFF(iters, 1)));
    return project_host(ctx, 1, n, q, nullptr, nullptr, q_out, ok, converged, iters, nullptr, false);
}
/* compact results of a multi-device context: every child fills a private buffer, the parts are concatenated in device order */
static int compact_multi(ckc_ctx* ctx, int flavour, int64_t n, const double* q, const int8_t* ac, const int8_t* ik, uint8_t* valid,
                         int64_t* n_valid, int32_t* idx, double* q_valid, int64_t cap) {
    const int G = (int)ctx->children.size();
    std::vector<std::vector<int32_t>> pi(G); std::vector<std::vector<double>> pq(G); std::vector<int64_t> cntv(G, 0), offv(G, 0);
    int rc = multi_dispatch(ctx, n, [&](ckc_ctx* c, int64_t off, int64_t cnt) -> int {
        int g = 0; while (ctx->children[g] != c) g++;
        const int64_t room = std::min(cnt, cap);
        pi[g].resize((size_t)room); pq[g].resize((size_t)room * 12); offv[g] = off;
        return flavour == 0 ? ckc_pcs_validate_compact(c, cnt, q + off * 12, OFF(ac, 1), ik + off, OFF(valid, 1), &cntv[g], pi[g].data(), pq[g].data(), room)
                            : ckc_gd_validate_compact(c, cnt, q + off * 12, OFF(valid, 1), &cntv[g], pi[g].data(), pq[g].data(), room);
    });
    if (rc != CKC_OK && rc != CKC_ERR_CAPACITY) return rc;
    int64_t total = 0;
    for (int g = 0; g < G; g++) {
        const int64_t have = std::min<int64_t>(cntv[g], (int64_t)pi[g].size()), take = std::max<int64_t>(0, std::min(have, cap - total));
        for (int64_t i = 0; i < take; i++) idx[total + i] = pi[g][(size_t)i] + (int32_t)offv[g];
        if (take > 0) memcpy(q_valid + total * 12, pq[g].data(), (size_t)take * 96);
        total += cntv[g];
    }
    *n_valid = total;
    return total > cap ? fail(ctx, CKC_ERR_CAPACITY, "compact output capacity too small") : CKC_OK;
}

extern "C" int ckc_pcs_validate_compact(ckc_ctx* ctx, int64_t n, const double* q, const int8_t* ac, const int8_t* ik, uint8_t* valid,
                                        int64_t* n_valid, int32_t* idx, double* q_valid, int64_t cap) {
    if (!n_valid || !idx || !q_valid || cap < 0) return fail(ctx, CKC_ERR_ARG, "bad argument");
    if (ctx && !ctx->children.empty()) return (n < 0 || !q || !ik) ? fail(ctx, CKC_ERR_ARG, "bad argument") : compact_multi(ctx, 0, n, q, ac, ik, valid, n_valid, idx, q_valid, cap);
    if (ctx) ctx->host_stats.is_valid_calls += n;
    *n_valid = 0;
    compact_out co = { n_valid, idx, q_valid, cap };
    return project_host(ctx, 0, n, q, ac, ik, nullptr, nullptr, nullptr, nullptr, valid, true, &co);
}
extern "C" int ckc_gd_validate_compact(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* valid, int64_t* n_valid, int32_t* idx, double* q_valid, int64_t cap) {
    if (!n_valid || !idx || !q_valid || cap < 0) return fail(ctx, CKC_ERR_ARG, "bad argument");
    if (ctx && !ctx->children.empty()) return (n < 0 || !q) ? fail(ctx, CKC_ERR_ARG, "bad argument") : compact_multi(ctx, 1, n, q, nullptr, nullptr, valid, n_valid, idx, q_valid, cap);
    if (ctx) ctx->host_stats.is_valid_calls += n;
    *n_valid = 0;
    compact_out co = { n_valid, idx, q_valid, cap };
    return project_host(ctx, 1, n, q, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, valid, true, &co);
}
extern "C" int ckc_gd_validate(ckc_ctx* ctx, int64_t n, const double* q, double* q_out, uint8_t* valid) {
    MULTI(n, ckc_gd_validate(c, cnt, OFF(q, 12), OFF(q_out, 12), OFF(valid, 1)));
    if (!valid) return fail(ctx, CKC_ERR_ARG, "valid must not be NULL");
    if (ctx) ctx->host_stats.is_valid_calls += n;
    return project_host(ctx, 1, n, q, nullptr, nullptr, q_out, nullptr, nullptr, nullptr, valid, true);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* device-resident entry points: fork / join over the lanes                                                           */
/*   the batch is cut into up to CKC_FAN_LANES chunks that run on the lanes' own streams (forked from / joined to the   */
/*   caller's stream with events), so the latency-bound collision kernel of one chunk overlaps the FP64-bound projection */
/*   kernel of the next.  CKC_OVERLAP=0 (ckc_set_overlap) runs everything in order on the caller's stream instead.       */
/* ---------------------------------------------------------------------------------------------------------------- */
template <class F>
static int dev_fanout(ckc_ctx* ctx, cudaStream_t user, int64_t n, bool fan, F body) {
    if (n <= 0) return CKC_OK;
    if (!fan || !ctx->overlap || n < (1 << 17)) {
        /* the call runs in order on the caller's stream with the scratch buffers of ONE lane; consecutive calls rotate over the lanes, so
         * calls issued on different caller streams can be in flight together (one call's collision kernel then fills the tail of the next
         * call's projection kernel); a lane's previous user is awaited through the lane's event, whatever stream it ran on */
        ckc_lane& L = ctx->lanes[(ctx->dev_rot++) % CKC_FAN_LANES];
        if (L.busy_set) CU(cudaStreamWaitEvent(user, L.ev_busy, 0));
        for (int64_t off = 0; off < n; off += ctx->chunk) {
            int rc = body(L, off, std::min(ctx->chunk, n - off), user);
            if (rc) return rc;
        }
        CU(cudaEventRecord(L.ev_busy, user)); L.busy_set = true;
        return CKC_OK;
    }
    int64_t c = (n + CKC_FAN_LANES - 1) / CKC_FAN_LANES;
    c = std::max<int64_t>(c, 1 << 16);
    c = std::min<int64_t>(c, ctx->chunk);
    CU(cudaEventRecord(ctx->ev_fork, user));
    int used = 0, k = 0;
    for (int64_t off = 0; off < n; off += c, k++) {
        const int li = k % CKC_FAN_LANES;
        ckc_lane& L = ctx->lanes[li];
        if (k < CKC_FAN_LANES) { CU(cudaStreamWaitEvent(L.st, ctx->ev_fork, 0)); used = k + 1; }
        int rc = body(L, off, std::min(c, n - off), L.st);
        if (rc) return rc;
    }
    for (int li = 0; li < used; li++) {
        CU(cudaEventRecord(ctx->lanes[li].ev_done, ctx->lanes[li].st));
        CU(cudaStreamWaitEvent(user, ctx->lanes[li].ev_done, 0));
    }
    return CKC_OK;
}

extern "C" int ckc_set_overlap(ckc_ctx* ctx, int enable) {
    if (ctx && !ctx->children.empty()) return multi_dispatch(ctx, (int64_t)ctx->children.size(), [&](ckc_ctx* c, int64_t, int64_t) { return ckc_set_overlap(c, enable); });
    if (!ctx) return CKC_ERR_ARG;
    ctx->overlap = enable != 0;
    return CKC_OK;
}

extern "C" int ckc_pcs_validate_dev(ckc_ctx* ctx, int64_t n, const double* d_q, const int8_t* d_ac, const int8_t* d_ik, double* d_q_out,
                                    uint8_t* d_ok, uint8_t* d_valid, void* stream) {
    if (ctx && !ctx->children.empty()) return fail(ctx, CKC_ERR_ARG, "device-resident entry points take a single-device context");
    if (!ctx || n < 0 || !d_q || !d_ik || !d_valid) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return dev_fanout(ctx, (cudaStream_t)stream, n, ctx->pcs_fan, [&](ckc_lane& L, int64_t off, int64_t c, cudaStream_t st) -> int {
        /* the projected configurations are needed by the collision stage: the caller's buffer or an internal one */
        double* qo; ckc_layout lo;
        if (d_q_out) { qo = d_q_out + off; lo = soa(n); }
        else { CU(L.qout.reserve((size_t)c * 96)); qo = L.qout.as<double>(); lo = soa(c); }
        return run_pcs_chunk(ctx, L, c, d_q + off, soa(n), d_ac ? d_ac + off : nullptr, d_ik + off, 0, 0, 0, qo, lo, d_ok ? d_ok + off : nullptr,
                             d_valid + off, true, st);
    });
}

extern "C" int ckc_spcs_validate_seeded_dev(ckc_ctx* ctx, uint64_t seed, uint64_t i0, int64_t n, double* d_q_out, uint8_t* d_ok, uint8_t* d_valid, void* stream) {
    if (ctx && !ctx->children.empty()) return fail(ctx, CKC_ERR_ARG, "device-resident entry points take a single-device context");
    if (!ctx || n < 0 || !d_valid) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return dev_fanout(ctx, (cudaStream_t)stream, n, ctx->pcs_fan, [&](ckc_lane& L, int64_t off, int64_t c, cudaStream_t st) -> int {
        double* qo; ckc_layout l;
        if (d_q_out) { qo = d_q_out + off; l = soa(n); }
        else { CU(L.qout.reserve((size_t)c * 96)); qo = L.qout.as<double>(); l = soa(c); }
        return run_pcs_chunk(ctx, L, c, nullptr, l, nullptr, nullptr, seed, i0 + (uint64_t)off, 1, qo, l, d_ok ? d_ok + off : nullptr, d_valid + off, true, st);
    });
}

extern "C" int ckc_gd_validate_dev(ckc_ctx* ctx, int64_t n, const double* d_q, double* d_q_out, uint8_t* d_valid, void* stream) {
    if (ctx && !ctx->children.empty()) return fail(ctx, CKC_ERR_ARG, "device-resident entry points take a single-device context");
    if (!ctx || n < 0 || !d_q || !d_valid) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return dev_fanout(ctx, (cudaStream_t)stream, n, ctx->gd_fan /* off by default: the persistent Newton kernel fills the GPU; chunk tails cost more than the overlap gains (measured) */, [&](ckc_lane& L, int64_t off, int64_t c, cudaStream_t st) -> int {
        double* qo; ckc_layout lo;
        if (d_q_out) { qo = d_q_out + off; lo = soa(n); }
        else { CU(L.qout.reserve((size_t)c * 96)); qo = L.qout.as<double>(); lo = soa(c); }
        return run_gd_chunk(ctx, L, c, d_q + off, soa(n), 0, 0, 0, qo, lo, nullptr, nullptr, nullptr, d_valid + off, true, st);
    });
}

extern "C" int ckc_sgd_validate_seeded_dev(ckc_ctx* ctx, uint64_t seed, uint64_t i0, int64_t n, double* d_q_out, uint8_t* d_ok, uint8_t* d_valid, void* stream) {
    if (ctx && !ctx->children.empty()) return fail(ctx, CKC_ERR_ARG, "device-resident entry points take a single-device context");
    if (!ctx || n < 0 || !d_valid) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return dev_fanout(ctx, (cudaStream_t)stream, n, ctx->gd_fan /* off by default: the persistent Newton kernel fills the GPU; chunk tails cost more than the overlap gains (measured) */, [&](ckc_lane& L, int64_t off, int64_t c, cudaStream_t st) -> int {
        double* qo; ckc_layout l;
        if (d_q_out) { qo = d_q_out + off; l = soa(n); }
        else { CU(L.qout.reserve((size_t)c * 96)); qo = L.qout.as<double>(); l = soa(c); }
        return run_gd_chunk(ctx, L, c, nullptr, l, seed, i0 + (uint64_t)off, 1, qo, l, d_ok ? d_ok + off : nullptr, nullptr, nullptr, d_valid + off, true, st);
    });
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* collision entry points                                                                                             */
/* ---------------------------------------------------------------------------------------------------------------- */
static int collide_host(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* hit, uint8_t* flags) {
    if (!ctx || n < 0 || (n > 0 && (!q || (!hit && !flags)))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0];
    cudaStream_t st = L.st;
    const int64_t chunk = std::min<int64_t>(ctx->chunk, 1 << 18);
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t c = std::min(chunk, n - off);
        CU(L.qin.reserve((size_t)c * 96)); CU(L.ok.reserve((size_t)c));
        CU(cudaMemcpyAsync(L.qin.p, q + off * 12, (size_t)c * 96, cudaMemcpyHostToDevice, st));
        if (flags) { CU(L.flags.reserve((size_t)c * CKC_MAX_PAIRS)); CU(cudaMemsetAsync(L.flags.p, 0, (size_t)c * CKC_MAX_PAIRS, st)); }
        int rc = run_collision_stage(ctx, L, c, L.qin.as<double>(), AOS, nullptr, nullptr, L.ok.as<uint8_t>(), nullptr,
                                     flags ? L.flags.as<uint8_t>() : nullptr, st);
        if (rc) return rc;
        if (hit) CU(cudaMemcpyAsync(hit + off, L.ok.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        if (flags) CU(cudaMemcpyAsync(flags + off * CKC_MAX_PAIRS, L.flags.p, (size_t)c * CKC_MAX_PAIRS, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}
extern "C" int ckc_collide(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* hit) {
    MULTI(n, ckc_collide(c, cnt, OFF(q, 12), OFF(hit, 1))); return collide_host(ctx, n, q, hit, nullptr); }
extern "C" int ckc_collide_pairs(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* flags) {
    MULTI(n, ckc_collide_pairs(c, cnt, OFF(q, 12), OFF(flags, CKC_MAX_PAIRS))); return collide_host(ctx, n, q, nullptr, flags); }

/* debug aid (not part of include/ckc.h): box rounds | triangle steps << 16 that k_collide spent on each configuration */
extern uint32_t* g_ckc_dbg_rounds;
extern "C" int ckc_debug_collide_rounds(ckc_ctx* ctx, int64_t n, const double* q, uint32_t* rounds) {
    if (!ctx || !ctx->children.empty() || n <= 0 || n > (1 << 18) || !q || !rounds) return CKC_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    dev_buf d; CU(d.reserve((size_t)n * 4));
    std::vector<uint8_t> hit((size_t)n);
    g_ckc_dbg_rounds = d.as<uint32_t>();
    const int rc = collide_host(ctx, n, q, hit.data(), nullptr);
    g_ckc_dbg_rounds = nullptr;
    if (rc == CKC_OK) cudaMemcpy(rounds, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
    d.release();
    return rc;
}

/* ---------------------------------------------------------------------------------------------------------------- */
extern "C" int ckc_check_angle_limits(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* ok) {
    MULTI(n, ckc_check_angle_limits(c, cnt, OFF(q, 12), OFF(ok, 1)));
    if (!ctx || n < 0 || (n > 0 && (!q || !ok))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0];
    cudaStream_t st = L.st;
    for (int64_t off = 0; off < n; off += ctx->chunk) {
        const int64_t c = std::min(ctx->chunk, n - off);
        CU(L.qin.reserve((size_t)c * 96)); CU(L.ok.reserve((size_t)c));
        CU(cudaMemcpyAsync(L.qin.p, q + off * 12, (size_t)c * 96, cudaMemcpyHostToDevice, st));
        ckc_launch_limits(ctx->kin, c, L.qin.as<double>(), AOS, L.ok.as<uint8_t>(), st);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(ok + off, L.ok.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}

extern "C" int ckc_identify_ik(ckc_ctx* ctx, int64_t n, const double* q, int8_t* ik01) {
    MULTI(n, ckc_identify_ik(c, cnt, OFF(q, 12), OFF(ik01, 2)));
    if (!ctx || n < 0 || (n > 0 && (!q || !ik01))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0];
    cudaStream_t st = L.st;
    for (int64_t off = 0; off < n; off += ctx->chunk) {
        const int64_t c = std::min(ctx->chunk, n - off);
        CU(L.qin.reserve((size_t)c * 96)); CU(L.aux.reserve((size_t)c * 2));
        CU(cudaMemcpyAsync(L.qin.p, q + off * 12, (size_t)c * 96, cudaMemcpyHostToDevice, st));
        ckc_launch_identify_ik(ctx->kin, c, L.qin.as<double>(), AOS, L.aux.as<int8_t>(), st);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(ik01 + off * 2, L.aux.p, (size_t)c * 2, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}

/* isValid(const ob::State*)  StateValidityCheckerPCS.cpp:323-357 */
extern "C" int ckc_pcs_is_valid(ckc_ctx* ctx, int64_t n, const double* q, uint8_t* valid) {
    MULTI(n, ckc_pcs_is_valid(c, cnt, OFF(q, 12), OFF(valid, 1)));
    if (!ctx || n < 0 || (n > 0 && (!q || !valid))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0];
    cudaStream_t st = L.st;
    ctx->host_stats.is_valid_calls += n;
    const int64_t chunk = std::min<int64_t>(ctx->chunk, 1 << 17);          /* two collision jobs per sample */
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t c = std::min(chunk, n - off);
        CU(L.qin.reserve((size_t)c * 96)); CU(L.qout.reserve((size_t)c * 192)); CU(L.ok.reserve((size_t)c)); CU(L.valid.reserve((size_t)c));
        CU(L.aux.reserve((size_t)c * 2)); CU(L.list.reserve((size_t)c * 8));
        CU(cudaMemcpyAsync(L.qin.p, q + off * 12, (size_t)c * 96, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(L.list_count(), 0, 4, st));
        CU(cudaMemsetAsync(L.aux.p, 0, (size_t)c * 2, st));
        ckc_launch_isvalid_prepare(ctx->kin, c, L.qin.as<double>(), AOS, L.qout.as<double>(), L.ok.as<uint8_t>(), L.list.as<int32_t>(), L.list_count(), st);
        ctx->launches++; ctx->host_stats.ik_calls += 2 * c;
        CU(cudaGetLastError());
        int rc = run_collision_stage(ctx, L, 2 * c, L.qout.as<double>(), AOS, L.list.as<int32_t>(), L.list_count(), L.aux.as<uint8_t>(), nullptr, nullptr, st);
        if (rc) return rc;
        ckc_launch_isvalid_combine(c, L.ok.as<uint8_t>(), L.aux.as<uint8_t>(), L.valid.as<uint8_t>(), st);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(valid + off, L.valid.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}

/* FKsolve_rob / IKsolve_rob / kdl::FK one to one (the drop-in classes expose them; the batched entry points above fuse them) */
extern "C" int ckc_fk(ckc_ctx* ctx, int64_t n, const double* q6, const int8_t* robot, double* T16) {
    MULTI(n, ckc_fk(c, cnt, OFF(q6, 6), OFF(robot, 1), OFF(T16, 16)));
    if (!ctx || n < 0 || (n > 0 && (!q6 || !robot || !T16))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0]; cudaStream_t st = L.st;
    for (int64_t off = 0; off < n; off += ctx->chunk) {
        const int64_t c = std::min(ctx->chunk, n - off);
        CU(L.qin.reserve((size_t)c * 48)); CU(L.ik.reserve((size_t)c)); CU(L.qout.reserve((size_t)c * 128));
        CU(cudaMemcpyAsync(L.qin.p, q6 + off * 6, (size_t)c * 48, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(L.ik.p, robot + off, (size_t)c, cudaMemcpyHostToDevice, st));
        ckc_launch_fk(ctx->kin, c, L.qin.as<double>(), L.ik.as<int8_t>(), L.qout.as<double>(), st);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(T16 + off * 16, L.qout.p, (size_t)c * 128, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}
extern "C" int ckc_ik(ckc_ctx* ctx, int64_t n, const double* T16, const int8_t* robot, const int8_t* sol, double* q6, uint8_t* ok) {
    MULTI(n, ckc_ik(c, cnt, OFF(T16, 16), OFF(robot, 1), OFF(sol, 1), OFF(q6, 6), OFF(ok, 1)));
    if (!ctx || n < 0 || (n > 0 && (!T16 || !robot || !sol || !q6 || !ok))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0]; cudaStream_t st = L.st;
    for (int64_t off = 0; off < n; off += ctx->chunk) {
        const int64_t c = std::min(ctx->chunk, n - off);
        CU(L.qin.reserve((size_t)c * 128)); CU(L.ik.reserve((size_t)c)); CU(L.ac.reserve((size_t)c)); CU(L.qout.reserve((size_t)c * 48)); CU(L.ok.reserve((size_t)c));
        CU(cudaMemcpyAsync(L.qin.p, T16 + off * 16, (size_t)c * 128, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(L.ac.p, robot + off, (size_t)c, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(L.ik.p, sol + off, (size_t)c, cudaMemcpyHostToDevice, st));
        ckc_launch_ik(ctx->kin, c, L.qin.as<double>(), L.ac.as<int8_t>(), L.ik.as<int8_t>(), L.qout.as<double>(), L.ok.as<uint8_t>(), st);
        ctx->launches++; ctx->host_stats.ik_calls += c;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(q6 + off * 6, L.qout.p, (size_t)c * 48, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ok + off, L.ok.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}
extern "C" int ckc_kdl_fk(ckc_ctx* ctx, int64_t n, const double* q12, double* T16) {
    MULTI(n, ckc_kdl_fk(c, cnt, OFF(q12, 12), OFF(T16, 16)));
    if (!ctx || n < 0 || (n > 0 && (!q12 || !T16))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0]; cudaStream_t st = L.st;
    for (int64_t off = 0; off < n; off += ctx->chunk) {
        const int64_t c = std::min(ctx->chunk, n - off);
        CU(L.qin.reserve((size_t)c * 96)); CU(L.qout.reserve((size_t)c * 128));
        CU(cudaMemcpyAsync(L.qin.p, q12 + off * 12, (size_t)c * 96, cudaMemcpyHostToDevice, st));
        ckc_launch_kdl_fk(ctx->kin, c, L.qin.as<double>(), L.qout.as<double>(), st);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(T16 + off * 16, L.qout.p, (size_t)c * 128, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}

extern "C" int ckc_measure_fma_peak(ckc_ctx* ctx, int precision, double* tflops) {
    if (ctx && !ctx->children.empty()) return ckc_measure_fma_peak(ctx->children[0], precision, tflops);
    if (!ctx || !tflops || (precision != 32 && precision != 64)) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0];
    CU(L.aux.reserve(64));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t err = cudaEventCreate(&e0);
    if (err == cudaSuccess) err = cudaEventCreate(&e1);
    const int iters = precision == 64 ? 1 << 15 : 1 << 16;
    double best = 0;
    for (int rep = 0; rep < 4 && err == cudaSuccess; rep++) {
        err = cudaEventRecord(e0, ctx->stream);
        ckc_launch_fma_peak(precision, ctx->num_sms, iters, L.aux.as<float>(), ctx->stream);
        if (err == cudaSuccess) err = cudaEventRecord(e1, ctx->stream);
        if (err == cudaSuccess) err = cudaEventSynchronize(e1);
        float ms = 0;
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 8.0 * iters * 1024.0 * 2.0 * ctx->num_sms;
        if (err == cudaSuccess && rep > 0 && ms > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (err != cudaSuccess) return fail(ctx, CKC_ERR_CUDA, std::string("fma peak probe: ") + cudaGetErrorString(err));
    ctx->launches += 4;
    *tflops = best;
    return CKC_OK;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* GD / RX entry points                                                                                               */
/* ---------------------------------------------------------------------------------------------------------------- */
extern "C" int ckc_rx_validate(ckc_ctx* ctx, int64_t n, const double* q, double epsilon, uint8_t* valid, double* resid) {
    MULTI(n, ckc_rx_validate(c, cnt, OFF(q, 12), epsilon, OFF(valid, 1), OFF(resid, 1)));
    if (!ctx || n < 0 || (n > 0 && (!q || !valid))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0];
    cudaStream_t st = L.st;
    ctx->host_stats.is_valid_calls += n;
    for (int64_t off = 0; off < n; off += ctx->chunk) {
        const int64_t c = std::min(ctx->chunk, n - off);
        CU(L.qin.reserve((size_t)c * 96)); CU(L.valid.reserve((size_t)c)); CU(L.aux2.reserve((size_t)c * 8)); CU(L.list.reserve((size_t)c * 4));
        CU(cudaMemcpyAsync(L.qin.p, q + off * 12, (size_t)c * 96, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(L.list_count(), 0, 4, st));
        ckc_launch_rx_check(ctx->kin, c, L.qin.as<double>(), AOS, epsilon, nullptr, L.aux2.as<double>(), L.valid.as<uint8_t>(),
                            L.list.as<int32_t>(), L.list_count(), st);
        ctx->launches++; ctx->host_stats.ik_calls += c;
        CU(cudaGetLastError());
        int rc = run_collision_stage(ctx, L, c, L.qin.as<double>(), AOS, L.list.as<int32_t>(), L.list_count(), nullptr, L.valid.as<uint8_t>(), nullptr, st);
        if (rc) return rc;
        CU(cudaMemcpyAsync(valid + off, L.valid.p, (size_t)c, cudaMemcpyDeviceToHost, st));
        if (resid) CU(cudaMemcpyAsync(resid + off, L.aux2.p, (size_t)c * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return CKC_OK;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* RBS entry points                                                                                                   */
/* ---------------------------------------------------------------------------------------------------------------- */
/* grow a device buffer keeping its contents */
static cudaError_t reserve_keep(dev_buf& b, size_t bytes, size_t used, cudaStream_t st) {
    if (bytes <= b.cap) return cudaSuccess;
    size_t ncap = std::max(bytes, b.cap + b.cap / 2);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) return e;
    if (b.p && used) { e = cudaMemcpyAsync(np, b.p, used, cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; e = cudaStreamSynchronize(st); if (e != cudaSuccess) return e; }
    if (b.p) cudaFree(b.p);
    b.p = np; b.cap = ncap;
    return cudaSuccess;
}

struct rbs_recon { double* confs; int64_t cap; int64_t* n_confs; };       /* reconstructRBS: ordered configurations of ONE edge */

static int rbs_host(ckc_ctx* ctx, int flavour, int64_t n_edges, const double* qa, const double* qb, const int8_t* ac, const int8_t* ik,
                    uint8_t* ok, int32_t* nchecks, const rbs_recon* rec = nullptr) {
    if (!ctx || n_edges < 0 || n_edges > (1 << 28) || (n_edges > 0 && (!qa || !qb || !ok || (flavour == 0 && !ik)))) return fail(ctx, CKC_ERR_ARG, "bad argument");
    if (n_edges == 0) return CKC_OK;
    CU(cudaSetDevice(ctx->device));
    ckc_lane& L = ctx->lanes[0];
    cudaStream_t st = L.st;
    const int ne = (int)n_edges;
    dev_buf (&seg)[2] = ctx->rbs_seg;
    dev_buf &cand = ctx->rbs_cand, &pool = ctx->rbs_pool, &nck = ctx->rbs_nck, &hit = ctx->rbs_hit;
    /* thin-tail kernel (PCS flavour): frontiers of at most tail_max segments are bisected to the end inside ONE cooperative launch */
    static const int tail_max_env = [] { const char* e = getenv("CKC_RBS_TAIL"); return e ? atoi(e) : 4096; }();
    const int64_t tail_max = flavour == 0 ? tail_max_env : 0;
    int64_t seg_cap = std::max<int64_t>(std::max<int64_t>(2 * (int64_t)ne, 1024), 4 * tail_max + 1024);
    CU(seg[0].reserve((size_t)seg_cap * 16)); CU(seg[1].reserve((size_t)seg_cap * 16)); CU(cand.reserve((size_t)seg_cap * 8));
    CU(pool.reserve((size_t)(2 * (int64_t)ne + seg_cap) * 96));
    /* per-edge results live in ONE allocation, [nchecks (4 ne) | edge_ok (ne)], so that a small call reads them back with one copy */
    CU(nck.reserve((size_t)ne * 5));
    uint8_t* d_eok = nck.as<uint8_t>() + (size_t)ne * 4;
    dev_buf& nt = ctx->rbs_nt;
    if (rec) CU(nt.reserve(pool.cap / 12));
    /* inputs: [qa | qb | ac | ik] in one device block; k_rbs_init interleaves the end points: node 2e = a, node 2e+1 = b.  Small calls
     * (a planner's edges) pack them into the pinned staging block first: one upload instead of four */
    dev_buf& ein = ctx->rbs_in;
    CU(ein.reserve((size_t)ne * 194));
    const bool small_call = ne <= CKC_SMALL_MAX;
    if (small_call) {
        if (!ctx->h_small) CU(cudaMallocHost(&ctx->h_small, ((size_t)CKC_SMALL_MAX * 98 + 256) + (size_t)CKC_SMALL_MAX * 103));
        char* h = (char*)ctx->h_small;
        memcpy(h, qa, (size_t)ne * 96); memcpy(h + (size_t)ne * 96, qb, (size_t)ne * 96);
        if (ac) memcpy(h + (size_t)ne * 192, ac, ne);
        if (ik) memcpy(h + (size_t)ne * 193, ik, ne);
        CU(cudaMemcpyAsync(ein.p, h, (size_t)ne * 194, cudaMemcpyHostToDevice, st));
    } else {
        CU(cudaMemcpyAsync(ein.p, qa, (size_t)ne * 96, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync((char*)ein.p + (size_t)ne * 96, qb, (size_t)ne * 96, cudaMemcpyHostToDevice, st));
        if (ac) CU(cudaMemcpyAsync((char*)ein.p + (size_t)ne * 192, ac, ne, cudaMemcpyHostToDevice, st));
        if (ik) CU(cudaMemcpyAsync((char*)ein.p + (size_t)ne * 193, ik, ne, cudaMemcpyHostToDevice, st));
    }
    const int8_t* d_eac = (const int8_t*)((char*)ein.p + (size_t)ne * 192); const int8_t* d_eik = d_eac + ne;
    int32_t* d_cnt = L.rbs_counts();                       /* [0] candidates, [1] next segments */
    int cur = 0;
    int64_t nseg = ne, nnodes = 2 * (int64_t)ne;
    auto make = [&](int c) {
        ckc_rbs_dev r;
        r.nodes = pool.as<double>();
        r.node_t = rec ? nt.as<double>() : nullptr;
        int32_t* s0 = seg[c].as<int32_t>(); int32_t* s1 = seg[1 - c].as<int32_t>();
        const int64_t cap0 = (int64_t)(seg[c].cap / 16), cap1 = (int64_t)(seg[1 - c].cap / 16);
        r.seg_a = s0; r.seg_b = s0 + cap0; r.seg_edge = s0 + 2 * cap0; r.seg_depth = s0 + 3 * cap0;
        r.next_a = s1; r.next_b = s1 + cap1; r.next_edge = s1 + 2 * cap1; r.next_depth = s1 + 3 * cap1;
        r.cand_node = cand.as<int32_t>(); r.cand_seg = cand.as<int32_t>() + (int64_t)(cand.cap / 8);
        r.cand_count = d_cnt; r.next_count = d_cnt + 1;
        r.edge_ok = d_eok; r.nchecks = nck.as<int32_t>();
        r.edge_ac = ac ? d_eac : nullptr; r.edge_ik = d_eik;
        r.tol = 0.05; r.max_depth = 150;                    /* StateValidityCheckerPCS.h:218-220 */
        return r;
    };
    ckc_launch_rbs_init(make(cur), ne, ein.as<double>(), st);
    ctx->launches++;
    int rounds = 0;
    bool tail_ok = tail_max > 0, results_staged = false;
    while (nseg > 0) {
        if (tail_ok && nseg <= tail_max) {
            /* ---- the rest of the bisection in one cooperative launch (k_rbs_tail) ---- */
            const int64_t want_nodes = nnodes + 16 * tail_max;
            if ((size_t)want_nodes * 96 > pool.cap) CU(reserve_keep(pool, (size_t)(want_nodes + want_nodes / 2) * 96, (size_t)nnodes * 96, st));
            if ((size_t)want_nodes > hit.cap) CU(hit.reserve((size_t)(want_nodes + want_nodes / 2)));
            if (rec && (size_t)want_nodes * 8 > nt.cap) CU(reserve_keep(nt, (size_t)(want_nodes + want_nodes / 2) * 8, (size_t)nnodes * 8, st));
            CU(L.ovf.reserve((size_t)(4 * tail_max) * 4));
            CU(L.bigstack.reserve((size_t)CKC_HEAVY_MAX_BLOCKS * 32768 * 4));
            ckc_rbs_tail_state hs;
            hs.nseg = (int32_t)nseg; hs.nnodes = (int32_t)nnodes; hs.cur = cur; hs.levels = 0; hs.status = -1; hs.total_cand = 0;
            hs.max_seg = (int32_t)(2 * tail_max);
            hs.seg_cap = (int32_t)std::min<int64_t>(seg[0].cap / 16, seg[1].cap / 16);
            hs.cand_cap = (int32_t)(cand.cap / 8);
            hs.node_cap = (int32_t)std::min<int64_t>(std::min<int64_t>((int64_t)(pool.cap / 96), (int64_t)hit.cap), rec ? (int64_t)(nt.cap / 8) : (int64_t)1 << 30);
            hs.max_levels = 400;
            /* one upload zeroes the level counters (ints 2..9 of the lane's scratch block) and sets the kernel's state (ints 16..26) */
            int32_t blk[25]; memset(blk, 0, sizeof(blk));
            memcpy(blk + 14, &hs, sizeof(hs));
            void* dstate = (void*)(L.small.as<int32_t>() + 16);
            CU(cudaMemcpyAsync(L.small.as<int32_t>() + 2, blk, sizeof(blk), cudaMemcpyHostToDevice, st));
            ckc_world_dev w = ctx->world; w.pair_hits = L.pair_hits(); w.pair_order = L.pair_order();
            ckc_launch_cfg cfg = { ctx->num_sms };
            int lrc;
            { stage_scope ts(ctx, st, 3);
              lrc = ckc_launch_rbs_tail(ctx->kin, w, cfg, make(0), make(1), L.work_counter(), L.ovf.as<int32_t>(), L.ovf_count(), L.bigstack.as<uint32_t>(),
                                        ctx->d_ctr.as<ckc_counters_dev>(), dstate, hit.as<uint8_t>(), st); }
            if (lrc != 0) { cudaGetLastError(); tail_ok = false; continue; }      /* no cooperative launch on this device: host-driven levels */
            ctx->launches++;
            CU(cudaMemcpyAsync(&hs, dstate, sizeof(hs), cudaMemcpyDeviceToHost, st));
            if (small_call) { CU(cudaMemcpyAsync((char*)ctx->h_small + (size_t)ne * 194, nck.p, (size_t)ne * 5, cudaMemcpyDeviceToHost, st)); results_staged = true; }
            CU(cudaStreamSynchronize(st));
            CU(cudaGetLastError());
            ctx->host_stats.is_valid_calls += hs.total_cand;
            nnodes = hs.nnodes; nseg = hs.nseg; cur = hs.cur; rounds += hs.levels;
            if (hs.status == 0) break;
            if (hs.status == 2) {          /* a buffer would overflow: grow the node storage and go on */
                const int64_t more = nnodes + 64 * tail_max + 2 * nseg;
                CU(reserve_keep(pool, (size_t)more * 96, (size_t)nnodes * 96, st));
                CU(hit.reserve((size_t)more));
                if (rec) CU(reserve_keep(nt, (size_t)more * 8, (size_t)nnodes * 8, st));
                if (2 * nseg > hs.seg_cap || nseg > hs.cand_cap) tail_ok = false;   /* cannot happen below max_seg; be safe */
                continue;
            }
            if (hs.status == 3 && rounds > 400) return fail(ctx, CKC_ERR_CAPACITY, "RBS did not terminate");
            if (hs.status != 1) continue;
            /* status 1: the frontier grew beyond the tail's limit again -- host-driven levels until it is thin */
        }
        results_staged = false;
        /* capacity for this level: <= nseg new nodes / candidates, <= 2 nseg next segments */
        if ((int64_t)(seg[1 - cur].cap / 16) < 2 * nseg) { seg[1 - cur].release(); CU(seg[1 - cur].reserve((size_t)(2 * nseg + 1024) * 16)); }
        if ((int64_t)(cand.cap / 8) < nseg) { cand.release(); CU(cand.reserve((size_t)(nseg + 1024) * 8)); }
        /* grow geometrically: a cudaMalloc / cudaFree pair per level would dominate the deep, thin tail of the frontier */
        if ((size_t)(nnodes + nseg) * 96 > pool.cap) CU(reserve_keep(pool, (size_t)((nnodes + nseg) * 3 / 2 + 4096) * 96, (size_t)nnodes * 96, st));
        if (rec && (size_t)(nnodes + nseg) * 8 > nt.cap) CU(reserve_keep(nt, pool.cap / 12, (size_t)nnodes * 8, st));
        if ((size_t)(nnodes + nseg) > hit.cap) CU(hit.reserve((size_t)((nnodes + nseg) * 3 / 2 + 4096)));
        ckc_rbs_dev r = make(cur);
        CU(cudaMemsetAsync(d_cnt, 0, 8, st));
        { stage_scope ts(ctx, st, 3);
          ckc_launch_rbs_expand(ctx->kin, r, flavour, (int)nseg, (int)nnodes, st); }
        ctx->launches++;
        CU(cudaGetLastError());
        /* collision on the surviving midpoints (work list = candidate node ids).  Small levels keep the count on the device
         * (one host synchronisation per level); large levels read it back and run bounded sub-batches */
        int32_t h[2];
        if (nseg <= (1 << 20)) {
            int rc = run_collision_stage(ctx, L, nseg, pool.as<double>(), AOS, r.cand_node, d_cnt, hit.as<uint8_t>(), nullptr, nullptr, st, false);
            if (rc) return rc;
            { stage_scope ts(ctx, st, 3);
              ckc_launch_rbs_push(r, (int)nseg, d_cnt, hit.as<uint8_t>(), st); }
        } else {
            CU(cudaMemcpyAsync(h, d_cnt, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            const int64_t sub = 1 << 20;
            for (int64_t off = 0; off < h[0]; off += sub) {
                int rc = run_collision_stage(ctx, L, std::min<int64_t>(sub, h[0] - off), pool.as<double>(), AOS, r.cand_node + off, nullptr, hit.as<uint8_t>(), nullptr, nullptr, st, false);
                if (rc) return rc;
            }
            { stage_scope ts(ctx, st, 3);
              ckc_launch_rbs_push(r, h[0], d_cnt, hit.as<uint8_t>(), st); }
        }
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(h, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        ctx->host_stats.is_valid_calls += h[0];
        nnodes += h[0];
        nseg = h[1];
        cur = 1 - cur;
        if (++rounds > 200) return fail(ctx, CKC_ERR_CAPACITY, "RBS did not terminate");
    }
    if (results_staged) {              /* came back together with the tail kernel's state */
        const char* o = (const char*)ctx->h_small + (size_t)ne * 194;
        memcpy(ok, o + (size_t)ne * 4, ne);
        if (nchecks) memcpy(nchecks, o, (size_t)ne * 4);
    } else {
        CU(cudaMemcpyAsync(ok, d_eok, ne, cudaMemcpyDeviceToHost, st));
        if (nchecks) CU(cudaMemcpyAsync(nchecks, nck.p, (size_t)ne * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    if (rec) {
        /* every node of the bisection tree with its position t along the edge (dyadic, exact for the depths a valid edge reaches):
         * sorted by t they are the matrix reconstructRBS builds -- a, the midpoints in order, b */
        std::vector<double> hn((size_t)nnodes * 12), ht((size_t)nnodes);
        CU(cudaMemcpy(hn.data(), pool.p, (size_t)nnodes * 96, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(ht.data(), nt.p, (size_t)nnodes * 8, cudaMemcpyDeviceToHost));
        std::vector<int> order((size_t)nnodes);
        for (int64_t i = 0; i < nnodes; i++) order[(size_t)i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ht[x] < ht[y]; });
        *rec->n_confs = nnodes;
        for (int64_t i = 0; i < std::min(nnodes, rec->cap); i++) memcpy(rec->confs + i * 12, hn.data() + (size_t)order[(size_t)i] * 12, 96);
        if (nnodes > rec->cap) return fail(ctx, CKC_ERR_CAPACITY, "reconstructRBS: output capacity too small");
    }
    return CKC_OK;
}

extern "C" int ckc_pcs_check_motion_rbs(ckc_ctx* ctx, int64_t n_edges, const double* qa, const double* qb, const int8_t* ac, const int8_t* ik,
                                        uint8_t* ok, int32_t* nchecks) {
    MULTI(n_edges, ckc_pcs_check_motion_rbs(c, cnt, OFF(qa, 12), OFF(qb, 12), OFF(ac, 1), OFF(ik, 1), OFF(ok, 1), OFF(nchecks, 1)));
    return rbs_host(ctx, 0, n_edges, qa, qb, ac, ik, ok, nchecks);
}
extern "C" int ckc_gd_check_motion_rbs(ckc_ctx* ctx, int64_t n_edges, const double* qa, const double* qb, uint8_t* ok, int32_t* nchecks) {
    MULTI(n_edges, ckc_gd_check_motion_rbs(c, cnt, OFF(qa, 12), OFF(qb, 12), OFF(ok, 1), OFF(nchecks, 1)));
    return rbs_host(ctx, 1, n_edges, qa, qb, nullptr, nullptr, ok, nchecks);
}
/* reconstructRBS: the same bisection with the node positions kept; one edge per call (post-processing of a solved path) */
extern "C" int ckc_pcs_reconstruct_rbs(ckc_ctx* ctx, const double* qa, const double* qb, int active_chain, int ik_sol, double* confs, int64_t cap,
                                       int64_t* n_confs, uint8_t* ok) {
    if (!ctx || !confs || !n_confs || !ok || cap < 2) return fail(ctx, CKC_ERR_ARG, "bad argument");
    if (!ctx->children.empty()) return ckc_pcs_reconstruct_rbs(ctx->children[0], qa, qb, active_chain, ik_sol, confs, cap, n_confs, ok);
    const int8_t ac = (int8_t)active_chain, ik = (int8_t)ik_sol;
    rbs_recon rec = { confs, cap, n_confs };
    return rbs_host(ctx, 0, 1, qa, qb, &ac, &ik, ok, nullptr, &rec);
}
extern "C" int ckc_gd_reconstruct_rbs(ckc_ctx* ctx, const double* qa, const double* qb, double* confs, int64_t cap, int64_t* n_confs, uint8_t* ok) {
    if (!ctx || !confs || !n_confs || !ok || cap < 2) return fail(ctx, CKC_ERR_ARG, "bad argument");
    if (!ctx->children.empty()) return ckc_gd_reconstruct_rbs(ctx->children[0], qa, qb, confs, cap, n_confs, ok);
    rbs_recon rec = { confs, cap, n_confs };
    return rbs_host(ctx, 1, 1, qa, qb, nullptr, nullptr, ok, nullptr, &rec);
}


/* ---------------------------------------------------------------------------------------------------------------- */
/* per-stage timing                                                                                                   */
/* ---------------------------------------------------------------------------------------------------------------- */
extern "C" int ckc_set_stage_timing(ckc_ctx* ctx, int enable) {
    if (ctx && !ctx->children.empty()) return multi_dispatch(ctx, (int64_t)ctx->children.size(), [&](ckc_ctx* c, int64_t, int64_t) { return ckc_set_stage_timing(c, enable); });
    if (!ctx) return CKC_ERR_ARG;
    ctx->stage_timing = enable != 0;
    for (auto& p : ctx->pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    ctx->pending.clear();
    for (int i = 0; i < 4; i++) { ctx->stage_ms[i] = 0; ctx->stage_n[i] = 0; }
    return CKC_OK;
}

extern "C" int ckc_get_stage_times(ckc_ctx* ctx, double ms[4], int64_t launches[4]) {
    if (ctx && !ctx->children.empty()) return ckc_get_stage_times(ctx->children[0], ms, launches);
    if (!ctx || !ms || !launches) return CKC_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    for (auto& p : ctx->pending) {
        CU(cudaEventSynchronize(p.e1));
        float t = 0;
        CU(cudaEventElapsedTime(&t, p.e0, p.e1));
        ctx->stage_ms[p.stage] += t; ctx->stage_n[p.stage]++;
        cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
    }
    ctx->pending.clear();
    for (int i = 0; i < 4; i++) { ms[i] = ctx->stage_ms[i]; launches[i] = ctx->stage_n[i]; }
    return CKC_OK;
}


/* debug aid (not part of include/ckc.h): prints every pending stage interval relative to the first recorded event */
extern "C" int ckc_debug_timeline(ckc_ctx* ctx) {
    if (!ctx || ctx->pending.empty()) return CKC_ERR_ARG;
    cudaSetDevice(ctx->device);
    for (int l = 0; l < CKC_NUM_LANES; l++) cudaStreamSynchronize(ctx->lanes[l].st);
    cudaEvent_t base = ctx->pending[0].e0;
    static const char* nm[4] = { "project", "upload", "collide", "rbs" };
    for (auto& p : ctx->pending) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, base, p.e0); cudaEventElapsedTime(&b, base, p.e1);
        fprintf(stderr, "  %-8s %8.3f -> %8.3f ms  (%.3f)\n", nm[p.stage], a, b, b - a);
    }
    return CKC_OK;
}
