/*
 * ckc_kernels.cu -- the sm_100a kernels of libckc and their launchers.  Product code (no oracle includes).
 *
 *   k_pcs_project   K1  one thread per sample: FK(active arm) -> closure target -> one analytic IK branch -> joint
 *                       limits; FP64 in registers; warp-aggregated append of the IK-feasible samples to a work list.
 *   k_collide       K3  one WARP per feasible configuration: the 18 configuration-dependent link frames computed by the warp
 *                       itself into shared memory (FP64 for the exact leaf test + an FP32 copy for the box tests; the former
 *                       K2 kernel and its 1 728 B per configuration HBM round trip are gone), sphere pre-test of the 116 root
 *                       pairs, then warp-cooperative
 *                       traversal of the OBB hierarchies with a shared-memory task stack; FP32 conservative box culling
 *                       (guard band), interpenetrating pairs first and split four ways, warp-cooperative FP64 PQP-exact
 *                       triangle distance at the leaves (9 lanes per triangle pair), __ballot_sync early-out on the first
 *                       hit, adaptive per-lane pair order; configurations that stay undecided go to a deferred list and are finished
 *                       by whole blocks (one test per thread) at the end of the same launch.
 *   k_gd_project    K5  persistent grid with lane-level work stealing: Jacobian-free two-walk Newton iteration (kdl::GD), both walks
 *                       unrolled over the compile-time axis sequence, the two collinear wrist joints merged (variant 2);
 *   k_gd_finish         dense post-processing, joint limits, work-list append.
 *   k_rbs_expand / k_rbs_push  K4: one level of the recursive-bisection frontier; k_rbs_tail: the thin levels in one cooperative launch.
 *   k_isvalid_prepare / k_isvalid_combine: isValid(const ob::State*) around K3.
 *   k_fk, k_ik, k_kdl_fk: the single FK / IK members as batches.
 *   k_rx_check, k_limits, k_identify_ik, k_compact_valid, k_update_pair_order, k_fma_peak, k_transpose: helpers.
 */
#include "ckc_device.cuh"
#include <cstdlib>
#include <cstring>
#include <algorithm>

using namespace ckc;

/* -DCKC_BOUNDS: every write into a task stack / leaf-pair queue and every node / triangle index is checked on the device (the kernel
 * traps on a violation); compute-sanitizer is not available on the GPU pool, this build is its stand-in (tools/gpu/gpu_bounds.sh) */
#ifdef CKC_BOUNDS
#include <cassert>
#define CKC_CHECK(cond) assert(cond)
#else
#define CKC_CHECK(cond) ((void)0)
#endif

#ifndef CKC_STACK_CAP
#define CKC_STACK_CAP 512           /* per-warp shared-memory task stack (entries); 8 warps x (2016 + 1008 + 2048 + 256) B + 3712 B = 46.3 kB per block: two blocks stay inside the 100 kB shared-memory carve-out (the L1 keeps 156 kB for the hierarchy nodes) */
#endif
#define CKC_BIG_STACK_CAP 32768     /* global-memory task stack of the heavy routine (per block) and of the diagnostic per-pair form (per warp) */
#define CKC_TRIQ_CAP 64
#ifndef CKC_COLLIDE_WARPS
#define CKC_COLLIDE_WARPS 8
#endif

/* ================================================================================================================ */
/* K1                                                                                                                 */
/* ================================================================================================================ */
__device__ __forceinline__ void append_to_list(bool flag, int32_t value, int32_t* list, int32_t* list_count) {
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (m == 0 || list == nullptr) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(list_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (flag) list[base + __popc(m & ((1u << lane) - 1))] = value;
}

__global__ void __launch_bounds__(256)
k_pcs_project(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q_in, ckc_layout li, const int8_t* __restrict__ ac_in,
              const int8_t* __restrict__ ik_in, uint64_t key, uint64_t i0, int seeded, double* __restrict__ q_out, ckc_layout lo,
              uint8_t* __restrict__ ok_out, uint8_t* __restrict__ valid_out, int32_t* __restrict__ list, int32_t* list_count,
              ckc_counters_dev* ctr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < n) {
        double q[12];
        int ac = 0, sol;
        if (seeded) {
            /* S-PCS: uniform in the rand_q_ambient box (apc_class.cpp:643-661), ik = floor(8 u) */
#pragma unroll
            for (int j = 0; j < 12; j++) {
                const int jj = j % 6;
                const double lo_j = (jj == 5) ? -CKC_PI_ : kin.lim_lo[jj], hi_j = (jj == 5) ? CKC_PI_ : kin.lim_hi[jj];
                q[j] = lo_j + uniform01(key, i0 + (uint64_t)i, j) * (hi_j - lo_j);
            }
            sol = (int)(uniform01(key, i0 + (uint64_t)i, 12) * 8.0);
            sol = sol > 7 ? 7 : sol;
        } else {
#pragma unroll
            for (int j = 0; j < 12; j++) q[j] = q_in[i * li.stride_i + j * li.stride_j];
            ac = ac_in ? ac_in[i] : 0;
            sol = ik_in[i] & 7;
        }
        double qp[6];
        if (ac == 0) {
            ok = project_passive(kin, q, 0, sol, qp);
            if (ok) {
#pragma unroll
                for (int j = 0; j < 6; j++) q[6 + j] = qp[j];
            }
        } else {
            ok = project_passive(kin, q + 6, 1, sol, qp);
            if (ok) {
#pragma unroll
                for (int j = 0; j < 6; j++) q[j] = qp[j];
            }
        }
        if (q_out) {
#pragma unroll
            for (int j = 0; j < 12; j++) q_out[i * lo.stride_i + j * lo.stride_j] = q[j];
        }
        if (ok_out) ok_out[i] = ok ? 1 : 0;
        if (valid_out) valid_out[i] = 0;
    }
    append_to_list(ok, (int32_t)i, list, list_count);
    if (ctr) {
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&ctr->ik_feasible, (unsigned long long)__popc(m));
    }
}

void ckc_launch_pcs_project(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout li, const int8_t* ac, const int8_t* ik,
                            uint64_t seed, uint64_t i0, int seeded, double* q_out, ckc_layout lo, uint8_t* ok, uint8_t* valid,
                            int32_t* list, int32_t* list_count, ckc_counters_dev* ctr, cudaStream_t st) {
    if (n <= 0) return;
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    k_pcs_project<<<(unsigned)blocks, threads, 0, st>>>(kin, n, q, li, ac, ik, mix64(seed), i0, seeded, q_out, lo, ok, valid, list, list_count, ctr);
}

/* ================================================================================================================ */
/* K3: warp-cooperative collision                                                                                     */
/* ================================================================================================================ */
struct collide_args {
    const double* q;             /* configurations, element (sample, j) = q[sample * lq.stride_i + j * lq.stride_j] */
    ckc_layout lq;
    const int32_t* list;         /* sample index of configuration k (NULL: k) */
    const int32_t* d_m;          /* number of configurations (device), or NULL */
    int64_t m_max;
    uint8_t* hit_out;            /* hit_out[sample] = 1/0   (may be NULL) */
    uint8_t* valid_out;          /* valid_out[sample] = !hit (may be NULL) */
    uint8_t* pair_flags;         /* [sample][116] per-pair flags, disables early exit (may be NULL) */
    int32_t* work_counter;
    int32_t* ovf_list;           /* configurations whose task stack overflowed (k indices) */
    int32_t* ovf_count;
    uint32_t* big_stack;         /* overflow path: per-warp global stacks */
    ckc_counters_dev* ctr;
    int flush_n;                 /* queued triangle pairs that trigger a drain of the queue */
    float likely_gap;            /* face-axis gap below which a leaf pair is tested at once */
    int deep4;                   /* interpenetrating box pairs split both boxes at once (4 children) */
    float deep_gap;              /* face-axis gap below which a box pair counts as interpenetrating (children pushed on top) */
    int count_rounds;            /* debug: the bv / tri counters count ROUNDS (warp steps) instead of tests */
    int wide_cnt;                /* rounds with at most this many tasks split BOTH boxes of a task (4 children) */
    uint32_t* dbg_rounds;        /* debug (may be NULL): per sample box rounds | triangle steps << 16 */
    int defer_rounds, defer_tri; /* a configuration still undecided after this many box rounds / triangle steps goes to k_collide_heavy */
};

__device__ __forceinline__ ckc_pair_dev smem_pair(const int4* sp, int p) {
    const int4 a = sp[2 * p], b = sp[2 * p + 1];
    ckc_pair_dev r; r.node_a = a.x; r.node_b = a.y; r.tri_a = a.z; r.tri_b = a.w; r.frame_a = b.x; r.frame_b = b.y; r.pad0 = b.z; r.pad1 = 0;
    return r;
}
/* leaf operands with PQP_Tolerance's conventions: R = Ra^T Rb, T = Ra^T (Tb - Ta); triangle ta of mesh A as stored, triangle tb
 * of mesh B transformed into A's frame */
__device__ __forceinline__ void load_leaf_pair(const ckc_world_dev& w, const double* F64, const ckc_pair_dev pr, int ta, int tb,
                                               v3& S0, v3& S1, v3& S2, v3& T0, v3& T1, v3& T2) {
    const double* Fa = F64 + 12 * pr.frame_a;        /* FP64 frame table of this warp (shared memory), static poses included */
    const double* Fb = F64 + 12 * pr.frame_b;
    double R[9], T[3];
    const double tx = Fb[9] - Fa[9], ty = Fb[10] - Fa[10], tz = Fb[11] - Fa[11];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) R[3 * i + j] = Fa[i] * Fb[j] + Fa[3 + i] * Fb[3 + j] + Fa[6 + i] * Fb[6 + j];
        T[i] = Fa[i] * tx + Fa[3 + i] * ty + Fa[6 + i] * tz;
    }
    const double* pa = w.tris + 9 * (int64_t)(pr.tri_a + ta);
    const double* pb = w.tris + 9 * (int64_t)(pr.tri_b + tb);
    S0 = { pa[0], pa[1], pa[2] }; S1 = { pa[3], pa[4], pa[5] }; S2 = { pa[6], pa[7], pa[8] };
    v3* Tt[3] = { &T0, &T1, &T2 };
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const double x = pb[3 * v], y = pb[3 * v + 1], z = pb[3 * v + 2];
        *Tt[v] = { R[0] * x + R[1] * y + R[2] * z + T[0], R[3] * x + R[4] * y + R[5] * z + T[1], R[6] * x + R[7] * y + R[8] * z + T[2] };
    }
}

#ifndef CKC_COLLIDE_MINBLOCKS
#define CKC_COLLIDE_MINBLOCKS 2
#endif
/* root pre-test of one pair-table entry: spheres around the two root boxes (centre = box centre, radius sum + tolerance + FP32 slack
 * precomputed per entry); NaN poses stay in (the reference reports them in collision) */
__device__ __forceinline__ bool root_spheres_close(const ckc_world_dev& w, const ckc_pair_dev pr, const float* F) {
    const float4 ca = __ldg(reinterpret_cast<const float4*>(w.nodes + pr.node_a));      /* root box centre (mesh frame) */
    const float4 cb = __ldg(reinterpret_cast<const float4*>(w.nodes + pr.node_b));
    const float* Fa = F + 12 * pr.frame_a; const float* Fb = F + 12 * pr.frame_b;
    float d2 = 0;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const float xa = Fa[3 * r] * ca.x + Fa[3 * r + 1] * ca.y + Fa[3 * r + 2] * ca.z + Fa[9 + r];
        const float xb = Fb[3 * r] * cb.x + Fb[3 * r + 1] * cb.y + Fb[3 * r + 2] * cb.z + Fb[9 + r];
        d2 += (xb - xa) * (xb - xa);
    }
    const float rs = __int_as_float(pr.pad0);
    return !(d2 > rs * rs);
}

/* static shared memory of one block (8 warps): FP64 frames 8 x 2016 B, FP32 frames 8 x 1008 B, task stacks 8 x 2048 B, triangle
 * queues 8 x 256 B, pair table 3712 B = 46.3 kB (static __shared__ arrays: the compiler keeps the shared address space -- with one
 * carved-up extern buffer it fell back to generic loads / stores plus a window-base computation in every round).
 * kPerPair = the diagnostic form (ckc_collide_pairs): per-pair flags, no early exit, task stack in global memory, no deferral. */
/* shared memory of one 8-warp block of the collision kernels, declared by the KERNEL and handed to the block routines below (the RBS tail
 * kernel runs the warp routine, the heavy routine and the frontier steps in one launch and lets them share this storage) */
struct collide_smem {
    double f64[CKC_COLLIDE_WARPS][CKC_NUM_FRAMES * 12];
    float f32[CKC_COLLIDE_WARPS][CKC_NUM_FRAMES * 12];
    uint32_t stack[CKC_COLLIDE_WARPS][CKC_STACK_CAP];
    uint32_t triq[CKC_COLLIDE_WARPS][CKC_TRIQ_CAP];
    int4 pairs[CKC_MAX_PAIRS * 2];       /* the pair table, once per block (32 bytes per entry) */
};
__device__ __forceinline__ void collide_smem_init(const ckc_world_dev& w, collide_smem& sm) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < w.npairs * 2; e += blockDim.x) sm.pairs[e] = __ldg(reinterpret_cast<const int4*>(w.pair_tab) + e);
    /* the three static obstacle poses never change: once per warp, in both precisions */
    for (int e = lane; e < 36; e += 32) { const double v = w.static_frames_dev[e]; sm.f64[wib][CKC_FRAME_DOUBLES + e] = v; sm.f32[wib][CKC_FRAME_DOUBLES + e] = (float)v; }
    __syncthreads();
}

/* the warps of one block take configurations from the work counter until it runs out (collide_smem_init done by the caller) */
template <bool kPerPair>
__device__ __forceinline__ void collide_block(const ckc_world_dev& w, const collide_args& a, collide_smem& sm) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    double* F64 = sm.f64[wib];
    float* F = sm.f32[wib];
    uint32_t* triq = sm.triq[wib];
    const int4* s_pairs = sm.pairs;

    const unsigned lt_mask = (1u << lane) - 1;
    uint32_t* stack = kPerPair ? a.big_stack + (size_t)(blockIdx.x * CKC_COLLIDE_WARPS + wib) * CKC_BIG_STACK_CAP : sm.stack[wib];
    const int cap = kPerPair ? CKC_BIG_STACK_CAP : CKC_STACK_CAP;

    int64_t m = a.d_m ? (int64_t)*a.d_m : a.m_max;
    if (m > a.m_max) m = a.m_max;

    unsigned int n_bv = 0, n_tri = 0, n_cfg = 0, n_hit = 0;      /* per warp and launch: far below 2^32 */

    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(a.work_counter, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= m) break;
        const int cfg = k;                                             /* index into list[] */
        const int64_t sample = a.list ? (int64_t)a.list[cfg] : (int64_t)cfg;

        /* this configuration's 18 link frames, FP64, by the warp itself (lane j < 12 fetches joint j), then the FP32 copy the box
         * tests read.  Measured and rejected: an FP32 chain for the box tests with the FP64 frames computed only when a configuration
         * reaches its first leaf pair -- colliding configurations then pay both chains (94 k feasible: 0.97 -> 1.08 ms) and the
         * collision-free ones gain nothing (0.39 ms either way): the chain is latency, not issue slots. */
        const double qj = lane < 12 ? a.q[sample * a.lq.stride_i + lane * a.lq.stride_j] : 0.0;
        warp_link_frames(w.static_frames_dev + 36, qj, lane, F64);
        __syncwarp();
        for (int e = lane; e < CKC_FRAME_DOUBLES; e += 32) F[e] = (float)F64[e];
        /* seed the stack with the root task of every table entry that passes the sphere pre-test, in decreasing order of each
         * entry's (decayed) first-hit frequency, first entry on top: a colliding configuration usually reaches its hit inside the
         * first round and leaves through the ballot early-out.  Most of the 116 entries are far apart in most configurations: the
         * pre-test replaces ~3 of the ~4 root rounds (480 instructions per lane each) by ~35 instructions per lane and entry */
        __syncwarp();
        /* the root-box centre of every body (a (frame, mesh) combination the pair table names; <= 32) in world coordinates, one body per
         * lane, once per configuration: a table entry's sphere test is then two shared-memory reads and seven FP32 operations instead of
         * two box transforms (same arithmetic as root_spheres_close, same decisions) */
        /* storage: the last 96 words of this warp's task stack -- seeding writes at most CKC_MAX_PAIRS entries from the bottom, and only
         * after the last pre-test has been read (the ballots of the loop below order the two) */
        static_assert(CKC_MAX_PAIRS + 96 <= CKC_STACK_CAP, "body centres share the warp's task stack during seeding");
        float (*ctr)[3] = reinterpret_cast<float (*)[3]>(&sm.stack[wib][CKC_STACK_CAP - 96]);
        if (lane < w.nbodies) {
            const float4 b = __ldg(w.body_tab + lane);
            const float* Fb = F + 12 * __float_as_int(b.w);
#pragma unroll
            for (int r = 0; r < 3; r++) ctr[lane][r] = Fb[3 * r] * b.x + Fb[3 * r + 1] * b.y + Fb[3 * r + 2] * b.z + Fb[9 + r];
        }
        __syncwarp();
        int top = 0;
        {
            uint32_t task[(CKC_MAX_PAIRS + 31) / 32]; bool keep[(CKC_MAX_PAIRS + 31) / 32]; int before[(CKC_MAX_PAIRS + 31) / 32];
#pragma unroll
            for (int it = 0; it < (CKC_MAX_PAIRS + 31) / 32; it++) {
                const int k2 = it * 32 + lane;
                keep[it] = false; task[it] = 0;
                if (k2 < w.npairs) {
                    const int p = __ldg(w.pair_order + k2);
                    const int4 pb = s_pairs[2 * p + 1];                       /* frame_a, frame_b, radius sum (float bits), body_a | body_b << 8 */
                    const float* ca = ctr[pb.w & 255]; const float* cb = ctr[pb.w >> 8];
                    float d2 = 0;
#pragma unroll
                    for (int r = 0; r < 3; r++) d2 += (cb[r] - ca[r]) * (cb[r] - ca[r]);
                    const float rs = __int_as_float(pb.z);
                    keep[it] = !(d2 > rs * rs);                               /* NaN poses stay in (the reference reports them in collision) */
                    task[it] = (uint32_t)p << 22;
                }
                const unsigned mk = __ballot_sync(0xffffffffu, keep[it]);
                before[it] = top + __popc(mk & lt_mask);
                top += __popc(mk);
            }
#pragma unroll
            for (int it = 0; it < (CKC_MAX_PAIRS + 31) / 32; it++)
                if (keep[it]) { CKC_CHECK(top - 1 - before[it] >= 0 && top - 1 - before[it] < CKC_STACK_CAP - 96 && top <= cap); stack[top - 1 - before[it]] = task[it]; }      /* first entry of the order on top */
        }
        int ntri = 0;
        int rounds_box = 0, steps_tri = 0;
        bool hit = false, defer = false, hot = false, draining = false;      /* hot: a queued triangle pair looks like a hit -> test it now */
        /* env 2: "path not from above" rule, collisionDetection.cpp:427-428 */
        if (w.env == 2 && (F64[7 * 12 + 11] > 800.0 || F64[15 * 12 + 11] > 800.0)) { hit = true; top = 0; }
        __syncwarp();

        while (!hit) {
            /* a configuration that is still undecided after this much work (defaults: 128 box rounds or 24 triangle steps, ~0.1 % of
             * the IK-feasible S-PCS configurations: hundreds of triangle pairs within a few mm of the tolerance) would keep one
             * warp busy for 0.1 - 0.3 ms and set the duration of a small launch.  It goes to k_collide_heavy, where a block works on
             * it with one thread per test (~25 us of one SM per configuration, whatever its weight: lower thresholds lose, measured). */
            if (!kPerPair && (rounds_box > a.defer_rounds || steps_tri > a.defer_tri)) { defer = true; break; }
            if (ntri > 0 && (draining || ntri >= a.flush_n || top == 0 || hot)) {
                hot = false; draining = ntri > 3;
                /* ---- exact triangle tests: three pairs per step, nine lanes (one per edge pair) each ---- */
                const int g = lane / 9;
                const int take = ntri < 3 ? ntri : 3;
                const bool act = g < take;
                uint32_t t = 0;
                v3 S0 = { 0, 0, 0 }, S1 = S0, S2 = S0, T0 = S0, T1 = S0, T2 = S0;
                if (act) {
                    t = triq[ntri - 1 - g];
                    const ckc_pair_dev pr = smem_pair(s_pairs, t >> 20);
                    load_leaf_pair(w, F64, pr, (t >> 10) & 1023, t & 1023, S0, S1, S2, T0, T1, T2);
                }
                const double d = tri_dist_coop(act, lane, S0, S1, S2, T0, T1, T2);
                const bool h = act && (lane - 9 * g) == 0 && d <= w.tol;      /* PQP: closer_than_tolerance iff d <= tolerance */
                n_tri += a.count_rounds ? 1 : take;
                steps_tri++;
                ntri -= take;
                if (kPerPair) {
                    if (h) a.pair_flags[sample * CKC_MAX_PAIRS + (t >> 20)] = 1;
                } else {
                    const unsigned hm = __ballot_sync(0xffffffffu, h);
                    if (hm) {
                        hit = true;
                        if (lane == __ffs(hm) - 1) atomicAdd(w.pair_hits + (t >> 20), 1u);
                    }
                }
                __syncwarp();
                continue;
            }
            if (top == 0) break;
            /* ---- box tests, one task per lane ---- */
            const int cnt = top < 32 ? top : 32;
            const bool active = lane < cnt;
            bool expand = false, expand4 = false, leafpair = false, likely = false, deep = false;
            const bool wide = cnt <= a.wide_cnt;       /* few tasks in flight: idle lanes are free, halve the number of descent rounds */
            uint32_t c0 = 0, c1 = 0, tt = 0;
            if (active) {
                const uint32_t t = stack[top - 1 - lane];
                const int p = t >> 22, na = (t >> 11) & 2047, nb = t & 2047;
                const ckc_pair_dev pr = smem_pair(s_pairs, p);
                bool skip = false;
                if (kPerPair) skip = a.pair_flags[sample * CKC_MAX_PAIRS + p] != 0;      /* pair already decided */
                if (!skip) {
                    CKC_CHECK(p < w.npairs && pr.frame_a >= 0 && pr.frame_a < CKC_NUM_FRAMES && pr.frame_b >= 0 && pr.frame_b < CKC_NUM_FRAMES);
                    const box32 A = load_box(w.nodes + pr.node_a + na);
                    const box32 B = load_box(w.nodes + pr.node_b + nb);
                    CKC_CHECK((A.child >= 0 ? A.child + 1 : ~A.child) < 2048 && (B.child >= 0 ? B.child + 1 : ~B.child) < 2048);
                    float gap6;
                    const bool far = boxes_farther_than(A, F + 12 * pr.frame_a, B, F + 12 * pr.frame_b, w.cull_tol, gap6);
                    if (!far) {
                        if (A.child < 0 && B.child < 0) {
                            CKC_CHECK(~A.child < 1024 && ~B.child < 1024);
                            leafpair = true; tt = ((uint32_t)p << 20) | ((uint32_t)(~A.child) << 10) | (uint32_t)(~B.child);
                            likely = gap6 < a.likely_gap;             /* the two triangles' boxes are within ~2 mm along every face axis */
                        }
                        else if ((wide || (a.deep4 && gap6 < a.deep_gap)) && A.child >= 0 && B.child >= 0) {
                            expand4 = true;
                            c0 = ((uint32_t)p << 22) | ((uint32_t)A.child << 11) | (uint32_t)B.child;
                        }
                        else {
                            expand = true;
                            deep = gap6 < a.deep_gap;                 /* the boxes interpenetrate along every face axis */
                            if (B.child < 0 || (A.child >= 0 && A.size2 > B.size2)) {       /* descend the larger box */
                                c0 = ((uint32_t)p << 22) | ((uint32_t)A.child << 11) | (uint32_t)nb;
                                c1 = c0 + (1u << 11);
                            } else {
                                c0 = ((uint32_t)p << 22) | ((uint32_t)na << 11) | (uint32_t)B.child;
                                c1 = c0 + 1u;
                            }
                        }
                    }
                }
            }
            n_bv += a.count_rounds ? 1 : cnt;
            rounds_box++;
            top -= cnt;
            __syncwarp();
            if (!kPerPair && __any_sync(0xffffffffu, likely)) hot = true;
            const unsigned mt = __ballot_sync(0xffffffffu, leafpair);
            const unsigned me = __ballot_sync(0xffffffffu, expand);
            const unsigned m4 = __ballot_sync(0xffffffffu, expand4);
            if (leafpair) { CKC_CHECK(ntri + __popc(mt & lt_mask) < CKC_TRIQ_CAP); triq[ntri + __popc(mt & lt_mask)] = tt; }
            ntri += __popc(mt);
            /* children of interpenetrating boxes go on top of the others: they are popped first, which takes a colliding
             * configuration to its hit in fewer rounds (any order gives the same boolean) */
            const unsigned md = __ballot_sync(0xffffffffu, expand && deep);
            const int npush = 2 * __popc(me) + 4 * __popc(m4);
            if (top + npush > cap) { defer = true; break; }           /* stack overflow: the heavy kernel's stack lives in global memory */
            const unsigned mshallow = me & ~md;
            int pos = top + 4 * __popc(m4 & lt_mask);
            if (expand) pos = top + 4 * __popc(m4) + (deep ? 2 * __popc(mshallow) + 2 * __popc(md & lt_mask) : 2 * __popc(mshallow & lt_mask));
            CKC_CHECK(!(expand || expand4) || (pos >= 0 && pos + (expand4 ? 4 : 2) <= cap));
            if (expand) { stack[pos] = c1; stack[pos + 1] = c0; }
            if (expand4) { stack[pos] = c0 + (1u << 11) + 1u; stack[pos + 1] = c0 + (1u << 11); stack[pos + 2] = c0 + 1u; stack[pos + 3] = c0; }
            top += npush;
            __syncwarp();
        }

        if (defer) {
            if (!kPerPair) {
                if (lane == 0) {           /* ticket, entry, fence, then publish in ticket order: the published count never runs ahead of the entries */
                    const int o = atomicAdd(&a.work_counter[3], 1);
                    a.ovf_list[o] = cfg;
                    __threadfence();
                    while (atomicCAS(&a.work_counter[2], o, o + 1) != o) { }
                }
                __syncwarp();
                continue;                                             /* decided by the heavy routine at the end of some block */
            }
            hit = true;                                               /* diagnostic form, 32 768-entry stack exhausted: report conservatively */
            if (lane == 0 && a.ctr) atomicAdd(&a.ctr->queue_overflows, 1ull << 32);
        }
        if (kPerPair) {
            __syncwarp();
            unsigned any = 0;
            for (int p = lane; p < CKC_MAX_PAIRS; p += 32) any |= a.pair_flags[sample * CKC_MAX_PAIRS + p];
            hit = __any_sync(0xffffffffu, any != 0) || hit;
        }
        if (lane == 0) {
            if (a.dbg_rounds) a.dbg_rounds[sample] = (unsigned)rounds_box | ((unsigned)steps_tri << 16);
            if (a.hit_out) a.hit_out[sample] = hit ? 1 : 0;
            if (a.valid_out) a.valid_out[sample] = hit ? 0 : 1;
        }
        n_cfg++; n_hit += hit ? 1 : 0;
        __syncwarp();
    }
    if (lane == 0 && a.ctr && n_cfg) {
        atomicAdd(&a.ctr->bv_tests, (unsigned long long)n_bv);
        atomicAdd(&a.ctr->tri_tests, (unsigned long long)n_tri);
        atomicAdd(&a.ctr->collision_calls, (unsigned long long)n_cfg);
        atomicAdd(&a.ctr->collisions, (unsigned long long)n_hit);
    }
}

__device__ __forceinline__ void collide_heavy_block(const ckc_world_dev& w, const collide_args& a, collide_smem& sm);

/* One launch: every block's warps drain the work counter; a block whose warps have all run out then works through the deferred
 * (heavy) configurations published so far, as a block, and leaves.  Early blocks thereby do the heavy work in the shadow of the
 * stragglers; a configuration deferred late is seen at the latest by its own block (its warps have all finished by then).
 * Nobody waits for anybody: concurrent launches on other streams cannot dead-lock each other. */
template <bool kPerPair>
__global__ void __launch_bounds__(CKC_COLLIDE_WARPS * 32, CKC_COLLIDE_MINBLOCKS)
k_collide(const ckc_world_dev w, const collide_args a) {
    __shared__ collide_smem sm;
    collide_smem_init(w, sm);
    collide_block<kPerPair>(w, a, sm);
    if (!kPerPair) {
        __syncthreads();
        collide_heavy_block(w, a, sm);
    }
}

/* ---- K3h: the heavy configurations, one BLOCK per configuration ---------------------------------------------------
 * The configurations k_collide gave up on (work above its thresholds, or a task-stack overflow) are started again from the pair
 * table by a whole block: every round pops up to 256 tasks -- one box test per THREAD --, leaf pairs collect in a 1024-entry
 * queue that is drained with one exact FP64 TriDist per thread (tri_dist_seq: same arithmetic as the cooperative routine, 256 pairs
 * per step instead of 3), the task stack lives in global memory (32 768 entries per block).  A configuration with 450 near-tolerance
 * triangle pairs and 50 descent rounds costs one warp ~0.3 ms; here it is two triangle steps and ~25 rounds.  Same boolean:
 * the pruning test and the leaf test are the same functions, only the order of evaluation differs. */
#define CKC_HEAVY_THREADS (CKC_COLLIDE_WARPS * 32)
/* CKC_HEAVY_MAX_BLOCKS (ckc_internal.h): size of the global stack buffer (ckc_host.cu reserves CKC_HEAVY_MAX_BLOCKS x CKC_BIG_STACK_CAP entries) */
#define CKC_HEAVY_TRIQ 1024
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp_tot, int& total) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp_tot[wib] = inc;
    __syncthreads();
    int base = 0; total = 0;
#pragma unroll
    for (int i = 0; i < CKC_HEAVY_THREADS / 32; i++) { const int t = s_warp_tot[i]; if (i < wib) base += t; total += t; }
    __syncthreads();
    return base + inc - v;
}

/* storage: warp 0's frame tables, the first 1024 entries of the warps' stacks as the leaf-pair queue, the block's pair table */
__device__ __forceinline__ void collide_heavy_block(const ckc_world_dev& w, const collide_args& a, collide_smem& sm) {
    static_assert(CKC_HEAVY_TRIQ <= CKC_COLLIDE_WARPS * CKC_STACK_CAP && CKC_HEAVY_THREADS == CKC_COLLIDE_WARPS * 32, "heavy routine reuses the warp routine's shared memory");
    __shared__ int s_tot[CKC_HEAVY_THREADS / 32];
    __shared__ int s_k, s_hit;
    double* F64 = sm.f64[0];
    float* F = sm.f32[0];
    uint32_t* triq = &sm.stack[0][0];
    const int4* s_pairs = sm.pairs;
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t* stack = a.big_stack + (size_t)blockIdx.x * CKC_BIG_STACK_CAP;
    unsigned int n_bv = 0, n_tri = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            /* claim the next PUBLISHED entry of the deferred list (work_counter[1] = next, [2] = published count); never wait:
             * whatever is published later belongs to a block that is still running and will drain it itself */
            int claimed = -1;
            for (;;) {
                const int nx = *(volatile int32_t*)&a.work_counter[1], cnt = *(volatile int32_t*)&a.work_counter[2];
                if (nx >= cnt) break;
                if (atomicCAS(&a.work_counter[1], nx, nx + 1) == nx) { claimed = nx; break; }
            }
            s_k = claimed; s_hit = 0;
        }
        __syncthreads();
        const int k = s_k;
        if (k < 0) break;
        const int cfg = *(volatile int32_t*)&a.ovf_list[k];
        const int64_t sample = a.list ? (int64_t)a.list[cfg] : (int64_t)cfg;
        if (tid < 32) {                 /* warp 0: exact frames, then the FP32 copy (rounded from FP64: inside the guard band like the FP32 chain) */
            warp_link_frames(w.static_frames_dev + 36, lane < 12 ? a.q[sample * a.lq.stride_i + lane * a.lq.stride_j] : 0.0, lane, F64);
            __syncwarp();
            for (int e = lane; e < CKC_FRAME_DOUBLES; e += 32) F[e] = (float)F64[e];
        }
        __syncthreads();
        int top, ntri = 0;
        {   /* root tasks of the entries that pass the sphere pre-test */
            bool keep = false; int p = 0;
            if (tid < w.npairs) { p = tid; keep = root_spheres_close(w, smem_pair(s_pairs, p), F); }
            const int pos = block_exclusive_scan(keep ? 1 : 0, s_tot, top);
            if (keep) stack[pos] = (uint32_t)p << 22;
        }
        bool hit = w.env == 2 && (F64[7 * 12 + 11] > 800.0 || F64[15 * 12 + 11] > 800.0);
        bool overflow = false;
        __syncthreads();
        while (!hit) {
            if (ntri > 0 && (top == 0 || ntri > CKC_HEAVY_TRIQ - CKC_HEAVY_THREADS)) {
                /* ---- exact triangle tests, one pair per thread ---- */
                const int take = ntri < CKC_HEAVY_THREADS ? ntri : CKC_HEAVY_THREADS;
                bool h = false;
                if (tid < take) {
                    const uint32_t t = triq[ntri - 1 - tid];
                    v3 S0, S1, S2, T0, T1, T2;
                    load_leaf_pair(w, F64, smem_pair(s_pairs, t >> 20), (t >> 10) & 1023, t & 1023, S0, S1, S2, T0, T1, T2);
                    h = tri_dist_seq(S0, S1, S2, T0, T1, T2) <= w.tol;
                    if (h) { s_hit = 1; atomicAdd(w.pair_hits + (t >> 20), 1u); }
                }
                n_tri += tid == 0 ? take : 0;
                ntri -= take;
                __syncthreads();
                hit = s_hit != 0;
                continue;
            }
            if (top == 0) break;
            /* ---- box tests, one task per thread ---- */
            const int cnt = top < CKC_HEAVY_THREADS ? top : CKC_HEAVY_THREADS;
            int nchild = 0; bool leafpair = false;
            uint32_t c0 = 0, c1 = 0, tt = 0; bool four = false;
            if (tid < cnt) {
                const uint32_t t = stack[top - 1 - tid];
                const int p = t >> 22, na = (t >> 11) & 2047, nb = t & 2047;
                const ckc_pair_dev pr = smem_pair(s_pairs, p);
                const box32 A = load_box(w.nodes + pr.node_a + na);
                const box32 B = load_box(w.nodes + pr.node_b + nb);
                float gap6;
                if (!boxes_farther_than(A, F + 12 * pr.frame_a, B, F + 12 * pr.frame_b, w.cull_tol, gap6)) {
                    if (A.child < 0 && B.child < 0) { leafpair = true; tt = ((uint32_t)p << 20) | ((uint32_t)(~A.child) << 10) | (uint32_t)(~B.child); }
                    else if (A.child >= 0 && B.child >= 0) {          /* both boxes split at once: half the descent rounds */
                        four = true; nchild = 4;
                        c0 = ((uint32_t)p << 22) | ((uint32_t)A.child << 11) | (uint32_t)B.child;
                    } else {
                        nchild = 2;
                        if (B.child < 0) { c0 = ((uint32_t)p << 22) | ((uint32_t)A.child << 11) | (uint32_t)nb; c1 = c0 + (1u << 11); }
                        else { c0 = ((uint32_t)p << 22) | ((uint32_t)na << 11) | (uint32_t)B.child; c1 = c0 + 1u; }
                    }
                }
            }
            n_bv += tid == 0 ? cnt : 0;
            __syncthreads();                                       /* every task of this round has been read */
            top -= cnt;
            int npush, nleaf;
            const int cpos = block_exclusive_scan(nchild, s_tot, npush);
            const int lpos = block_exclusive_scan(leafpair ? 1 : 0, s_tot, nleaf);
            if (top + npush > CKC_BIG_STACK_CAP) { overflow = true; break; }
            CKC_CHECK(top + cpos >= 0 && top + cpos + nchild <= CKC_BIG_STACK_CAP && (!leafpair || ntri + lpos < CKC_HEAVY_TRIQ));
            if (nchild == 2) { stack[top + cpos] = c1; stack[top + cpos + 1] = c0; }
            if (four) { stack[top + cpos] = c0 + (1u << 11) + 1u; stack[top + cpos + 1] = c0 + (1u << 11); stack[top + cpos + 2] = c0 + 1u; stack[top + cpos + 3] = c0; }
            if (leafpair) triq[ntri + lpos] = tt;
            top += npush; ntri += nleaf;
            __syncthreads();
        }
        if (overflow) { hit = true; if (tid == 0 && a.ctr) atomicAdd(&a.ctr->queue_overflows, 1ull << 32); }      /* 32 768 tasks: report conservatively */
        if (tid == 0) {
            if (a.dbg_rounds) a.dbg_rounds[sample] = 0xffffffffu;
            if (a.hit_out) a.hit_out[sample] = hit ? 1 : 0;
            if (a.valid_out) a.valid_out[sample] = hit ? 0 : 1;
            if (a.ctr) { atomicAdd(&a.ctr->collision_calls, 1ull); if (hit) atomicAdd(&a.ctr->collisions, 1ull); atomicAdd(&a.ctr->queue_overflows, 1ull); }
        }
    }
    if (tid == 0 && a.ctr) { atomicAdd(&a.ctr->bv_tests, (unsigned long long)n_bv); atomicAdd(&a.ctr->tri_tests, (unsigned long long)n_tri); }
}

/* rank sort of the (decayed) first-hit counters -> processing order for the next launches; one tiny block */
__global__ void k_update_pair_order(unsigned int* hits, int* order, int npairs) {
    __shared__ unsigned int h[CKC_MAX_PAIRS];
    const int i = threadIdx.x;
    if (i < npairs) h[i] = hits[i];
    __syncthreads();
    if (i < npairs) {
        int rank = 0;
        for (int j = 0; j < npairs; j++) rank += (h[j] > h[i]) || (h[j] == h[i] && j < i);
        order[rank] = i;
        hits[i] = h[i] - (h[i] >> 6);          /* slow exponential decay keeps the statistics adaptive */
    }
}
void ckc_launch_update_pair_order(const ckc_world_dev& w, int* order_out, cudaStream_t st) {
    k_update_pair_order<<<1, 128, 0, st>>>(w.pair_hits, order_out, w.npairs);
}

uint32_t* g_ckc_dbg_rounds = nullptr;      /* debug hook (ckc_debug_collide_rounds): device array indexed by sample */
void ckc_launch_collide(const ckc_world_dev& w, const ckc_launch_cfg& cfg, const double* q, ckc_layout lq, const int32_t* list, const int32_t* d_m,
                        int64_t m_max, uint8_t* hit_out, uint8_t* valid_out, uint8_t* pair_flags, int32_t* work_counter,
                        int32_t* ovf_list, int32_t* ovf_count, uint32_t* big_stack, ckc_counters_dev* ctr, cudaStream_t st) {
    if (m_max <= 0) return;
    collide_args a;
    a.q = q; a.lq = lq; a.list = list; a.d_m = d_m; a.m_max = m_max; a.hit_out = hit_out; a.valid_out = valid_out; a.pair_flags = pair_flags;
    a.work_counter = work_counter; a.ovf_list = ovf_list; a.ovf_count = ovf_count; a.big_stack = big_stack; a.ctr = ctr;
    a.dbg_rounds = g_ckc_dbg_rounds;
    /* tuning / debug knobs, read once (function-local statics: initialisation is thread safe, contexts may live on several threads) */
    static const int wide_cnt = [] { const char* e = getenv("CKC_WIDE_CNT"); return e ? atoi(e) : 8; }();
    static const int count_rounds = [] { const char* e = getenv("CKC_COUNT_ROUNDS"); return e ? atoi(e) : 0; }();
    static const float deep_gap = [] { const char* e = getenv("CKC_DEEP_GAP"); return e ? (float)atof(e) : 0.0f; }();
    a.wide_cnt = wide_cnt;
    a.count_rounds = count_rounds;
    a.deep_gap = deep_gap;
    static const int deep4 = [] { const char* e = getenv("CKC_DEEP4"); return e ? atoi(e) : 1; }();
    a.deep4 = deep4;
    static const int flush_n = [] { const char* e = getenv("CKC_FLUSH_N"); return e ? atoi(e) : 6; }();
    static const float likely_gap = [] { const char* e = getenv("CKC_LIKELY_GAP"); return e ? (float)atof(e) : 2.0125f; }();
    a.flush_n = flush_n < 1 ? 1 : (flush_n > CKC_TRIQ_CAP - 32 ? CKC_TRIQ_CAP - 32 : flush_n);      /* a round queues up to 32 leaf pairs on top of flush_n - 1 */
    a.likely_gap = likely_gap;

    static const int defer_rounds = [] { const char* e = getenv("CKC_DEFER_ROUNDS"); return e ? atoi(e) : 128; }();
    static const int defer_tri = [] { const char* e = getenv("CKC_DEFER_TRI"); return e ? atoi(e) : 24; }();
    a.defer_rounds = defer_rounds; a.defer_tri = defer_tri;
    /* work_counter[0] next configuration, [1] next deferred entry, [2] deferred entries published (= *ovf_count), [3] deferred tickets */
    if (ovf_count != work_counter + 2) return;                  /* the lanes' scratch block lays them out like this */
    cudaMemsetAsync(work_counter, 0, 4 * sizeof(int32_t), st);
    if (pair_flags) {           /* diagnostic form: global-memory stacks, 16 blocks (the big-stack buffer holds 128 warps) */
        k_collide<true><<<16, CKC_COLLIDE_WARPS * 32, 0, st>>>(w, a);
        return;
    }
    /* persistent grid: exactly the resident blocks (the heavy routine's global-memory stack is indexed by the block) */
    int64_t blocks = (m_max + CKC_COLLIDE_WARPS - 1) / CKC_COLLIDE_WARPS;
    static const int blocks_per_sm = [] { const char* e = getenv("CKC_K3_BLOCKS_PER_SM"); const int v = e ? atoi(e) : CKC_COLLIDE_MINBLOCKS; return v < 1 || v > CKC_COLLIDE_MINBLOCKS ? CKC_COLLIDE_MINBLOCKS : v; }();      /* occupancy experiments */
    const int64_t resident = std::min<int64_t>((int64_t)cfg.num_sms * blocks_per_sm, CKC_HEAVY_MAX_BLOCKS);
    if (blocks > resident) blocks = resident;
    k_collide<false><<<(unsigned)blocks, CKC_COLLIDE_WARPS * 32, 0, st>>>(w, a);
}

/* ================================================================================================================ */
/* helpers                                                                                                            */
/* ================================================================================================================ */
__global__ void k_limits(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q, ckc_layout l, uint8_t* __restrict__ ok) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double qq[12];
#pragma unroll
    for (int j = 0; j < 12; j++) qq[j] = q[i * l.stride_i + j * l.stride_j];
    ok[i] = within_limits(kin, qq) ? 1 : 0;
}
void ckc_launch_limits(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, uint8_t* ok, cudaStream_t st) {
    if (n <= 0) return;
    k_limits<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kin, n, q, l, ok);
}

/* identify_state_ik (StateValidityCheckerPCS.cpp:275-316): 2 FK + up to 16 IK branches per configuration, one thread each */
__device__ __forceinline__ void identify_ik(const ckc_kin_dev& kin, const double* qq, int& ik0, int& ik1) {
    ik0 = -1; ik1 = -1;
    double R[9], p[3], Rt[9], pt[3], cand[6];
    fk_arm(kin, qq, 0, R, p);
#pragma unroll
    for (int r = 0; r < 3; r++) { Rt[3 * r] = -R[3 * r]; Rt[3 * r + 1] = -R[3 * r + 1]; Rt[3 * r + 2] = R[3 * r + 2]; pt[r] = R[3 * r] * kin.L + p[r]; }
    int nsol = 0;
#pragma unroll 1
    for (int s = 0; s < 8; s++) {
        if (!ik_arm(kin, Rt, pt, 1, s, cand)) continue;
        nsol++;
        double d2 = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) { const double d = cand[j] - qq[6 + j]; d2 += d * d; }
        if (sqrt(d2) < 1e-1) { ik0 = s; break; }
    }
    if (nsol > 0) {          /* :283-284: with no IK solution at all for the first chain the function returns early */
        fk_arm(kin, qq + 6, 1, R, p);
#pragma unroll
        for (int r = 0; r < 3; r++) { Rt[3 * r] = -R[3 * r]; Rt[3 * r + 1] = -R[3 * r + 1]; Rt[3 * r + 2] = R[3 * r + 2]; pt[r] = R[3 * r] * kin.L + p[r]; }
#pragma unroll 1
        for (int s = 0; s < 8; s++) {
            if (!ik_arm(kin, Rt, pt, 0, s, cand)) continue;
            double d2 = 0;
#pragma unroll
            for (int j = 0; j < 6; j++) { const double d = cand[j] - qq[j]; d2 += d * d; }
            if (sqrt(d2) < 1e-1) { ik1 = s; break; }
        }
    }
}

__global__ void __launch_bounds__(128)
k_identify_ik(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q, ckc_layout l, int8_t* __restrict__ ik01) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double qq[12];
#pragma unroll
    for (int j = 0; j < 12; j++) qq[j] = q[i * l.stride_i + j * l.stride_j];
    int ik0, ik1;
    identify_ik(kin, qq, ik0, ik1);
    ik01[2 * i] = (int8_t)ik0;
    ik01[2 * i + 1] = (int8_t)ik1;
}

/* isValid(const ob::State*) (StateValidityCheckerPCS.cpp:323-357), first half: identify both branches, re-project each arm
 * from the other one, apply the 0.05 / 0.12 rad agreement tests and emit the two collision jobs (q1, IK(q1)) and
 * (IK(q2), q2) of the samples that got that far.  jobs[2i], jobs[2i+1] are AoS configurations. */
__global__ void __launch_bounds__(128)
k_isvalid_prepare(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q, ckc_layout l, double* __restrict__ jobs,
                  uint8_t* __restrict__ pre, int32_t* __restrict__ list, int32_t* list_count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < n) {
        double qq[12], pa[6], pb[6];
#pragma unroll
        for (int j = 0; j < 12; j++) qq[j] = q[i * l.stride_i + j * l.stride_j];
        int ik0, ik1;
        identify_ik(kin, qq, ik0, ik1);
        /* an unidentified branch (-1) indexes the reference's branch tables out of range (:331/:342): defined as invalid here */
        if (ik0 >= 0 && ik1 >= 0 && project_passive(kin, qq, 0, ik0, pa)) {
            double d2 = 0;
#pragma unroll
            for (int j = 0; j < 6; j++) { const double d = qq[6 + j] - pa[j]; d2 += d * d; }
            if (!(sqrt(d2) > 0.5e-1) && project_passive(kin, qq + 6, 1, ik1, pb)) {
                d2 = 0;
#pragma unroll
                for (int j = 0; j < 6; j++) { const double d = qq[j] - pb[j]; d2 += d * d; }
                ok = !(sqrt(d2) > 0.12);
            }
        }
        if (ok) {
            double* o = jobs + i * 24;
#pragma unroll
            for (int j = 0; j < 6; j++) { o[j] = qq[j]; o[6 + j] = pa[j]; o[12 + j] = pb[j]; o[18 + j] = qq[6 + j]; }
        }
        pre[i] = ok ? 1 : 0;
    }
    append_to_list(ok, (int32_t)(2 * i), list, list_count);
    append_to_list(ok, (int32_t)(2 * i + 1), list, list_count);
}

__global__ void __launch_bounds__(256)
k_isvalid_combine(int64_t n, const uint8_t* __restrict__ pre, const uint8_t* __restrict__ hit, uint8_t* __restrict__ valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) valid[i] = (pre[i] && !hit[2 * i] && !hit[2 * i + 1]) ? 1 : 0;
}
void ckc_launch_isvalid_prepare(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, double* jobs, uint8_t* pre, int32_t* list,
                                int32_t* list_count, cudaStream_t st) {
    if (n <= 0) return;
    k_isvalid_prepare<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kin, n, q, l, jobs, pre, list, list_count);
}
void ckc_launch_isvalid_combine(int64_t n, const uint8_t* pre, const uint8_t* hit, uint8_t* valid, cudaStream_t st) {
    if (n <= 0) return;
    k_isvalid_combine<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, pre, hit, valid);
}
void ckc_launch_identify_ik(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, int8_t* ik01, cudaStream_t st) {
    if (n <= 0) return;
    k_identify_ik<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kin, n, q, l, ik01);
}

/* FMA-chain peak probe: 8 independent dependent-chains per thread, 2 flop per FMA */
template <typename T>
__global__ void __launch_bounds__(1024)
k_fma_peak(int iters, float* sink) {
    T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.9999, c = (T)1e-3;
    for (int i = 0; i < iters; i++) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    const T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == (T)-1) sink[0] = (float)s;
}
void ckc_launch_fma_peak(int precision, int num_sms, int iters, float* sink, cudaStream_t st) {
    if (precision == 64) k_fma_peak<double><<<num_sms * 2, 1024, 0, st>>>(iters, sink);
    else k_fma_peak<float><<<num_sms * 2, 1024, 0, st>>>(iters, sink);
}

/* valid samples of one chunk -> (index, projected configuration) pairs, warp-aggregated append; order unspecified */
__global__ void __launch_bounds__(256)
k_compact_valid(int64_t n, int32_t index_base, const uint8_t* __restrict__ valid, const double* __restrict__ q, ckc_layout l,
                int32_t* __restrict__ idx_out, double* __restrict__ q_out, int32_t* count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool v = i < n && valid[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (v) {
        const int pos = base + __popc(m & ((1u << lane) - 1));
        idx_out[pos] = index_base + (int32_t)i;
#pragma unroll
        for (int j = 0; j < 12; j++) q_out[(int64_t)pos * 12 + j] = q[i * l.stride_i + j * l.stride_j];
    }
}
void ckc_launch_compact_valid(int64_t n, int32_t index_base, const uint8_t* valid, const double* q, ckc_layout l, int32_t* idx_out, double* q_out,
                              int32_t* count, cudaStream_t st) {
    if (n <= 0) return;
    k_compact_valid<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, index_base, valid, q, l, idx_out, q_out, count);
}

/* FKsolve_rob / IKsolve_rob / kdl::FK as batch kernels (one thread per sample); robot numbers are the reference's 1 and 2 */
__global__ void __launch_bounds__(128)
k_fk(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q6, const int8_t* __restrict__ robot, double* __restrict__ T16) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q[6], R[9], p[3];
#pragma unroll
    for (int j = 0; j < 6; j++) q[j] = q6[6 * i + j];
    fk_arm(kin, q, robot[i] == 2 ? 1 : 0, R, p);
    double* T = T16 + 16 * i;
#pragma unroll
    for (int r = 0; r < 3; r++) { T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2]; T[4 * r + 3] = p[r]; }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}
__global__ void __launch_bounds__(128)
k_ik(const ckc_kin_dev kin, int64_t n, const double* __restrict__ T16, const int8_t* __restrict__ robot, const int8_t* __restrict__ sol,
     double* __restrict__ q6, uint8_t* __restrict__ ok) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* T = T16 + 16 * i;
    double Rt[9], pt[3], q[6] = { 0, 0, 0, 0, 0, 0 };
#pragma unroll
    for (int r = 0; r < 3; r++) { Rt[3 * r] = T[4 * r]; Rt[3 * r + 1] = T[4 * r + 1]; Rt[3 * r + 2] = T[4 * r + 2]; pt[r] = T[4 * r + 3]; }
    const bool v = ik_arm(kin, Rt, pt, robot[i] == 2 ? 1 : 0, sol[i] & 7, q);
    ok[i] = v ? 1 : 0;
#pragma unroll
    for (int j = 0; j < 6; j++) q6[6 * i + j] = q[j];          /* meaningful when ok[i] == 1 */
}
__global__ void __launch_bounds__(128)
k_kdl_fk(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q12, double* __restrict__ T16) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double qc[12], R[9], p[3];
    const double* q = q12 + 12 * i;
#pragma unroll
    for (int j = 0; j < 6; j++) { qc[j] = q[j]; qc[6 + j] = q[11 - j]; }      /* kdl_class.cpp:254-259 */
    qc[11] = -qc[11];
    chain_fk<false>(kin, qc, R, p, nullptr);
    double* T = T16 + 16 * i;
#pragma unroll
    for (int r = 0; r < 3; r++) { T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2]; T[4 * r + 3] = p[r]; }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}
void ckc_launch_fk(const ckc_kin_dev& kin, int64_t n, const double* q6, const int8_t* robot, double* T16, cudaStream_t st) {
    if (n > 0) k_fk<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kin, n, q6, robot, T16);
}
void ckc_launch_ik(const ckc_kin_dev& kin, int64_t n, const double* T16, const int8_t* robot, const int8_t* sol, double* q6, uint8_t* ok, cudaStream_t st) {
    if (n > 0) k_ik<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kin, n, T16, robot, sol, q6, ok);
}
void ckc_launch_kdl_fk(const ckc_kin_dev& kin, int64_t n, const double* q12, double* T16, cudaStream_t st) {
    if (n > 0) k_kdl_fk<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kin, n, q12, T16);
}

__global__ void k_transpose(const double* __restrict__ src, ckc_layout ls, double* __restrict__ dst, ckc_layout ld, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int j = 0; j < 12; j++) dst[i * ld.stride_i + j * ld.stride_j] = src[i * ls.stride_i + j * ls.stride_j];
}
void ckc_launch_transpose(const double* src, ckc_layout ls, double* dst, ckc_layout ld, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    k_transpose<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, ls, dst, ld, n);
}

/* ================================================================================================================ */
/* GD / RX flavours                                                                                                   */
/* ================================================================================================================ */
/* K5: Newton-Raphson closure projection on the 12-joint chain (kdl::GD), joint limits, append of the surviving samples
 * to the collision work list.  Persistent grid with LANE-level work stealing: a lane that has converged writes its
 * result and takes the next sample (one warp-aggregated atomicAdd per refill), so the spread of the iteration count
 * (median 9, max > 25) does not idle the other 31 lanes of its warp. */
#ifndef CKC_GD_MINBLOCKS
#define CKC_GD_MINBLOCKS 5
#endif
#ifndef CKC_GD_MINBLOCKS_U
#define CKC_GD_MINBLOCKS_U 4        /* the unrolled walks keep more values live: 128 registers */
#endif
#ifndef CKC_GD_DEFAULT_VARIANT
#define CKC_GD_DEFAULT_VARIANT 2
#endif
/* kVariant: 0 = rolled joint loops (gd_iterate), 1 = unrolled chain (gd_iterate_u), 2 = unrolled + collinear joints 5 / 6 merged */
template <bool seeded, int kVariant>
__global__ void __launch_bounds__(CKC_GD_BLOCK, kVariant == 0 ? CKC_GD_MINBLOCKS : CKC_GD_MINBLOCKS_U)
k_gd_project(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q_in, ckc_layout li, uint64_t key, uint64_t i0,
             double* __restrict__ q_raw, ckc_layout lo, int32_t* __restrict__ state_out, unsigned long long* work_counter) {
    __shared__ double sh_state[36 * CKC_GD_BLOCK];
    const int lane = threadIdx.x & 31;
    const gd_ref s = { sh_state + threadIdx.x };
    int64_t i = -1;
    int it = 0;
    bool active = false, exhausted = false;
    for (;;) {
        /* ---- refill idle lanes ---- */
        const unsigned need = __ballot_sync(0xffffffffu, !active && !exhausted);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(work_counter, (unsigned long long)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!active && !exhausted) {
                i = (int64_t)(base + __popc(need & ((1u << lane) - 1)));
                if (i < n) {
                    double q[12];
                    if (seeded) {
#pragma unroll
                        for (int j = 0; j < 12; j++) q[j] = -CKC_PI_ + uniform01(key, i0 + (uint64_t)i, j) * (2 * CKC_PI_);   /* GD.cpp:60-61 */
                    } else {
#pragma unroll
                        for (int j = 0; j < 12; j++) q[j] = q_in[i * li.stride_i + j * li.stride_j];
                    }
                    gd_init(s, q);
                    it = 0; active = true;
                } else exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        /* ---- one Newton iteration on every busy lane ---- */
        if (active) {
            const bool conv = kVariant == 0 ? gd_iterate(kin, s) : (kVariant == 1 ? gd_iterate_u<false>(kin, s) : gd_iterate_u<true>(kin, s));
            it++;
            if (conv || it >= 10000) {                                 /* ChainIkSolverPos_NR(..., 10000, 1e-5) */
                /* retire: the raw chain-order joints + (iterations | converged << 31); k_gd_finish post-processes densely */
#pragma unroll
                for (int j = 0; j < 12; j++) q_raw[i * lo.stride_i + j * lo.stride_j] = s.qc(j);
                state_out[i] = it | (conv ? (int)0x80000000 : 0);
                active = false;
            }
        }
    }
}

/* dense post-pass of kdl::GD (kdl_class.cpp:110-130): zero snap, fmod / wrap, user joint order, joint limits, outputs and
 * the warp-aggregated append of the surviving samples to the collision work list (one thread per sample, coalesced) */
__global__ void __launch_bounds__(256)
k_gd_finish(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q_in, ckc_layout li, int seeded, double* __restrict__ q_io, ckc_layout lo,
            const int32_t* __restrict__ state, uint8_t* __restrict__ ok_out, uint8_t* __restrict__ conv_out, int32_t* __restrict__ iters_out,
            uint8_t* __restrict__ valid_out, int32_t* __restrict__ list, int32_t* list_count, ckc_counters_dev* ctr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    int it = 0;
    if (i < n) {
        const int st = state[i];
        const bool conv = st < 0;
        it = st & 0x7fffffff;
        double q[12];
        if (conv) {
            double qc[12];
#pragma unroll
            for (int j = 0; j < 12; j++) qc[j] = q_io[i * lo.stride_i + j * lo.stride_j];
            gd_finish_regs(qc, q);
            ok = within_limits(kin, q);
        } else {
#pragma unroll
            for (int j = 0; j < 12; j++) q[j] = seeded ? 0.0 : q_in[i * li.stride_i + j * li.stride_j];
        }
#pragma unroll
        for (int j = 0; j < 12; j++) q_io[i * lo.stride_i + j * lo.stride_j] = q[j];
        if (ok_out) ok_out[i] = ok ? 1 : 0;
        if (conv_out) conv_out[i] = conv ? 1 : 0;
        if (iters_out) iters_out[i] = it;
        if (valid_out) valid_out[i] = 0;
    }
    append_to_list(ok, (int32_t)i, list, list_count);
    if (ctr) {
        unsigned long long sum = (unsigned long long)it;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0) { if (sum) atomicAdd(&ctr->gd_iterations, sum); if (m) atomicAdd(&ctr->ik_feasible, (unsigned long long)__popc(m)); }
    }
}

void ckc_launch_gd_project(const ckc_kin_dev& kin, int num_sms, int64_t n, const double* q, ckc_layout li, uint64_t seed, uint64_t i0, int seeded,
                           double* q_out, ckc_layout lo, uint8_t* ok, uint8_t* converged, int32_t* iters, uint8_t* valid,
                           int32_t* list, int32_t* list_count, int32_t* state_scratch, unsigned long long* work_counter, ckc_counters_dev* ctr, cudaStream_t st) {
    if (n <= 0) return;
    static const int variant = [] { const char* e = getenv("CKC_GD_VARIANT"); const int v = e ? atoi(e) : CKC_GD_DEFAULT_VARIANT; return v < 0 || v > 2 ? CKC_GD_DEFAULT_VARIANT : v; }();
    static const int blocks_per_sm = [] {
        int b = 0;
        const cudaError_t e = variant == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_gd_project<false, 0>, CKC_GD_BLOCK, 0)
                            : variant == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_gd_project<false, 1>, CKC_GD_BLOCK, 0)
                                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_gd_project<false, 2>, CKC_GD_BLOCK, 0);
        if (e != cudaSuccess || b < 1) b = 2;
        return b;
    }();
    int64_t blocks = (int64_t)num_sms * blocks_per_sm;
    /* at least ~2 samples per lane so that the lane-level work stealing can balance the iteration counts (measured on
     * 2^16..2^18-sample chunks: 1-3 equal, 6 slower by 5 %; a 1e6-sample launch is capped by the resident grid anyway) */
    static const int spl = [] { const char* e = getenv("CKC_GD_SPL"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    const int64_t needed = (n + spl * CKC_GD_BLOCK - 1) / (spl * CKC_GD_BLOCK);
    if (blocks > needed) blocks = needed;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), st);
#define CKC_GD_LAUNCH(S, V) k_gd_project<S, V><<<(unsigned)blocks, CKC_GD_BLOCK, 0, st>>>(kin, n, q, li, mix64(seed), i0, q_out, lo, state_scratch, work_counter)
    if (seeded) { if (variant == 0) CKC_GD_LAUNCH(true, 0); else if (variant == 1) CKC_GD_LAUNCH(true, 1); else CKC_GD_LAUNCH(true, 2); }
    else { if (variant == 0) CKC_GD_LAUNCH(false, 0); else if (variant == 1) CKC_GD_LAUNCH(false, 1); else CKC_GD_LAUNCH(false, 2); }
#undef CKC_GD_LAUNCH
    k_gd_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kin, n, q, li, seeded, q_out, lo, state_scratch, ok, converged, iters, valid, list, list_count, ctr);
}

/* RX: relaxed-constraint residual + joint limits; survivors go to the collision work list */
__global__ void __launch_bounds__(256)
k_rx_check(const ckc_kin_dev kin, int64_t n, const double* __restrict__ q_in, ckc_layout l, double eps, uint8_t* __restrict__ ok_out,
           double* __restrict__ resid, uint8_t* __restrict__ valid_out, int32_t* __restrict__ list, int32_t* list_count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < n) {
        double q[12];
#pragma unroll
        for (int j = 0; j < 12; j++) q[j] = q_in[i * l.stride_i + j * l.stride_j];
        const double r = rx_residual(kin, q);
        ok = (r < eps) && within_limits(kin, q);                   /* RX.cpp:209: !check_relax_constraint || !check_angle_limits || collision */
        if (resid) resid[i] = r;
        if (ok_out) ok_out[i] = ok ? 1 : 0;
        if (valid_out) valid_out[i] = 0;
    }
    append_to_list(ok, (int32_t)i, list, list_count);
}

void ckc_launch_rx_check(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, double eps, uint8_t* ok, double* resid,
                         uint8_t* valid, int32_t* list, int32_t* list_count, cudaStream_t st) {
    if (n <= 0) return;
    k_rx_check<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kin, n, q, l, eps, ok, resid, valid, list, list_count);
}

/* ================================================================================================================ */
/* K4: recursive bi-section (checkMotionRBS) as a level-synchronous frontier                                          */
/*   StateValidityCheckerPCS.cpp:486-511, StateValidityCheckerGD.cpp:214-241.  The reference recurses depth first and   */
/*   stops at the first invalid midpoint; the boolean it returns is "every node of the bisection tree is valid and no  */
/*   branch passes depth 150", which is order independent, so all segments of one level are evaluated together.        */
/* ================================================================================================================ */
/* flavour 0 = PCS (analytic projection of the passive arm), 1 = GD (Newton projection of all 12 joints).
 * One segment per thread; called with a segment index that is the same for all lanes of a warp up to + lane (ballots inside). */
template <int kFlavour>
__device__ __forceinline__ void rbs_expand_one(const ckc_kin_dev& kin, const ckc_rbs_dev& r, int s, int nseg, int node_base, double* sh_state) {
    bool cand = false;
    int my_node = -1;
    double m[12];
    if (s < nseg) {
        const int e = r.seg_edge[s];
        if (r.edge_ok[e]) {
            const double* a = r.nodes + 12 * (int64_t)r.seg_a[s];
            const double* b = r.nodes + 12 * (int64_t)r.seg_b[s];
            double sum = 0;
#pragma unroll
            for (int j = 0; j < 12; j++) { const double d = a[j] - b[j]; sum += d * d; m[j] = (a[j] + b[j]) / 2; }
            if (!(sqrt(sum) < r.tol)) {                               /* d < RBS_tol: this branch is done */
                if (r.seg_depth[s] > r.max_depth) r.edge_ok[e] = 0;   /* recursion_depth > RBS_max_depth */
                else {
                    if (r.nchecks) atomicAdd(&r.nchecks[e], 1);       /* one isValidRBS evaluation */
                    bool ok;
                    if (kFlavour == 0) {
                        double qp[6];
                        const int ac = r.edge_ac ? r.edge_ac[e] : 0, sol = r.edge_ik[e] & 7;
                        if (ac == 0) { ok = project_passive(kin, m, 0, sol, qp); if (ok) { for (int j = 0; j < 6; j++) m[6 + j] = qp[j]; } }
                        else { ok = project_passive(kin, m + 6, 1, sol, qp); if (ok) { for (int j = 0; j < 6; j++) m[j] = qp[j]; } }
                    } else {
                        int it;
                        const gd_ref gs = { sh_state + (kFlavour == 1 ? threadIdx.x : 0) };
                        ok = gd_project(kin, gs, m, 10000, &it) && within_limits(kin, m);
                    }
                    if (ok) cand = true; else r.edge_ok[e] = 0;
                }
            }
        }
    }
    /* warp-aggregated allocation of one new node + one candidate slot per surviving midpoint */
    const unsigned mk = __ballot_sync(0xffffffffu, cand);
    if (mk) {
        const int lane = threadIdx.x & 31, leader = __ffs(mk) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(r.cand_count, __popc(mk));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (cand) {
            const int c = base + __popc(mk & ((1u << lane) - 1));
            my_node = node_base + c;
            if (r.node_t) r.node_t[my_node] = (r.node_t[r.seg_a[s]] + r.node_t[r.seg_b[s]]) / 2;
            double* dst = r.nodes + 12 * (int64_t)my_node;
#pragma unroll
            for (int j = 0; j < 12; j++) dst[j] = m[j];
            r.cand_node[c] = my_node;
            r.cand_seg[c] = s;
        }
    }
}

template <int kFlavour>
__global__ void __launch_bounds__(CKC_GD_BLOCK)
k_rbs_expand(const ckc_kin_dev kin, const ckc_rbs_dev r, int nseg, int node_base) {
    __shared__ double sh_state[kFlavour == 1 ? 36 * CKC_GD_BLOCK : 1];
    rbs_expand_one<kFlavour>(kin, r, blockIdx.x * blockDim.x + threadIdx.x, nseg, node_base, sh_state);
}

/* one candidate per thread (index c the same for all lanes of a warp up to + lane) */
__device__ __forceinline__ void rbs_push_one(const ckc_rbs_dev& r, int c, int ncand, const uint8_t* __restrict__ hit) {
    bool push = false;
    int s = 0, node = 0;
    if (c < ncand) {
        s = r.cand_seg[c]; node = r.cand_node[c];
        const int e = r.seg_edge[s];
        if (hit[node]) r.edge_ok[e] = 0;
        else {
            push = true;       /* edge_ok may still be cleared by a sibling; stale segments are dropped at the next expand */
            /* stationary recursion: once the active joints of a and b coincide to the last bit (a projection discontinuity
             * between them, ~53 halvings in), the projected midpoint IS one of the end points, so one child is the segment
             * itself and the other has length 0 -- the reference keeps recursing on the identical segment until
             * recursion_depth > RBS_max_depth and returns false (PCS.cpp:495).  Same answer, without the remaining levels. */
            const long long* pm = reinterpret_cast<const long long*>(r.nodes + 12 * (int64_t)node);
            const long long* pa = reinterpret_cast<const long long*>(r.nodes + 12 * (int64_t)r.seg_a[s]);
            const long long* pb = reinterpret_cast<const long long*>(r.nodes + 12 * (int64_t)r.seg_b[s]);
            bool eqa = true, eqb = true;
#pragma unroll
            for (int j = 0; j < 12; j++) { const long long v = pm[j]; eqa = eqa && (v == pa[j]); eqb = eqb && (v == pb[j]); }
            if (eqa || eqb) {
                push = false;
                r.edge_ok[e] = 0;
                if (r.nchecks) atomicAdd(&r.nchecks[e], r.max_depth - r.seg_depth[s]);      /* the checks of the skipped levels */
            }
        }
    }
    const unsigned mk = __ballot_sync(0xffffffffu, push);
    if (mk) {
        const int lane = threadIdx.x & 31, leader = __ffs(mk) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(r.next_count, 2 * __popc(mk));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (push) {
            const int o = base + 2 * __popc(mk & ((1u << lane) - 1));
            const int e = r.seg_edge[s], d = r.seg_depth[s] + 1;
            r.next_a[o] = r.seg_a[s]; r.next_b[o] = node; r.next_edge[o] = e; r.next_depth[o] = d;
            r.next_a[o + 1] = node; r.next_b[o + 1] = r.seg_b[s]; r.next_edge[o + 1] = e; r.next_depth[o + 1] = d;
        }
    }
}

__global__ void __launch_bounds__(256)
k_rbs_push(const ckc_rbs_dev r, const int32_t* __restrict__ d_ncand, const uint8_t* __restrict__ hit) {
    rbs_push_one(r, blockIdx.x * blockDim.x + threadIdx.x, *d_ncand, hit);
}

/* ---- the thin tail of the frontier in ONE launch --------------------------------------------------------------------
 * The deep levels of a bisection hold few segments (a planner's single edge: 1, 2, 4 ... 32; the 1e5-edge sweep: ~60 of its ~70 levels
 * below a few thousand), and a host-driven level costs a stream synchronisation, an 8-byte read-back and five launches.  Once the
 * frontier is small the host hands it to this cooperative kernel (one block per SM): expand -> grid sync -> collide (warp per
 * midpoint) -> heavy configurations (block per configuration) -> push -> swap, level after level, until the frontier is empty, grows
 * beyond `max_seg` again or a buffer would overflow (the host then grows it and continues).  Same routines as the per-level kernels. */
struct rbs_tail_state {
    int32_t nseg, nnodes, cur, levels;       /* in / out */
    int32_t status;                          /* out: 0 = frontier empty, 1 = frontier above max_seg, 2 = capacity, 3 = level cap */
    int32_t total_cand;                      /* out: isValidRBS evaluations made */
    int32_t max_seg, seg_cap, cand_cap, node_cap, max_levels;
};

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(CKC_COLLIDE_WARPS * 32, 1)
k_rbs_tail(const ckc_kin_dev kin, const ckc_world_dev w, ckc_rbs_dev r0, ckc_rbs_dev r1, collide_args ca, rbs_tail_state* st, uint8_t* hit) {
    __shared__ collide_smem sm;
    cg::grid_group grid = cg::this_grid();
    collide_smem_init(w, sm);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int gwarp0 = (gtid >> 5) << 5;            /* first thread index of this warp */
    for (;;) {
        const int nseg = st->nseg, nnodes = st->nnodes, cur = st->cur, levels = st->levels;
        int status = -1;
        if (nseg == 0) status = 0;
        else if (nseg > st->max_seg) status = 1;
        else if (2 * nseg > st->seg_cap || nseg > st->cand_cap || nnodes + nseg > st->node_cap) status = 2;
        else if (levels >= st->max_levels) status = 3;
        if (status >= 0) { if (gtid == 0) st->status = status; break; }
        const ckc_rbs_dev& r = cur ? r1 : r0;
        /* ---- expand: one segment per thread ---- */
        for (int base = gwarp0; base < nseg; base += gthreads) rbs_expand_one<0>(kin, r, base + (threadIdx.x & 31), nseg, nnodes, nullptr);
        grid.sync();
        const int ncand = *r.cand_count;
        /* ---- collide the surviving midpoints (list = candidate node ids, flags indexed by node) ---- */
        collide_args a = ca;
        a.list = r.cand_node; a.d_m = r.cand_count; a.m_max = ncand; a.hit_out = hit; a.valid_out = nullptr;
        collide_block<false>(w, a, sm);
        grid.sync();
        if (*(volatile int32_t*)&a.work_counter[2] > 0) {      /* uniform across the grid: every block sees the same published count */
            collide_heavy_block(w, a, sm);
            __syncthreads();
            collide_smem_init(w, sm);               /* the heavy routine used warp 0's tables and the stacks: restore the static poses */
        }
        grid.sync();
        /* ---- push the children of the valid midpoints ---- */
        for (int base = gwarp0; base < ncand; base += gthreads) rbs_push_one(r, base + (threadIdx.x & 31), ncand, hit);
        grid.sync();
        if (gtid == 0) {
            st->nnodes = nnodes + ncand; st->nseg = *r.next_count; st->cur = cur ^ 1; st->levels = levels + 1; st->total_cand += ncand;
            *r.cand_count = 0; *r.next_count = 0;
            a.work_counter[0] = 0; a.work_counter[1] = 0; a.work_counter[2] = 0; a.work_counter[3] = 0;
        }
        grid.sync();
    }
}

/* in: the end points as uploaded, [qa (n_edges x 12) | qb (n_edges x 12)] (two contiguous H2D copies); nodes 2e / 2e + 1 of the pool */
__global__ void k_rbs_init(const ckc_rbs_dev r, int n_edges, const double* __restrict__ in) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
#pragma unroll
    for (int j = 0; j < 12; j++) { r.nodes[24 * (int64_t)e + j] = in[12 * (int64_t)e + j]; r.nodes[24 * (int64_t)e + 12 + j] = in[12 * ((int64_t)n_edges + e) + j]; }
    r.seg_a[e] = 2 * e; r.seg_b[e] = 2 * e + 1; r.seg_edge[e] = e; r.seg_depth[e] = 0;
    r.edge_ok[e] = 1;
    if (r.nchecks) r.nchecks[e] = 0;
    if (r.node_t) { r.node_t[2 * e] = 0.0; r.node_t[2 * e + 1] = 1.0; }
}

/* cooperative launch of the thin-tail kernel: one block per SM (all resident); dstate = device copy of rbs_tail_state */
int ckc_launch_rbs_tail(const ckc_kin_dev& kin, const ckc_world_dev& w, const ckc_launch_cfg& cfg, const ckc_rbs_dev& r0, const ckc_rbs_dev& r1,
                        int32_t* work_counter, int32_t* ovf_list, int32_t* ovf_count, uint32_t* big_stack, ckc_counters_dev* ctr, void* dstate, uint8_t* hit,
                        cudaStream_t st) {
    collide_args a;
    memset(&a, 0, sizeof(a));
    a.work_counter = work_counter; a.ovf_list = ovf_list; a.ovf_count = ovf_count; a.big_stack = big_stack; a.ctr = ctr;
    a.q = r0.nodes; a.lq.stride_i = 12; a.lq.stride_j = 1;
    a.wide_cnt = 8; a.deep_gap = 0.0f; a.deep4 = 1; a.flush_n = 6; a.likely_gap = 2.0125f; a.count_rounds = 0;
    a.defer_rounds = 64; a.defer_tri = 24; a.dbg_rounds = nullptr; a.pair_flags = nullptr;
    static const int blocks_per_sm = [] {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_rbs_tail, CKC_COLLIDE_WARPS * 32, 0) != cudaSuccess) b = 0;
        return b;
    }();
    if (blocks_per_sm < 1) return -1;
    int blocks = cfg.num_sms;
    if (blocks > CKC_HEAVY_MAX_BLOCKS) blocks = CKC_HEAVY_MAX_BLOCKS;
    ckc_kin_dev kin_c = kin; ckc_world_dev w_c = w; ckc_rbs_dev r0_c = r0, r1_c = r1;
    rbs_tail_state* ds = (rbs_tail_state*)dstate;
    void* args[] = { &kin_c, &w_c, &r0_c, &r1_c, &a, &ds, &hit };
    return cudaLaunchCooperativeKernel((const void*)k_rbs_tail, dim3(blocks), dim3(CKC_COLLIDE_WARPS * 32), args, 0, st) == cudaSuccess ? 0 : -1;
}

void ckc_launch_rbs_init(const ckc_rbs_dev& r, int n_edges, const double* in, cudaStream_t st) {
    if (n_edges <= 0) return;
    k_rbs_init<<<(n_edges + 255) / 256, 256, 0, st>>>(r, n_edges, in);
}
void ckc_launch_rbs_expand(const ckc_kin_dev& kin, const ckc_rbs_dev& r, int flavour, int nseg, int node_base, cudaStream_t st) {
    if (nseg <= 0) return;
    if (flavour == 0) k_rbs_expand<0><<<(nseg + CKC_GD_BLOCK - 1) / CKC_GD_BLOCK, CKC_GD_BLOCK, 0, st>>>(kin, r, nseg, node_base);
    else k_rbs_expand<1><<<(nseg + CKC_GD_BLOCK - 1) / CKC_GD_BLOCK, CKC_GD_BLOCK, 0, st>>>(kin, r, nseg, node_base);
}
void ckc_launch_rbs_push(const ckc_rbs_dev& r, int max_cand, const int32_t* d_ncand, const uint8_t* hit, cudaStream_t st) {
    if (max_cand <= 0) return;
    k_rbs_push<<<(max_cand + 255) / 256, 256, 0, st>>>(r, d_ncand, hit);
}
