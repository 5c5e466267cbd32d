/*
 * ckc_internal.h -- data layout shared by the host side (ckc_host.cu) and the kernels (ckc_kernels.cu) of libckc.
 * Product code: must not include anything of the CPU checker (the oracle directory).
 */
#ifndef CKC_INTERNAL_H_
#define CKC_INTERNAL_H_

#include <stdint.h>
#include <cuda_runtime.h>

#define CKC_NUM_MESHES 12       /* base, link1..link6, EE, table, obs (pole), cone, rod */
#define CKC_NUM_FRAMES 21       /* 8 per robot + 2 quirk frames + 3 static obstacle poses */
#define CKC_NUM_DYN_FRAMES 18   /* frames that depend on the configuration */
#define CKC_MAX_PAIRS 116
#define CKC_FRAME_DOUBLES (CKC_NUM_DYN_FRAMES * 12)
#ifndef CKC_HEAVY_MAX_BLOCKS
#define CKC_HEAVY_MAX_BLOCKS 320    /* most blocks a collision launch may have: each owns a 32 768-entry global task stack for the heavy routine */
#endif

enum { CKC_M_BASE = 0, CKC_M_LINK1, CKC_M_LINK2, CKC_M_LINK3, CKC_M_LINK4, CKC_M_LINK5, CKC_M_LINK6, CKC_M_EE,
       CKC_M_TABLE, CKC_M_OBS, CKC_M_CONE, CKC_M_ROD };

/* k_collide packs a task as pair << 22 | nodeA << 11 | nodeB and a leaf pair as pair << 20 | triA << 10 | triB */
#define CKC_MAX_MESH_NODES 2048
#define CKC_MAX_MESH_TRIS 1024

/* One node of a per-mesh oriented-bounding-box hierarchy, 64 bytes (two 32-byte sectors, four float4 loads).
 * Box = { c + sum_i s_i h_i a_i, |s_i| <= 1 }, a2 = a0 x a1, all in the mesh's own frame.
 * child >= 0: inner node, children are `child` and `child + 1` (indices relative to the mesh's first node);
 * child <  0: leaf holding triangle ~child of the mesh. */
struct __align__(16) ckc_node {
    float c[3];  float size2;       /* size2 = |h|^2: the "descend the larger box" rule */
    float h[3];  int child;
    float a0[3]; float pad0;
    float a1[3]; float pad1;
};

struct ckc_pair { uint8_t mesh_a, frame_a, mesh_b, frame_b; };
/* device copy of one pair-table entry, 32 bytes, read with two 16-byte loads by the lane that owns the task */
struct __align__(16) ckc_pair_dev { int node_a, node_b, tri_a, tri_b, frame_a, frame_b, pad0, pad1; };

/* Everything the collision kernels need; passed BY VALUE as a kernel parameter (constant bank, broadcast reads). */
struct ckc_world_dev {
    const ckc_node* nodes;              /* all meshes, concatenated */
    const ckc_pair_dev* pair_tab;       /* [npairs] in global memory: indexed per lane (a kernel-parameter array would be copied to local memory) */
    const double* static_frames_dev;    /* [5][12] in global memory: the 3 obstacle poses (same values as static_frames), then the two robot base frames */
    unsigned int* pair_hits;            /* [CKC_MAX_PAIRS] how often each table entry produced the first hit (decayed): adaptive processing order */
    const int* pair_order;              /* [CKC_MAX_PAIRS] table entries sorted by decreasing pair_hits; any order yields the same boolean */
    const double*   tris;               /* all meshes, concatenated, 9 doubles per triangle (FP64, as parsed) */
    const float4* body_tab;             /* [nbodies] distinct (frame, mesh) combinations of the pair table: root-box centre (mesh frame) xyz, frame index as int bits in w;
                                           a pair entry names its two bodies in pad1 (body_a | body_b << 8) -- k_collide transforms each centre once per configuration */
    int nbodies;
    int node_off[CKC_NUM_MESHES];       /* first node of each mesh */
    int tri_off[CKC_NUM_MESHES];        /* first triangle of each mesh */
    int npairs;
    int env;
    ckc_pair pairs[CKC_MAX_PAIRS];
    double static_frames[3][12];        /* obstacle poses (frames 18..20): R row-major (9) + T (3) */
    double tol;                         /* 8.0 mm, collisionDetection.cpp:263 */
    float  cull_tol;                    /* tol + FP32 guard band used by the conservative box test */
    double D;                           /* robots distance (offsetX) */
    double base2_rot[9];                /* R02 = Rz(3.14159265/2)^2 as the reference computes it */
    double base_t[2][3];                /* T0, T02 (incl. the aliased MxV quirk) */
};

/* Kinematic constants (apc_class.cpp:5-25, kdl_class.cpp:5-45), precomputed on the host with the same libm calls */
struct ckc_kin_dev {
    double b, l1, l2, l3a, l3b, l4, l5, lee;    /* APC values: l5 = 65.5, lee = 66.5 */
    double L, D;
    double lim_lo[6], lim_hi[6];
    double base_x[2], base_y[2], base_c[2], base_s[2];   /* robot base poses: x, y, cos(theta), sin(theta) */
    double l34, alpha;                           /* sqrt(l3a^2+(l3b+l4)^2), atan2(l3b+l4, l3a) */
    double kdl_tip[13][3];                       /* KDL chain segment tips (integer-division l5/lee = 66) */
    int    kdl_axis[13];                         /* -1 none, 0 X, 1 Y, 2 Z */
    int    env_index;                            /* env - 1: selects the __constant__ chain table */
};

/* how a batch of configurations is addressed: element (i, j) = base[i * stride_i + j * stride_j] */
struct ckc_layout { int64_t stride_i, stride_j; };

struct ckc_counters_dev {       /* device-side accumulators, one struct per context */
    unsigned long long ik_calls, ik_feasible, collision_calls, collisions, bv_tests, tri_tests, queue_overflows,
        gd_iterations, is_valid_calls;
};

/* RBS frontier state (device pointers). Nodes = configurations (AoS, 12 doubles); a segment joins two nodes. */
struct ckc_rbs_dev {
    double* nodes;                       /* node pool: [2*n_edges endpoints | midpoints ...] */
    double* node_t;                      /* reconstructRBS only (else NULL): position of each node along its edge, 0 = a, 1 = b, midpoints (ta + tb) / 2 */
    int32_t *seg_a, *seg_b, *seg_edge, *seg_depth;          /* current level */
    int32_t *next_a, *next_b, *next_edge, *next_depth;      /* next level */
    int32_t *cand_node, *cand_seg;       /* midpoints that passed projection and await the collision check */
    int32_t *cand_count, *next_count;
    uint8_t* edge_ok;                    /* per edge: 1 until any node of its bisection tree is invalid */
    int32_t* nchecks;                    /* per edge: isValidRBS evaluations (may be NULL) */
    const int8_t *edge_ac, *edge_ik;     /* PCS flavour: active chain / IK branch per edge */
    double tol;                          /* RBS_tol = 0.05 */
    int max_depth;                       /* RBS_max_depth = 150 */
};

/* ---- launchers implemented in ckc_kernels.cu (all asynchronous on `st`) ---- */
void ckc_launch_rbs_init(const ckc_rbs_dev& r, int n_edges, const double* in, cudaStream_t st);
void ckc_launch_rbs_expand(const ckc_kin_dev& kin, const ckc_rbs_dev& r, int flavour, int nseg, int node_base, cudaStream_t st);
void ckc_launch_rbs_push(const ckc_rbs_dev& r, int max_cand, const int32_t* d_ncand, const uint8_t* hit, cudaStream_t st);
struct ckc_launch_cfg { int num_sms; };
/* host mirror of the state block of the RBS tail kernel (ckc_kernels.cu: rbs_tail_state) */
struct ckc_rbs_tail_state { int32_t nseg, nnodes, cur, levels, status, total_cand, max_seg, seg_cap, cand_cap, node_cap, max_levels; };
int ckc_launch_rbs_tail(const ckc_kin_dev& kin, const ckc_world_dev& w, const ckc_launch_cfg& cfg, const ckc_rbs_dev& r0, const ckc_rbs_dev& r1,
                        int32_t* work_counter, int32_t* ovf_list, int32_t* ovf_count, uint32_t* big_stack, ckc_counters_dev* ctr, void* dstate, uint8_t* hit,
                        cudaStream_t st);

/* K1: projection. src == NULL => generate S-PCS samples (seed, i0).  Writes q_out (layout lo), ok, zeroes valid,
 * appends feasible sample indices to list[] (count in *list_count). */
void ckc_launch_pcs_project(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout li, const int8_t* ac, const int8_t* ik,
                            uint64_t seed, uint64_t i0, int seeded, double* q_out, ckc_layout lo, uint8_t* ok, uint8_t* valid,
                            int32_t* list, int32_t* list_count, ckc_counters_dev* ctr, cudaStream_t st);
/* K3: link frames + collision of the configurations named by list[0..m) (or 0..m-1 when list == NULL); configuration
 * (sample, j) = q[sample * lq.stride_i + j * lq.stride_j]; hit flags / valid flags are indexed by the sample index (list[k]) */
void ckc_launch_collide(const ckc_world_dev& w, const ckc_launch_cfg& cfg, const double* q, ckc_layout lq, const int32_t* list, const int32_t* d_m,
                        int64_t m_max, uint8_t* hit_out, uint8_t* valid_out, uint8_t* pair_flags, int32_t* work_counter,
                        int32_t* ovf_list, int32_t* ovf_count, uint32_t* big_stack, ckc_counters_dev* ctr, cudaStream_t st);
void ckc_launch_update_pair_order(const ckc_world_dev& w, int* order_out, cudaStream_t st);
void ckc_launch_limits(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, uint8_t* ok, cudaStream_t st);
void ckc_launch_identify_ik(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, int8_t* ik01, cudaStream_t st);
void ckc_launch_isvalid_prepare(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, double* jobs, uint8_t* pre, int32_t* list,
                                int32_t* list_count, cudaStream_t st);
void ckc_launch_isvalid_combine(int64_t n, const uint8_t* pre, const uint8_t* hit, uint8_t* valid, cudaStream_t st);
void ckc_launch_gd_project(const ckc_kin_dev& kin, int num_sms, int64_t n, const double* q, ckc_layout li, uint64_t seed, uint64_t i0, int seeded,
                           double* q_out, ckc_layout lo, uint8_t* ok, uint8_t* converged, int32_t* iters, uint8_t* valid,
                           int32_t* list, int32_t* list_count, int32_t* state_scratch, unsigned long long* work_counter, ckc_counters_dev* ctr, cudaStream_t st);
void ckc_launch_rx_check(const ckc_kin_dev& kin, int64_t n, const double* q, ckc_layout l, double eps, uint8_t* ok, double* resid,
                         uint8_t* valid, int32_t* list, int32_t* list_count, cudaStream_t st);
void ckc_launch_fma_peak(int precision, int num_sms, int iters, float* sink, cudaStream_t st);
void ckc_launch_compact_valid(int64_t n, int32_t index_base, const uint8_t* valid, const double* q, ckc_layout l, int32_t* idx_out, double* q_out,
                              int32_t* count, cudaStream_t st);
/* single-arm kinematics exposed one to one: FKsolve_rob / IKsolve_rob (apc_class.cpp:51-76, :114-245), kdl::FK (kdl_class.cpp:251-278) */
void ckc_launch_fk(const ckc_kin_dev& kin, int64_t n, const double* q6, const int8_t* robot, double* T16, cudaStream_t st);
void ckc_launch_ik(const ckc_kin_dev& kin, int64_t n, const double* T16, const int8_t* robot, const int8_t* sol, double* q6, uint8_t* ok, cudaStream_t st);
void ckc_launch_kdl_fk(const ckc_kin_dev& kin, int64_t n, const double* q12, double* T16, cudaStream_t st);
void ckc_launch_transpose(const double* src, ckc_layout ls, double* dst, ckc_layout ld, int64_t n, cudaStream_t st);

#endif
