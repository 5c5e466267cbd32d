/*
 * ckc_device.cuh -- device-side math of libckc (sm_100a): counter RNG, IRB120 kinematics in registers, link frames,
 * the conservative FP32 box test and the FP64 triangle-pair distance.  Product code (no oracle includes).
 *
 * Reference behaviour each routine reproduces (paths relative to the reference root):
 *   fk_arm            two_robots::FKsolve_rob          proj_classes/apc_class.cpp:51-76
 *   ik_arm            two_robots::IKsolve_rob          proj_classes/apc_class.cpp:114-245
 *   project_passive   calc_specific_IK_solution_R1/R2  proj_classes/apc_class.cpp:356-389
 *   warp_link_frames  collision_state link frames      validity_checkers/collisionDetection.cpp:72-260 (one warp per configuration)
 *   tri_dist_coop     PQP v1.3 TriDist / SegPoints (external library used by collision_state), warp-cooperative
 */
#ifndef CKC_DEVICE_CUH_
#define CKC_DEVICE_CUH_

#include "ckc_internal.h"

#define CKC_PI_ 3.1416          /* apc_class.h:10, kdl_class.h:24 -- the reference's pi */

namespace ckc {

/* ---------------------------------------------------------------------------------------------------------------- */
/* counter-based generator (DESIGN.md "Synthetic inputs"): u(seed,i,k) = (mix(mix(seed) ^ (16 i + k)) >> 11) 2^-53     */
/* ---------------------------------------------------------------------------------------------------------------- */
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ double uniform01(uint64_t key, uint64_t i, uint32_t k) {
    return (double)(mix64(key ^ (i * 16ULL + k)) >> 11) * (1.0 / 9007199254740992.0);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* small rotation helpers: R is a 3x3 row-major array held in registers                                               */
/* ---------------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void post_rot_z(double* R, double c, double s) {     /* R <- R * Rz */
#pragma unroll
    for (int r = 0; r < 3; r++) { const double a = R[3 * r], b = R[3 * r + 1]; R[3 * r] = a * c + b * s; R[3 * r + 1] = b * c - a * s; }
}
__device__ __forceinline__ void post_rot_y(double* R, double c, double s) {     /* R <- R * Ry */
#pragma unroll
    for (int r = 0; r < 3; r++) { const double a = R[3 * r], b = R[3 * r + 2]; R[3 * r] = a * c - b * s; R[3 * r + 2] = a * s + b * c; }
}
__device__ __forceinline__ void post_rot_x(double* R, double c, double s) {     /* R <- R * Rx */
#pragma unroll
    for (int r = 0; r < 3; r++) { const double a = R[3 * r + 1], b = R[3 * r + 2]; R[3 * r + 1] = a * c + b * s; R[3 * r + 2] = b * c - a * s; }
}

/* tool frame of one arm: T = Trans(x0,y0,0) Rz(th) Tz(b+l1) Rz(q1) Ry(q2) Tz(l2) Ry(q3) T(l3b+l4,0,l3a) Rx(q4) Ry(q5) Tx(l5+lee) Rx(q6) */
__device__ __forceinline__ void fk_arm(const ckc_kin_dev& k, const double* q, int robot, double* R, double* p) {
    const double bc = k.base_c[robot], bs = k.base_s[robot];
    R[0] = bc; R[1] = -bs; R[2] = 0; R[3] = bs; R[4] = bc; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    p[0] = k.base_x[robot]; p[1] = k.base_y[robot]; p[2] = k.b + k.l1;
    double s, c;
    sincos(q[0], &s, &c); post_rot_z(R, c, s);
    sincos(q[1], &s, &c); post_rot_y(R, c, s);
    p[0] += R[2] * k.l2; p[1] += R[5] * k.l2; p[2] += R[8] * k.l2;
    sincos(q[2], &s, &c); post_rot_y(R, c, s);
    const double ax = k.l3b + k.l4;
    p[0] += R[0] * ax + R[2] * k.l3a; p[1] += R[3] * ax + R[5] * k.l3a; p[2] += R[6] * ax + R[8] * k.l3a;
    sincos(q[3], &s, &c); post_rot_x(R, c, s);
    sincos(q[4], &s, &c); post_rot_y(R, c, s);
    const double le = k.l5 + k.lee;
    p[0] += R[0] * le; p[1] += R[3] * le; p[2] += R[6] * le;
    sincos(q[5], &s, &c); post_rot_x(R, c, s);
}

__device__ __forceinline__ double wrap_pi_(double a) {           /* apc_class.cpp:142-145 */
    if (a > CKC_PI_) a -= 2 * CKC_PI_;
    if (a < -CKC_PI_) a += 2 * CKC_PI_;
    return a;
}

/* One analytic IK branch.  Rt (row-major) / pt: target pose in the world frame; returns false on the reference's
 * early-outs (joint limit breach, unreachable). */
__device__ __forceinline__ bool ik_arm(const ckc_kin_dev& k, const double* Rt, const double* pt, int robot, int sol, double* q) {
    const double add1 = (sol & 4) ? CKC_PI_ : 0.0;                        /* q1_add_sol  :117 */
    const double sg2 = (sol & 4) ? 1.0 : -1.0;                            /* q2_sign     :118 */
    const double sg3 = (sol & 1) ? -1.0 : 1.0;                            /* q3_sign     :119 */
    const double sg = ((0x96 >> sol) & 1) ? -1.0 : 1.0;                   /* sign456 = {1,-1,-1,1,-1,1,1,-1}  :120 */
    const double cth = k.base_c[robot], sth = k.base_s[robot], x0 = k.base_x[robot], y0 = k.base_y[robot];
    double R[9], p[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {                                         /* :128-130 */
        R[j] = cth * Rt[j] + sth * Rt[3 + j];
        R[3 + j] = cth * Rt[3 + j] - sth * Rt[j];
        R[6 + j] = Rt[6 + j];
    }
    p[0] = cth * pt[0] + sth * pt[1] - cth * x0 - sth * y0;              /* :132-134 */
    p[1] = cth * pt[1] - sth * pt[0] - cth * y0 + sth * x0;
    p[2] = pt[2];
    const double le = k.l5 + k.lee;
    const double p5x = p[0] - R[0] * le, p5y = p[1] - R[3] * le, p5z = p[2] - R[6] * le;     /* :137 */

    q[0] = wrap_pi_(atan2(p5y, p5x) + add1);                              /* :141-148 */
    if (fabs(q[0]) > k.lim_hi[0]) return false;

    const double k2 = p5x * p5x + p5y * p5y;
    const double zz = p5z - (k.l1 + k.b);
    const double D2 = k2 + zz * zz;
    const double cosphi = (D2 - k.l2 * k.l2 - k.l34 * k.l34) / (2 * k.l2 * k.l34);   /* :160 */
    double sinphi = 1 - cosphi * cosphi;
    if (sinphi < 0) return false;                                         /* :164 */
    sinphi = sg3 * sqrt(sinphi);
    q[2] = wrap_pi_(-(atan2(sinphi, cosphi) + k.alpha));                  /* :168-172 */
    if (q[2] < k.lim_lo[2] || q[2] > k.lim_hi[2]) return false;

    const double sa1 = -k.l34 * sinphi / sqrt(D2);                        /* :180 */
    const double alpha1 = atan2(-sg2 * sa1, sqrt(1 - sa1 * sa1));
    const double alpha2 = atan2(p5z - k.l1 - k.b, sqrt(k2));
    q[1] = wrap_pi_(sg2 * (alpha1 + alpha2 - CKC_PI_ / 2));               /* :182-186 */
    if (fabs(q[1]) > k.lim_hi[1]) return false;

    double s23, c23, s1, c1;
    sincos(q[1] + q[2], &s23, &c23);
    sincos(q[0], &s1, &c1);
    const double S = R[0] * c23 * c1 - R[6] * s23 + R[3] * c23 * s1;      /* :206 */
    q[4] = atan2(sg * sqrt(1 - S * S), S);
    q[4] = fabs(q[4]) < 1e-4 ? 0 : q[4];
    if (fabs(q[4]) > k.lim_hi[4]) return false;
    if (q[4] != 0) {                                                      /* :215-225 */
        const double sinq4 = R[3] * c1 - R[0] * s1;
        const double cosq4 = R[6] * c23 + R[0] * s23 * c1 + R[3] * s23 * s1;
        q[3] = atan2(sg * sinq4, -sg * cosq4);
        const double sinq6 = R[1] * c23 * c1 - R[7] * s23 + R[4] * c23 * s1;
        const double cosq6 = R[2] * c23 * c1 - R[8] * s23 + R[5] * c23 * s1;
        q[5] = atan2(sg * sinq6, sg * cosq6);
    } else {                                                              /* :226-231 */
        q[3] = 0;
        q[5] = atan2(R[7], R[8]);
    }
    if (fabs(q[3]) > k.lim_hi[3] || fabs(q[5]) > k.lim_hi[5]) return false;
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (fabs(q[i]) < 1e-5) q[i] = 0;                                  /* :236-238 */
    return true;
}

/* closure: the passive arm must hold the other end of the rod: T_passive = T_active * Trans(+-L) * diag(-1,-1,1,1)
 * (apc_class.cpp:360-362 and :381-383; both reduce to: columns 0,1 negated, origin moved by +L along tool x) */
__device__ __forceinline__ bool project_passive(const ckc_kin_dev& k, const double* q_active, int active_robot, int sol, double* q_passive) {
    double R[9], p[3];
    fk_arm(k, q_active, active_robot, R, p);
    double Rt[9], pt[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        Rt[3 * r] = -R[3 * r]; Rt[3 * r + 1] = -R[3 * r + 1]; Rt[3 * r + 2] = R[3 * r + 2];
        pt[r] = R[3 * r] * k.L + p[r];
    }
    return ik_arm(k, Rt, pt, 1 - active_robot, sol, q_passive);
}

__device__ __forceinline__ bool within_limits(const ckc_kin_dev& k, const double* q) {       /* apc_class.cpp:304-334 */
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const double* a = q + 6 * r;
        ok = ok && !(fabs(a[0]) > k.lim_hi[0]) && !(fabs(a[1]) > k.lim_hi[1]) && !(a[2] < k.lim_lo[2] || a[2] > k.lim_hi[2]) &&
             !(fabs(a[3]) > k.lim_hi[3]) && !(fabs(a[4]) > k.lim_hi[4]) && !(fabs(a[5]) > k.lim_hi[5]);
    }
    return ok;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* link frames of both robots for the collision meshes (collisionDetection.cpp:72-166 / :177-260), computed by the     */
/* WARP that collides the configuration, straight into its shared-memory frame table (no HBM round trip):              */
/*   lanes 0..11   sin / cos of joint `lane` (FP64), one joint each;                                                     */
/*   lanes 0..5    "row lanes": lane = 3 * robot + row carries row `row` of that robot's running rotation and component  */
/*                 `row` of its running translation along the chain -- MRot{Z,Y,Y,X,Y,X} only mix entries of one row,    */
/*                 and T += R * offset only needs that row, so the three rows of a robot are independent.               */
/* F64: [CKC_NUM_FRAMES][12] doubles (R row-major 9 + T 3): 8 frames per robot (base, link1..link6, EE / rod pose        */
/* (R6, T7)), frames 16 / 17 = the two poses the reference's pair table mixes up (collisionDetection.cpp:373, :397,     */
/* :415: robot-2 rotation with robot-1 translation of link 2 and vice versa).  base: [2][12] robot base frames.         */
/* ALL 32 lanes must call this (shuffles inside).                                                                       */
/* ---------------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ double shfl_f64(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}
__device__ __forceinline__ void sincos_compact(double x, double* sn, double* cs);      /* defined with the GD kernel's tables below */
__device__ __noinline__ void warp_link_frames(const double* __restrict__ base, double qj, int lane, double* F64) {
    /* noinline + the compact sin / cos (no inlined Payne-Hanek path): k_collide's loop body must stay inside the 32 kB
     * instruction cache level (measured: with libdevice's sincos inlined here the kernel grew by 17 kB and `no_instruction`
     * became its second largest stall) */
    double sj = 0.0, cj = 1.0;
    if (lane < 12) sincos_compact(qj, &sj, &cj);
    const int rob = lane >= 3 ? 1 : 0, row = lane - 3 * rob;
    const bool rl = lane < 6;
    double R0 = 0, R1 = 0, R2 = 0, T = 0;
    if (rl) { const double* b = base + 12 * rob; R0 = b[3 * row]; R1 = b[3 * row + 1]; R2 = b[3 * row + 2]; T = b[9 + row]; }
    double* out = F64 + 96 * rob;
    auto put = [&](int f) { if (rl) { double* d = out + 12 * f; d[3 * row] = R0; d[3 * row + 1] = R1; d[3 * row + 2] = R2; d[9 + row] = T; } };
    const int src = 6 * rob;
    double c, s, a, b;
    put(0);                                                               /* base */
    T += R2 * 290.0;                                                      /* T1 = T0 + R0 (0,0,145*2) */
    c = shfl_f64(cj, src); s = shfl_f64(sj, src);
    a = R0; b = R1; R0 = a * c + b * s; R1 = b * c - a * s;               /* R1 = R0 Rz(q1) */
    put(1);
    c = shfl_f64(cj, src + 1); s = shfl_f64(sj, src + 1);
    a = R0; b = R2; R0 = a * c - b * s; R2 = a * s + b * c;               /* R2 = R1 Ry(q2), T2 = T1 */
    put(2);
    if (rl) {                                                             /* the pair table's mixed-up link-2 poses */
        double* mixR = F64 + 12 * (rob == 0 ? 17 : 16); double* mixT = F64 + 12 * (rob == 0 ? 16 : 17);
        mixR[3 * row] = R0; mixR[3 * row + 1] = R1; mixR[3 * row + 2] = R2; mixT[9 + row] = T;
    }
    T += R2 * 270.0;                                                      /* T3 = T2 + R2 (0,0,270) */
    c = shfl_f64(cj, src + 2); s = shfl_f64(sj, src + 2);
    a = R0; b = R2; R0 = a * c - b * s; R2 = a * s + b * c;               /* R3 = R2 Ry(q3) */
    put(3);
    T += R0 * 134.0 + R2 * 70.0;                                          /* T4 */
    c = shfl_f64(cj, src + 3); s = shfl_f64(sj, src + 3);
    a = R1; b = R2; R1 = a * c + b * s; R2 = b * c - a * s;               /* R4 = R3 Rx(q4) */
    put(4);
    T += R0 * 168.0;                                                      /* T5 */
    c = shfl_f64(cj, src + 4); s = shfl_f64(sj, src + 4);
    a = R0; b = R2; R0 = a * c - b * s; R2 = a * s + b * c;               /* R5 = R4 Ry(q5) */
    put(5);
    T += R0 * 66.0;                                                       /* T6: 72 - 13/2 = 66 (integer division) */
    c = shfl_f64(cj, src + 5); s = shfl_f64(sj, src + 5);
    a = R1; b = R2; R1 = a * c + b * s; R2 = b * c - a * s;               /* R6 = R5 Rx(q6) */
    put(6);
    T += R0 * 66.0;                                                       /* T7: 13/2 + 60 = 66 */
    put(7);                                                               /* EE and rod pose (R6, T7) */
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* conservative FP32 box-box test: true => the two boxes are farther apart than `tol` (tol already includes the       */
/* guard band).  Fa/Fb = world frames (R row-major 9 + T 3) of the meshes the boxes belong to.                        */
/* ---------------------------------------------------------------------------------------------------------------- */
struct box32 { float c[3], h[3], a0[3], a1[3]; float size2; int child; };

__device__ __forceinline__ box32 load_box(const ckc_node* n) {
    const float4* p = reinterpret_cast<const float4*>(n);
    const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
    box32 b;
    b.c[0] = v0.x; b.c[1] = v0.y; b.c[2] = v0.z; b.size2 = v0.w;
    b.h[0] = v1.x; b.h[1] = v1.y; b.h[2] = v1.z; b.child = __float_as_int(v1.w);
    b.a0[0] = v2.x; b.a0[1] = v2.y; b.a0[2] = v2.z;
    b.a1[0] = v3.x; b.a1[1] = v3.y; b.a1[2] = v3.z;
    return b;
}

__device__ __forceinline__ void box_to_world(const box32& b, const float* F, float* c, float (*ax)[3]) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
        c[r] = F[3 * r] * b.c[0] + F[3 * r + 1] * b.c[1] + F[3 * r + 2] * b.c[2] + F[9 + r];
        ax[0][r] = F[3 * r] * b.a0[0] + F[3 * r + 1] * b.a0[1] + F[3 * r + 2] * b.a0[2];
        ax[1][r] = F[3 * r] * b.a1[0] + F[3 * r + 1] * b.a1[1] + F[3 * r + 2] * b.a1[2];
    }
    ax[2][0] = ax[0][1] * ax[1][2] - ax[0][2] * ax[1][1];
    ax[2][1] = ax[0][2] * ax[1][0] - ax[0][0] * ax[1][2];
    ax[2][2] = ax[0][0] * ax[1][1] - ax[0][1] * ax[1][0];
}

/* gmax6 (out): largest gap along the 6 face axes -- a cheap lower bound of the box distance used as a traversal hint only */
__device__ __forceinline__ bool boxes_farther_than(const box32& A, const float* Fa, const box32& B, const float* Fb, float tol, float& gmax6) {
    float ca[3], cb[3], ua[3][3], ub[3][3];
    box_to_world(A, Fa, ca, ua);
    box_to_world(B, Fb, cb, ub);
    const float d[3] = { cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2] };
    float t[3], R[3][3], AR[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t[i] = d[0] * ua[i][0] + d[1] * ua[i][1] + d[2] * ua[i][2];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            R[i][j] = ua[i][0] * ub[j][0] + ua[i][1] * ub[j][1] + ua[i][2] * ub[j][2];
            AR[i][j] = fabsf(R[i][j]) + 1e-5f;        /* keeps the test conservative under FP32 rounding */
        }
    }
    bool far = false;
    float gm = -1e30f;
#pragma unroll
    for (int i = 0; i < 3; i++)
        gm = fmaxf(gm, fabsf(t[i]) - (A.h[i] + B.h[0] * AR[i][0] + B.h[1] * AR[i][1] + B.h[2] * AR[i][2]));
#pragma unroll
    for (int j = 0; j < 3; j++)
        gm = fmaxf(gm, fabsf(t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j]) - (B.h[j] + A.h[0] * AR[0][j] + A.h[1] * AR[1][j] + A.h[2] * AR[2][j]));
    gmax6 = gm;
    far = gm > tol;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            const float len2 = 1.0f - R[i][j] * R[i][j];             /* |a_i x b_j|^2 */
            const float g = fabsf(t[i2] * R[i1][j] - t[i1] * R[i2][j]) - (A.h[i1] * AR[i2][j] + A.h[i2] * AR[i1][j]) -
                            (B.h[j1] * AR[i][j2] + B.h[j2] * AR[i][j1]);
            far = far | ((len2 > 1e-3f) & (g > 0.0f) & (g * g > tol * tol * len2));       /* bitwise: no short-circuit branches */
        }
    }
    return far;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* FP64 triangle-pair distance with PQP v1.3 TriDist / SegPoints semantics (the value PQP_Tolerance compares to 8 mm)  */
/* ---------------------------------------------------------------------------------------------------------------- */
struct v3 { double x, y, z; };
__device__ __forceinline__ v3 operator-(const v3& a, const v3& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
__device__ __forceinline__ v3 operator+(const v3& a, const v3& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
__device__ __forceinline__ v3 madd(const v3& a, const v3& b, double s) { return { a.x + b.x * s, a.y + b.y * s, a.z + b.z * s }; }
__device__ __forceinline__ double dot(const v3& a, const v3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ v3 cross(const v3& a, const v3& b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }

/* vertex-vs-face case of TriDist: the vertices V[] of one triangle against the plane of the other (F0.., edges E[], normal N) */
__device__ __forceinline__ bool face_case(const v3* F, const v3* E, const v3* V, bool& shown_disjoint, double& out_d) {
    const v3 N = cross(E[0], E[1]);
    const double Nl = dot(N, N);
    if (!(Nl > 1e-15)) return false;
    double pr[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pr[k] = dot(F[0] - V[k], N);
    int point = -1;
    if ((pr[0] > 0) && (pr[1] > 0) && (pr[2] > 0)) { point = (pr[0] < pr[1]) ? 0 : 1; if (pr[2] < pr[point]) point = 2; }
    else if ((pr[0] < 0) && (pr[1] < 0) && (pr[2] < 0)) { point = (pr[0] > pr[1]) ? 0 : 1; if (pr[2] > pr[point]) point = 2; }
    if (point < 0) return false;
    shown_disjoint = true;
    const v3 Vp = point == 0 ? V[0] : (point == 1 ? V[1] : V[2]);
    const double prp = point == 0 ? pr[0] : (point == 1 ? pr[1] : pr[2]);
    if (dot(Vp - F[0], cross(N, E[0])) > 0 && dot(Vp - F[1], cross(N, E[1])) > 0 && dot(Vp - F[2], cross(N, E[2])) > 0) {
        const v3 Pf = madd(Vp, N, prp / Nl);          /* foot of the perpendicular on the face */
        const v3 dd = Pf - Vp;
        out_d = sqrt(dot(dd, dd));
        return true;
    }
    return false;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Warp-cooperative TriDist: NINE lanes per triangle pair (one per edge pair (i, j), i = e / 3, j = e % 3), three pairs per    */
/* warp step, everything in registers.  The sequential PQP loop visits the edge pairs in order k = 0..8 and                 */
/*   - "updates" at k when dd_k <= min(init, dd_0..dd_{k-1})   (init = |S0 - T0|^2 + 1)                                    */
/*   - returns sqrt(dd_k) at the FIRST updating k with a_k <= 0 and b_k >= 0                                                */
/*   - otherwise ORs (p_k - max(a_k,0) + min(b_k,0) > 0) over the updating k into shown_disjoint and keeps the prefix min. */
/* Each dd_k, a_k, b_k, p_k is produced by exactly the arithmetic of the scalar routine, so prefix-min + ballots reproduce    */
/* its result bit for bit; the vertex-face cases and the final "shown_disjoint ? sqrt(mindd) : 0" run on lane e = 0.         */
/* ALL 32 lanes must call this (shuffles inside); lanes with act == false carry dummy data.                                  */
/* ---------------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void seg_points_inl(v3& VEC, v3& X, v3& Y, const v3& P, const v3& A, const v3& Q, const v3& B) {
    v3 T = Q - P;
    const double AA = dot(A, A), BB = dot(B, B), AB = dot(A, B), AT = dot(A, T), BT = dot(B, T);
    const double denom = AA * BB - AB * AB;
    double t = (AT * BB - BT * AB) / denom;
    if ((t < 0) || isnan(t)) t = 0; else if (t > 1) t = 1;
    const double u = (t * AB - BT) / BB;
    if ((u <= 0) || isnan(u)) {
        Y = Q;
        t = AT / AA;
        if ((t <= 0) || isnan(t)) { X = P; VEC = Q - P; }
        else if (t >= 1) { X = P + A; VEC = Q - X; }
        else { X = madd(P, A, t); VEC = cross(A, cross(T, A)); }
    } else if (u >= 1) {
        Y = Q + B;
        t = (AB + AT) / AA;
        if ((t <= 0) || isnan(t)) { X = P; VEC = Y - P; }
        else if (t >= 1) { X = P + A; VEC = Y - X; }
        else { X = madd(P, A, t); T = Y - P; VEC = cross(A, cross(T, A)); }
    } else {
        Y = madd(Q, B, u);
        if ((t <= 0) || isnan(t)) { X = P; VEC = cross(B, cross(T, B)); }
        else if (t >= 1) { X = P + A; T = Q - X; VEC = cross(B, cross(T, B)); }
        else {
            X = madd(P, A, t);
            VEC = cross(A, B);
            if (dot(VEC, T) < 0) { VEC.x = -VEC.x; VEC.y = -VEC.y; VEC.z = -VEC.z; }
        }
    }
}

__device__ __noinline__ double face_and_final(const v3 S0, const v3 S1, const v3 S2, const v3 T0, const v3 T1, const v3 T2, bool shown_disjoint, double mindd) {
    const v3 S[3] = { S0, S1, S2 }, T[3] = { T0, T1, T2 };
    const v3 Sv[3] = { S1 - S0, S2 - S1, S0 - S2 }, Tv[3] = { T1 - T0, T2 - T1, T0 - T2 };
    double d;
    if (face_case(S, Sv, T, shown_disjoint, d)) return d;
    if (face_case(T, Tv, S, shown_disjoint, d)) return d;
    return shown_disjoint ? sqrt(mindd) : 0.0;
}

/* S0..S2: triangle of mesh A; T0..T2: triangle of mesh B already in A's frame (identical values on the 9 lanes of a group) */
__device__ __forceinline__ double tri_dist_coop(bool act, int lane, const v3& S0, const v3& S1, const v3& S2, const v3& T0, const v3& T1, const v3& T2) {
    const int g = lane / 9, e = lane - 9 * g;
    const int i = e / 3, j = e - 3 * i;
    /* this lane's edge pair: edge i of S = (Si, Si+1), opposite vertex Si+2; edge j of T likewise */
    const v3 Sa = i == 0 ? S0 : (i == 1 ? S1 : S2), Sb = i == 0 ? S1 : (i == 1 ? S2 : S0), Sc = i == 0 ? S2 : (i == 1 ? S0 : S1);
    const v3 Ta = j == 0 ? T0 : (j == 1 ? T1 : T2), Tb = j == 0 ? T1 : (j == 1 ? T2 : T0), Tc = j == 0 ? T2 : (j == 1 ? T0 : T1);
    v3 VEC, P, Q;
    seg_points_inl(VEC, P, Q, Sa, Sb - Sa, Ta, Tb - Ta);
    const v3 V = Q - P;
    const double dd = dot(V, V);
    double a = dot(Sc - P, VEC);
    double b = dot(Tc - Q, VEC);
    const bool ret = (a <= 0) && (b >= 0);
    const double pp = dot(V, VEC);
    if (a < 0) a = 0;
    if (b > 0) b = 0;
    const bool sd = (pp - a + b) > 0;
    /* prefix minimum in edge-pair order over the 9 lanes of the group */
    double m;
    { const v3 d0 = S0 - T0; m = dot(d0, d0) + 1; }            /* mindd initialisation of the scalar routine */
    const int base = 9 * g;
    bool upd = false;
    double mall = m;
#pragma unroll
    for (int t = 0; t < 9; t++) {
        const int src = base + t > 31 ? 31 : base + t;
        const double dt = shfl_f64(dd, src);
        if (t == e) upd = dd <= mall;                           /* compares with min(init, dd_0..dd_{e-1}) */
        mall = (dt <= mall) ? dt : mall;                        /* running minimum exactly as "if (dd <= mindd) mindd = dd" */
    }
    const unsigned grp = (g < 3) ? (0x1ffu << base) : 0u;
    const unsigned m_ret = __ballot_sync(0xffffffffu, act && upd && ret) & grp;
    const unsigned m_sd = __ballot_sync(0xffffffffu, act && upd && sd) & grp;
    /* every shuffle below is executed by all 32 lanes (group-specific source lanes, no divergent *_sync) */
    const int src_first = m_ret ? (__ffs(m_ret) - 1) : (base > 31 ? 31 : base);     /* first updating edge pair with the early-return condition */
    const double dd_first = shfl_f64(dd, src_first);
    double result = 0;
    if (m_ret) result = sqrt(dd_first);                       /* shown_disjoint contributions after `first` are irrelevant: the routine has returned */
    else if (act && e == 0) result = face_and_final(S0, S1, S2, T0, T1, T2, m_sd != 0, mall);
    result = shfl_f64(result, base > 31 ? 31 : base);         /* lane e = 0 of the group holds the value in both cases */
    return result;
}

/* Sequential TriDist, one THREAD per triangle pair (the heavy-configuration kernel): the loop of the scalar routine over the nine
 * edge pairs with exactly the per-edge-pair arithmetic of the cooperative version above (same seg_points_inl, same a / b / p), so
 * both return the same value. */
__device__ __noinline__ double tri_dist_seq(const v3 S0, const v3 S1, const v3 S2, const v3 T0, const v3 T1, const v3 T2) {
    double mindd;
    { const v3 d0 = S0 - T0; mindd = dot(d0, d0) + 1; }
    bool shown_disjoint = false;
#pragma unroll 1
    for (int e = 0; e < 9; e++) {
        const int i = e / 3, j = e - 3 * i;
        const v3 Sa = i == 0 ? S0 : (i == 1 ? S1 : S2), Sb = i == 0 ? S1 : (i == 1 ? S2 : S0), Sc = i == 0 ? S2 : (i == 1 ? S0 : S1);
        const v3 Ta = j == 0 ? T0 : (j == 1 ? T1 : T2), Tb = j == 0 ? T1 : (j == 1 ? T2 : T0), Tc = j == 0 ? T2 : (j == 1 ? T0 : T1);
        v3 VEC, P, Q;
        seg_points_inl(VEC, P, Q, Sa, Sb - Sa, Ta, Tb - Ta);
        const v3 V = Q - P;
        const double dd = dot(V, V);
        if (dd <= mindd) {
            mindd = dd;
            double a = dot(Sc - P, VEC);
            double b = dot(Tc - Q, VEC);
            if ((a <= 0) && (b >= 0)) return sqrt(dd);
            const double pp = dot(V, VEC);
            if (a < 0) a = 0;
            if (b > 0) b = 0;
            if ((pp - a + b) > 0) shown_disjoint = true;
        }
    }
    return face_and_final(S0, S1, S2, T0, T1, T2, shown_disjoint, mindd);
}

}  // namespace ckc

/* ================================================================================================================ */
/* GD flavour: the 12-joint serial chain of kdl_class.cpp:32-45 and the Newton-Raphson closure projection of          */
/* kdl::GD (kdl_class.cpp:68-140) = KDL ChainIkSolverPos_NR(fk_recursive, vel_pinv(eps 1e-5), 10000, 1e-5).           */
/* ================================================================================================================ */
namespace ckc {

/* chain FK in chain joint order.  Optionally the columns of the 6x12 geometric Jacobian referred to the BASE ORIGIN:
 * col0_j = [ a_j x (-o_j) ; a_j ].  The tip-referred Jacobian of KDL's ChainJntToJacSolver is J = X col0 with
 * X = [[I, -[p]x],[0, I]] (velocity of the tip point = v_origin + w x p). */
template <bool kJac>
__device__ __forceinline__ void chain_fk(const ckc_kin_dev& k, const double* qc, double* R, double* p, double (*J0)[12]) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    p[0] = 0; p[1] = 0; p[2] = 0;
    /* joint axes of the 13 segments (kdl_class.cpp:32-45): None, Z, Y, Y, X, Y, X | X, Y, X, Y, Y, Z -- compile-time so
     * that the unrolled walk indexes registers statically */
    constexpr int kAxis[13] = { -1, 2, 1, 1, 0, 1, 0, 0, 1, 0, 1, 1, 2 };
    int j = 0;
#pragma unroll
    for (int i = 0; i < 13; i++) {
        const int ax = kAxis[i];
        if (i > 0) {            /* segment 0 has no joint (Joint::None) */
            if (kJac) {
                const double a0 = R[ax], a1 = R[3 + ax], a2 = R[6 + ax];
                J0[0][j] = p[1] * a2 - p[2] * a1;       /* a x (-o) = o x a */
                J0[1][j] = p[2] * a0 - p[0] * a2;
                J0[2][j] = p[0] * a1 - p[1] * a0;
                J0[3][j] = a0; J0[4][j] = a1; J0[5][j] = a2;
            }
            double s, c;
            sincos(qc[j], &s, &c);
            if (ax == 0) post_rot_x(R, c, s); else if (ax == 1) post_rot_y(R, c, s); else post_rot_z(R, c, s);
            j++;
        }
        const double tx = k.kdl_tip[i][0], ty = k.kdl_tip[i][1], tz = k.kdl_tip[i][2];
        p[0] += R[0] * tx + R[1] * ty + R[2] * tz;
        p[1] += R[3] * tx + R[4] * ty + R[5] * tz;
        p[2] += R[6] * tx + R[7] * ty + R[8] * tz;
    }
}

/* KDL Rotation::GetRot() = axis * angle via GetRotAngle(axis, eps = 1e-6): rotation vector of M (row-major) */
__device__ __forceinline__ void kdl_rotvec(const double* M, double* rv) {
    const double eps = 1e-6, eps2 = 1e-5;
    if (fabs(M[1] - M[3]) < eps && fabs(M[2] - M[6]) < eps && fabs(M[5] - M[7]) < eps) {
        if (fabs(M[1] + M[3]) < eps2 && fabs(M[2] + M[6]) < eps2 && fabs(M[5] + M[7]) < eps2 && fabs(M[0] + M[4] + M[8] - 3) < eps2) {
            rv[0] = rv[1] = rv[2] = 0; return;
        }
        const double angle = 3.14159265358979323846;
        const double xx = (M[0] + 1) / 2, yy = (M[4] + 1) / 2, zz = (M[8] + 1) / 2;
        const double xy = (M[1] + M[3]) / 4, xz = (M[2] + M[6]) / 4, yz = (M[5] + M[7]) / 4;
        double x, y, z;
        if (xx > yy && xx > zz) { x = sqrt(xx); y = xy / x; z = xz / x; }
        else if (yy > zz) { y = sqrt(yy); x = xy / y; z = yz / y; }
        else { z = sqrt(zz); x = xz / z; y = yz / z; }
        rv[0] = x * angle; rv[1] = y * angle; rv[2] = z * angle; return;
    }
    const double f = (M[0] + M[4] + M[8] - 1) / 2;
    double x = M[7] - M[5], y = M[2] - M[6], z = M[3] - M[1];
    const double n = sqrt(x * x + y * y + z * z);
    const double angle = atan2(n / 2, f);
    if (n > 0) { x /= n; y /= n; z /= n; }
    rv[0] = x * angle; rv[1] = y * angle; rv[2] = z * angle;
}

/* exact truncated pseudo-inverse step dq = V S^+ U^T tw (singular values < eps dropped) by one-sided Jacobi on J^T.
 * Only reached near singular configurations (the Cholesky path below covers the regular case). */
__device__ __noinline__ void pinv_step_svd(const double (*J)[12], const double* tw, double eps, double* dq) {
    double B[12][6], W[6][6];
    for (int r = 0; r < 12; r++) for (int c = 0; c < 6; c++) B[r][c] = J[c][r];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) W[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 150; sweep++) {
        double off = 0;
        for (int a = 0; a < 6; a++)
            for (int b = a + 1; b < 6; b++) {
                double aa = 0, bb = 0, ab = 0;
                for (int r = 0; r < 12; r++) { aa += B[r][a] * B[r][a]; bb += B[r][b] * B[r][b]; ab += B[r][a] * B[r][b]; }
                if (fabs(ab) <= 1e-300 || fabs(ab) <= 1e-17 * sqrt(aa * bb)) continue;
                off = fmax(off, fabs(ab) / sqrt(aa * bb));
                const double zeta = (bb - aa) / (2 * ab);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                const double c = 1 / sqrt(1 + t * t), s = c * t;
                for (int r = 0; r < 12; r++) { const double x = B[r][a], y = B[r][b]; B[r][a] = c * x - s * y; B[r][b] = s * x + c * y; }
                for (int r = 0; r < 6; r++) { const double x = W[r][a], y = W[r][b]; W[r][a] = c * x - s * y; W[r][b] = s * x + c * y; }
            }
        if (off < 1e-15) break;
    }
    for (int r = 0; r < 12; r++) dq[r] = 0;
    for (int c = 0; c < 6; c++) {
        double s2 = 0;
        for (int r = 0; r < 12; r++) s2 += B[r][c] * B[r][c];
        if (sqrt(s2) < eps) continue;
        double ut = 0;
        for (int r = 0; r < 6; r++) ut += W[r][c] * tw[r];
        const double coef = ut / s2;
        for (int r = 0; r < 12; r++) dq[r] += B[r][c] * coef;
    }
}

/* ---- Newton iteration without a stored Jacobian -------------------------------------------------------------------
 * State per sample: the 12 chain-order joints and their cached sin / cos, held in SHARED memory (column per thread,
 * conflict free) so that the joint loops can stay rolled: one copy of the per-joint code instead of twelve keeps the
 * kernel inside the instruction cache and frees ~70 registers.  One iteration = two walks along the chain:
 *   walk 1: sin / cos of every joint (the only transcendental site), FK, and A0 = J0 J0^T accumulated column by column
 *           (J0 = Jacobian referred to the base origin; a column is known as soon as its joint frame is reached:
 *           col0_j = [o_j x a_j ; a_j]);
 *   solve : KDL's tip-referred J = X J0, X = [[I, -[p]x],[0, I]], so  pinv(J) tw = J0^T A0^-1 (X^-1 tw)  whenever no
 *           singular value is truncated (ChainIkSolverVel_pinv drops sigma < 1e-5).  A0 z = X^-1 tw by Cholesky;
 *           regularity is PROVEN by  sigma_min(J) >= sigma_min(J0) / ||X^-1|| >= 1 / (sqrt(trace(A0^-1)) (1 + |p|)) > 1.5e-5,
 *           otherwise the exact truncated SVD step runs (needs sigma_min < ~1e-4: essentially never for random seeds);
 *   walk 2: the columns again from the cached sin / cos, dq_j = col0_j . z, q_j += dq_j. */
#define CKC_GD_BLOCK 128
/* per-joint constants of the chain in __constant__ memory: the rolled joint loops index them with a warp-uniform
 * runtime index (an indexed kernel-parameter struct would be copied to local memory).  [env-1][segment][xyz];
 * kdl_class.cpp:11-12 (integer-division l5 = lee = 66) and :32-45; segment 6 carries the rod: 2*lee + L. */
__constant__ int c_kdl_axis[13] = { -1, 2, 1, 1, 0, 1, 0, 0, 1, 0, 1, 1, 2 };
__constant__ double c_kdl_tip[2][13][3] = {
    { { 0, 0, 166 }, { 0, 0, 124 }, { 0, 0, 270 }, { 149.6, 0, 70 }, { 152.4, 0, 0 }, { 66, 0, 0 }, { 2 * 66 + 300.0, 0, 0 },
      { 66, 0, 0 }, { 152.4, 0, 0 }, { 149.6, 0, -70 }, { 0, 0, -270 }, { 0, 0, -124 }, { 0, 0, -166 } },
    { { 0, 0, 166 }, { 0, 0, 124 }, { 0, 0, 270 }, { 149.6, 0, 70 }, { 152.4, 0, 0 }, { 66, 0, 0 }, { 2 * 66 + 500.0, 0, 0 },
      { 66, 0, 0 }, { 152.4, 0, 0 }, { 149.6, 0, -70 }, { 0, 0, -270 }, { 0, 0, -124 }, { 0, 0, -166 } } };
/* sin / cos for the Newton loop with every coefficient a constant-bank operand of its FMA (libdevice's sincos materialises
 * its 64-bit coefficients with two moves each inside the joint loop).  Cody-Waite reduction by pi/2 in two FMAs (exact
 * product, |x| < 1e4), fdlibm minimax kernels on [-pi/4, pi/4] (< 1 ulp); larger arguments take the library path. */
__constant__ double c_sc[16] = {
    6.36619772367581382433e-01,      /* 0: 2/pi */
    1.57079632679489655800e+00,      /* 1: pi/2 high */
    6.12323399573676603587e-17,      /* 2: pi/2 low */
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,        /* 3..8: S1..S6 */
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,         /* 9..14: C1..C6 */
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11, 0.0 };

/* the same chain packed for the two rolled walks: joint j = segment j + 1; its tip is (x, 0, z) in every segment (no y
 * component anywhere in kdl_class.cpp:32-45), so the position update needs two FMAs per coordinate */
struct gd_joint { double tx, tz; int axis, pad; };
__constant__ gd_joint c_gd_joint[2][12] = {
    { { 0, 124, 2, 0 }, { 0, 270, 1, 0 }, { 149.6, 70, 1, 0 }, { 152.4, 0, 0, 0 }, { 66, 0, 1, 0 }, { 2 * 66 + 300.0, 0, 0, 0 },
      { 66, 0, 0, 0 }, { 152.4, 0, 1, 0 }, { 149.6, -70, 0, 0 }, { 0, -270, 1, 0 }, { 0, -124, 1, 0 }, { 0, -166, 2, 0 } },
    { { 0, 124, 2, 0 }, { 0, 270, 1, 0 }, { 149.6, 70, 1, 0 }, { 152.4, 0, 0, 0 }, { 66, 0, 1, 0 }, { 2 * 66 + 500.0, 0, 0, 0 },
      { 66, 0, 0, 0 }, { 152.4, 0, 1, 0 }, { 149.6, -70, 0, 0 }, { 0, -270, 1, 0 }, { 0, -124, 1, 0 }, { 0, -166, 2, 0 } } };

__device__ __noinline__ double2 sincos_lib(double x) { double2 r; sincos(x, &r.x, &r.y); return r; }     /* cold: |x| >= 1e4 */

__device__ __forceinline__ void sincos_fast(double x, double* sn, double* cs) {
    if (!(fabs(x) < 1.0e4)) { const double2 r = sincos_lib(x); *sn = r.x; *cs = r.y; return; }
    const double kd = rint(x * c_sc[0]);
    const int k = (int)kd;
    double r = fma(-kd, c_sc[1], x);
    r = fma(-kd, c_sc[2], r);
    const double z = r * r;
    double ps = fma(z, c_sc[8], c_sc[7]);
    ps = fma(z, ps, c_sc[6]); ps = fma(z, ps, c_sc[5]); ps = fma(z, ps, c_sc[4]); ps = fma(z, ps, c_sc[3]);
    const double s0 = fma(r * z, ps, r);                               /* r + r^3 (S1 + z (S2 + ...)) */
    double pc = fma(z, c_sc[14], c_sc[13]);
    pc = fma(z, pc, c_sc[12]); pc = fma(z, pc, c_sc[11]); pc = fma(z, pc, c_sc[10]); pc = fma(z, pc, c_sc[9]);
    const double c0 = fma(z * z, pc, fma(-0.5, z, 1.0));               /* 1 - z/2 + z^2 (C1 + z (C2 + ...)) */
    const double ss = (k & 1) ? c0 : s0, cc = (k & 1) ? s0 : c0;
    *sn = (k & 2) ? -ss : ss;
    *cs = ((k + 1) & 2) ? -cc : cc;
}

struct gd_ref {                     /* view of one thread's column of the block's shared state array [36][CKC_GD_BLOCK] */
    double* b;
    __device__ __forceinline__ double& qc(int j) const { return b[j * CKC_GD_BLOCK]; }
    __device__ __forceinline__ double& cs(int j) const { return b[(12 + j) * CKC_GD_BLOCK]; }
    __device__ __forceinline__ double& sn(int j) const { return b[(24 + j) * CKC_GD_BLOCK]; }
};

__device__ __forceinline__ void gd_init(const gd_ref& s, const double* q) {
#pragma unroll
    for (int i = 0; i < 6; i++) { s.qc(i) = q[i]; s.qc(6 + i) = q[11 - i]; }     /* kdl_class.cpp:73-78 */
    s.qc(11) = -q[6];
}

/* rare path: tip-referred Jacobian in local memory + exact truncated pseudo inverse */
__device__ __noinline__ void gd_singular_step(const ckc_kin_dev& k, const gd_ref& s, const double* tw) {
    double qc[12], R[9], p[3], J0[6][12], J[6][12], dq[12];
    for (int c = 0; c < 12; c++) qc[c] = s.qc(c);
    chain_fk<true>(k, qc, R, p, J0);
    for (int c = 0; c < 12; c++) {
        const double w0 = J0[3][c], w1 = J0[4][c], w2 = J0[5][c];
        J[0][c] = J0[0][c] + (w1 * p[2] - w2 * p[1]);
        J[1][c] = J0[1][c] + (w2 * p[0] - w0 * p[2]);
        J[2][c] = J0[2][c] + (w0 * p[1] - w1 * p[0]);
        J[3][c] = w0; J[4][c] = w1; J[5][c] = w2;
    }
    pinv_step_svd(J, tw, 1e-5, dq);
    for (int c = 0; c < 12; c++) s.qc(c) = qc[c] + dq[c];
}

__device__ __forceinline__ void rot_by_axis(int ax, double* R, double c, double s) {     /* ax is warp-uniform */
    if (ax == 0) post_rot_x(R, c, s); else if (ax == 1) post_rot_y(R, c, s); else post_rot_z(R, c, s);
}
template <int AX> __device__ __forceinline__ void post_rot(double* R, double c, double s) {
    if (AX == 0) post_rot_x(R, c, s); else if (AX == 1) post_rot_y(R, c, s); else post_rot_z(R, c, s);
}
/* walk 1, joint about the compile-time axis AX: column of J0 (the axis is column AX of R, which the joint's own rotation
 * leaves unchanged), rank-1 update of A0 = J0 J0^T, then the joint rotation */
template <int AX> __device__ __forceinline__ void gd_walk1_joint(double* R, const double* p, double (*A)[6], double c, double s) {
    double col[6];
    col[3] = R[AX]; col[4] = R[3 + AX]; col[5] = R[6 + AX];
    col[0] = p[1] * col[5] - p[2] * col[4];
    col[1] = p[2] * col[3] - p[0] * col[5];
    col[2] = p[0] * col[4] - p[1] * col[3];
    /* explicit fma: with `+= col*col` the compiler merges the three arms' tails into 21 DMUL per arm + 21 shared DADD */
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) A[a][b] = fma(col[a], col[b], A[a][b]);
    post_rot<AX>(R, c, s);
}
/* walk 2: dq = col0 . z = a . (z_w + z_v x o), w = z_w + z_v x o precomputed by the caller */
template <int AX> __device__ __forceinline__ double gd_walk2_joint(double* R, const double* w, double c, double s, bool rotate) {
    const double dq = R[AX] * w[0] + R[3 + AX] * w[1] + R[6 + AX] * w[2];
    if (rotate) post_rot<AX>(R, c, s);
    return dq;
}

/* one ChainIkSolverPos_NR iteration; returns Equal(delta_twist, 0, 1e-5) evaluated on the twist BEFORE the update */
__device__ __forceinline__ bool gd_iterate(const ckc_kin_dev& k, const gd_ref& s) {
    const int e = k.env_index;
    double R[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    double p[3] = { c_kdl_tip[e][0][0], c_kdl_tip[e][0][1], c_kdl_tip[e][0][2] };  /* segment 0: Joint::None */
    double A[6][6];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) A[i][j] = 0;
    /* ---- walk 1 ---- */
    const gd_joint* jt = c_gd_joint[e];
#pragma unroll 1
    for (int j = 0; j < 12; j++) {
        double sj, cj;
        sincos_fast(s.qc(j), &sj, &cj);
        s.cs(j) = cj; s.sn(j) = sj;
        const int ax = jt[j].axis;                     /* warp-uniform: one three-way branch per joint */
        if (ax == 0) gd_walk1_joint<0>(R, p, A, cj, sj);
        else if (ax == 1) gd_walk1_joint<1>(R, p, A, cj, sj);
        else gd_walk1_joint<2>(R, p, A, cj, sj);
        const double tx = jt[j].tx, tz = jt[j].tz;
        p[0] = fma(R[0], tx, fma(R[2], tz, p[0]));
        p[1] = fma(R[3], tx, fma(R[5], tz, p[1]));
        p[2] = fma(R[6], tx, fma(R[8], tz, p[2]));
    }
    /* ---- delta_twist = diff(f, target): vel = (D,0,0) - p ; rot = R * GetRot(R^T) ---- */
    double tw[6];
    tw[0] = k.D - p[0]; tw[1] = -p[1]; tw[2] = -p[2];
    {
        double Rt[9], rv[3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) Rt[3 * a + b] = R[3 * b + a];
        kdl_rotvec(Rt, rv);
#pragma unroll
        for (int a = 0; a < 3; a++) tw[3 + a] = R[3 * a] * rv[0] + R[3 * a + 1] * rv[1] + R[3 * a + 2] * rv[2];
    }
    bool eq = true;
#pragma unroll
    for (int i = 0; i < 6; i++) eq = eq && (fabs(tw[i]) < 1e-5);
    /* ---- Cholesky of A0 (lower, in place; reciprocal pivots kept) ---- */
    bool regular = true;
    double inv[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double v = A[i][j];
#pragma unroll
            for (int kk = 0; kk < j; kk++) v -= A[i][kk] * A[j][kk];
            if (i == j) { if (!(v > 1e-300)) { regular = false; v = 1; } const double r = rsqrt(v); inv[i] = r; A[i][i] = v * r; }
            else A[i][j] = v * inv[j];
        }
    }
    if (regular) {
        /* trace(A0^-1) = ||L^-1||_F^2, one column of L^-1 at a time */
        double fro = 0;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            double Lc[6];
#pragma unroll
            for (int r = c; r < 6; r++) {
                double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
                for (int kk = c; kk < r; kk++) v -= A[r][kk] * Lc[kk];
                Lc[r] = v * inv[r];
                fro += Lc[r] * Lc[r];
            }
        }
        const double reach = 1.0 + sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        regular = fro * (1.5e-5 * 1.5e-5) * reach * reach < 1.0;
    }
    if (!regular) { gd_singular_step(k, s, tw); return eq; }
    /* ---- z = A0^-1 X^-1 tw,  X^-1 tw = [tw_v + p x tw_w ; tw_w] ---- */
    double z[6];
    z[0] = tw[0] + (p[1] * tw[5] - p[2] * tw[4]);
    z[1] = tw[1] + (p[2] * tw[3] - p[0] * tw[5]);
    z[2] = tw[2] + (p[0] * tw[4] - p[1] * tw[3]);
    z[3] = tw[3]; z[4] = tw[4]; z[5] = tw[5];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double v = z[i];
#pragma unroll
        for (int kk = 0; kk < i; kk++) v -= A[i][kk] * z[kk];
        z[i] = v * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double v = z[i];
#pragma unroll
        for (int kk = i + 1; kk < 6; kk++) v -= A[kk][i] * z[kk];
        z[i] = v * inv[i];
    }
    /* ---- walk 2: dq_j = col0_j . z with the frames of the configuration the twist was measured at ---- */
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    p[0] = c_kdl_tip[e][0][0]; p[1] = c_kdl_tip[e][0][1]; p[2] = c_kdl_tip[e][0][2];
#pragma unroll 1
    for (int j = 0; j < 12; j++) {
        double w[3];
        w[0] = fma(z[1], p[2], fma(-z[2], p[1], z[3]));
        w[1] = fma(z[2], p[0], fma(-z[0], p[2], z[4]));
        w[2] = fma(z[0], p[1], fma(-z[1], p[0], z[5]));
        const int ax = jt[j].axis;
        const double cj = s.cs(j), sj = s.sn(j);
        double dq;
        if (ax == 0) dq = gd_walk2_joint<0>(R, w, cj, sj, j < 11);
        else if (ax == 1) dq = gd_walk2_joint<1>(R, w, cj, sj, j < 11);
        else dq = gd_walk2_joint<2>(R, w, cj, sj, j < 11);
        s.qc(j) += dq;
        const double tx = jt[j].tx, tz = jt[j].tz;
        p[0] = fma(R[0], tx, fma(R[2], tz, p[0]));
        p[1] = fma(R[3], tx, fma(R[5], tz, p[1]));
        p[2] = fma(R[6], tx, fma(R[8], tz, p[2]));
    }
    return eq;
}

/* ---- the same iteration with the chain unrolled -------------------------------------------------------------------
 * The axis sequence Z,Y,Y,X,Y,X | X,Y,X,Y,Y,Z and the segment tips of kdl_class.cpp:32-45 are compile-time facts: with the two
 * walks fully unrolled every joint's axis arm is selected statically (no three-way branch, no constant-table loads, no loop
 * counters), tip components that are zero cost nothing (14 of the 24 are non-zero), and the |q| >= 1e4 library fallback of the
 * sin / cos moves out of the walk (a flag; the rolled gd_iterate above is the cold path that redoes such an iteration).
 * kMerge: joints 5 and 6 (chain order) are both about X with a pure X translation between them (the rod), i.e. COLLINEAR:
 * Rx(q5) Tx Rx(q6) = Tx Rx(q5 + q6), their Jacobian columns coincide, and the minimum-norm step gives both the same dq.  The
 * merged form evaluates one sin / cos of q5 + q6, one rotation and one column with weight 2 in J0 J0^T (rounding-level
 * differences against the two-joint evaluation, 1e-16 relative). */
struct gd_chain_u {
    /* tip (x, z) after each joint; joint 5 carries the rod: x = 2 * 66 + L, passed at run time */
    static constexpr double tx[12] = { 0, 0, 149.6, 152.4, 66, -1 /* rod */, 66, 152.4, 149.6, 0, 0, 0 };
    static constexpr double tz[12] = { 124, 270, 70, 0, 0, 0, 0, 0, -70, -270, -124, -166 };
};

/* sin / cos for the link frames of k_collide: Cody-Waite + fdlibm kernels as above (< 1 ulp), library call only for |x| >= 1e4 */
__device__ __forceinline__ void sincos_compact(double x, double* sn, double* cs) { sincos_fast(x, sn, cs); }

/* sin / cos without the large-argument branch: `bad` collects |x| >= 1e4 (or NaN), the caller redoes the iteration on the cold path */
__device__ __forceinline__ void sincos_nobranch(double x, double* sn, double* cs, bool& bad) {
    bad = bad || !(fabs(x) < 1.0e4);
    const double kd = rint(x * c_sc[0]);
    const int k = (int)kd;
    double r = fma(-kd, c_sc[1], x);
    r = fma(-kd, c_sc[2], r);
    const double z = r * r;
    double ps = fma(z, c_sc[8], c_sc[7]);
    ps = fma(z, ps, c_sc[6]); ps = fma(z, ps, c_sc[5]); ps = fma(z, ps, c_sc[4]); ps = fma(z, ps, c_sc[3]);
    const double s0 = fma(r * z, ps, r);
    double pc = fma(z, c_sc[14], c_sc[13]);
    pc = fma(z, pc, c_sc[12]); pc = fma(z, pc, c_sc[11]); pc = fma(z, pc, c_sc[10]); pc = fma(z, pc, c_sc[9]);
    const double c0 = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double ss = (k & 1) ? c0 : s0, cc = (k & 1) ? s0 : c0;
    /* sign flips on the high word (one integer instruction each instead of an FP64 negation + two selects) */
    *sn = __hiloint2double(__double2hiint(ss) ^ ((k & 2) << 30), __double2loint(ss));
    *cs = __hiloint2double(__double2hiint(cc) ^ (((k + 1) & 2) << 30), __double2loint(cc));
}

template <int AX> __device__ __forceinline__ void gd_col(const double* R, const double* p, double* col) {
    col[3] = R[AX]; col[4] = R[3 + AX]; col[5] = R[6 + AX];
    col[0] = p[1] * col[5] - p[2] * col[4];
    col[1] = p[2] * col[3] - p[0] * col[5];
    col[2] = p[0] * col[4] - p[1] * col[3];
}
template <int J> __device__ __forceinline__ void gd_tip_u(const double* R, double* p, double rodx) {
    constexpr double tz = gd_chain_u::tz[J];
    const double tx = J == 5 ? rodx : gd_chain_u::tx[J];
    if (J == 5 || gd_chain_u::tx[J] != 0.0) { p[0] = fma(R[0], tx, p[0]); p[1] = fma(R[3], tx, p[1]); p[2] = fma(R[6], tx, p[2]); }
    if (tz != 0.0) { p[0] = fma(R[2], tz, p[0]); p[1] = fma(R[5], tz, p[1]); p[2] = fma(R[8], tz, p[2]); }
}

template <bool kMerge> __device__ __noinline__ bool gd_iterate_cold(const ckc_kin_dev& k, const gd_ref& s) { return gd_iterate(k, s); }

template <bool kMerge>
__device__ __forceinline__ bool gd_iterate_u(const ckc_kin_dev& k, const gd_ref& s) {
    const double rodx = 2 * 66.0 + k.L;
    double R[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    double p[3] = { 0, 0, 166.0 };                 /* segment 0: Joint::None, Tz(b) */
    double A[6][6];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) A[i][j] = 0;
    bool bad = false;
    constexpr int kAx[12] = { 2, 1, 1, 0, 1, 0, 0, 1, 0, 1, 1, 2 };
    /* ---- walk 1 ---- */
#pragma unroll
    for (int j = 0; j < 12; j++) {
        if (kMerge && j == 6) continue;            /* folded into joint 5 */
        double sj, cj;
        const double qj = (kMerge && j == 5) ? s.qc(5) + s.qc(6) : s.qc(j);
        sincos_nobranch(qj, &sj, &cj, bad);
        s.cs(j) = cj; s.sn(j) = sj;
        double col[6];
        const int ax = kAx[j];
        if (ax == 0) gd_col<0>(R, p, col); else if (ax == 1) gd_col<1>(R, p, col); else gd_col<2>(R, p, col);
        if (kMerge && j == 5) {
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const double c2 = col[a] + col[a];
#pragma unroll
                for (int b = 0; b <= a; b++) A[a][b] = fma(c2, col[b], A[a][b]);
            }
        } else {
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) A[a][b] = fma(col[a], col[b], A[a][b]);
        }
        if (ax == 0) post_rot_x(R, cj, sj); else if (ax == 1) post_rot_y(R, cj, sj); else post_rot_z(R, cj, sj);
        if (j == 0) gd_tip_u<0>(R, p, rodx); else if (j == 1) gd_tip_u<1>(R, p, rodx); else if (j == 2) gd_tip_u<2>(R, p, rodx);
        else if (j == 3) gd_tip_u<3>(R, p, rodx); else if (j == 4) gd_tip_u<4>(R, p, rodx); else if (j == 5) gd_tip_u<5>(R, p, rodx);
        else if (j == 6) gd_tip_u<6>(R, p, rodx); else if (j == 7) gd_tip_u<7>(R, p, rodx); else if (j == 8) gd_tip_u<8>(R, p, rodx);
        else if (j == 9) gd_tip_u<9>(R, p, rodx); else if (j == 10) gd_tip_u<10>(R, p, rodx); else gd_tip_u<11>(R, p, rodx);
        if (kMerge && j == 5) gd_tip_u<6>(R, p, rodx);      /* joint 6's tip follows at once: Tx commutes with Rx */
    }
    if (bad) return gd_iterate_cold<kMerge>(k, s);          /* |q| >= 1e4: library sin / cos (state untouched so far) */
    /* ---- delta_twist = diff(f, target): vel = (D,0,0) - p ; rot = R * GetRot(R^T) ---- */
    double tw[6];
    tw[0] = k.D - p[0]; tw[1] = -p[1]; tw[2] = -p[2];
    {
        double Rt[9], rv[3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) Rt[3 * a + b] = R[3 * b + a];
        kdl_rotvec(Rt, rv);
#pragma unroll
        for (int a = 0; a < 3; a++) tw[3 + a] = R[3 * a] * rv[0] + R[3 * a + 1] * rv[1] + R[3 * a + 2] * rv[2];
    }
    bool eq = true;
#pragma unroll
    for (int i = 0; i < 6; i++) eq = eq && (fabs(tw[i]) < 1e-5);
    /* ---- Cholesky of A0 (lower, in place; reciprocal pivots kept) ---- */
    bool regular = true;
    double inv[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double v = A[i][j];
#pragma unroll
            for (int kk = 0; kk < j; kk++) v -= A[i][kk] * A[j][kk];
            if (i == j) { if (!(v > 1e-300)) { regular = false; v = 1; } const double r = rsqrt(v); inv[i] = r; A[i][i] = v * r; }
            else A[i][j] = v * inv[j];
        }
    }
    if (regular) {
        double fro = 0;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            double Lc[6];
#pragma unroll
            for (int r = c; r < 6; r++) {
                double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
                for (int kk = c; kk < r; kk++) v -= A[r][kk] * Lc[kk];
                Lc[r] = v * inv[r];
                fro += Lc[r] * Lc[r];
            }
        }
        const double reach = 1.0 + sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        regular = fro * (1.5e-5 * 1.5e-5) * reach * reach < 1.0;
    }
    if (!regular) { gd_singular_step(k, s, tw); return eq; }
    double z[6];
    z[0] = tw[0] + (p[1] * tw[5] - p[2] * tw[4]);
    z[1] = tw[1] + (p[2] * tw[3] - p[0] * tw[5]);
    z[2] = tw[2] + (p[0] * tw[4] - p[1] * tw[3]);
    z[3] = tw[3]; z[4] = tw[4]; z[5] = tw[5];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double v = z[i];
#pragma unroll
        for (int kk = 0; kk < i; kk++) v -= A[i][kk] * z[kk];
        z[i] = v * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double v = z[i];
#pragma unroll
        for (int kk = i + 1; kk < 6; kk++) v -= A[kk][i] * z[kk];
        z[i] = v * inv[i];
    }
    /* ---- walk 2 ---- */
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    p[0] = 0; p[1] = 0; p[2] = 166.0;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        if (kMerge && j == 6) continue;
        double w[3];
        w[0] = fma(z[1], p[2], fma(-z[2], p[1], z[3]));
        w[1] = fma(z[2], p[0], fma(-z[0], p[2], z[4]));
        w[2] = fma(z[0], p[1], fma(-z[1], p[0], z[5]));
        const int ax = kAx[j];
        const double dq = R[ax] * w[0] + R[3 + ax] * w[1] + R[6 + ax] * w[2];
        /* volatile reads: with the walks unrolled the compiler would otherwise forward the 36 values stored in walk 1 through
         * registers across the whole solve (and spill them) instead of re-reading the shared-memory column */
        s.qc(j) = *(volatile double*)&s.qc(j) + dq;
        if (kMerge && j == 5) s.qc(6) = *(volatile double*)&s.qc(6) + dq;
        if (j < 11) {
            const double cj = *(volatile double*)&s.cs(j), sj = *(volatile double*)&s.sn(j);
            if (ax == 0) post_rot_x(R, cj, sj); else if (ax == 1) post_rot_y(R, cj, sj); else post_rot_z(R, cj, sj);
            if (j == 0) gd_tip_u<0>(R, p, rodx); else if (j == 1) gd_tip_u<1>(R, p, rodx); else if (j == 2) gd_tip_u<2>(R, p, rodx);
            else if (j == 3) gd_tip_u<3>(R, p, rodx); else if (j == 4) gd_tip_u<4>(R, p, rodx); else if (j == 5) gd_tip_u<5>(R, p, rodx);
            else if (j == 6) gd_tip_u<6>(R, p, rodx); else if (j == 7) gd_tip_u<7>(R, p, rodx); else if (j == 8) gd_tip_u<8>(R, p, rodx);
            else if (j == 9) gd_tip_u<9>(R, p, rodx); else gd_tip_u<10>(R, p, rodx);
            if (kMerge && j == 5) gd_tip_u<6>(R, p, rodx);
        }
    }
    return eq;
}

/* exact fmod(v, P) for the magnitudes a Newton iterate reaches: fma(-k, P, v) is the exactly representable remainder when
 * k = trunc(v / P) is right; the rounded quotient can be off by one next to a multiple of P, which the fix-ups repair
 * (each fix-up result is again the exactly representable true remainder).  Falls back to fmod() for huge arguments. */
__device__ __forceinline__ double fmod_exact(double v, double P) {
    if (!(fabs(v) < 1e6)) return fmod(v, P);
    const double kq = trunc(v * (1.0 / P));       /* reciprocal multiply: an off-by-one quotient is repaired below */
    double r = fma(-kq, P, v);
    if (v >= 0) { if (r < 0) r += P; else if (r >= P) r -= P; }
    else { if (r > 0) r -= P; else if (r <= -P) r += P; }
    return r;
}

/* post-processing of kdl::GD (kdl_class.cpp:110-128): zero snap, fmod / wrap with PI = 3.1416, back to user order */
__device__ __forceinline__ void gd_finish(const gd_ref& s, double* q) {
    double qc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        double v = s.qc(i);
        v = (fabs(v) < 1e-4) ? 0.0 : v;
        v = fmod_exact(v, 2 * CKC_PI_);
        if (v > CKC_PI_) v -= 2 * CKC_PI_;
        if (v < -CKC_PI_) v += 2 * CKC_PI_;
        qc[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) { q[i] = qc[i]; q[6 + i] = qc[11 - i]; }
    q[6] = -q[6];
}

/* the same post-processing on a register copy of the raw chain-order joints */
__device__ __forceinline__ void gd_finish_regs(const double* raw, double* q) {
    double qc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        double v = raw[i];
        v = (fabs(v) < 1e-4) ? 0.0 : v;
        v = fmod_exact(v, 2 * CKC_PI_);
        if (v > CKC_PI_) v -= 2 * CKC_PI_;
        if (v < -CKC_PI_) v += 2 * CKC_PI_;
        qc[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) { q[i] = qc[i]; q[6 + i] = qc[11 - i]; }
    q[6] = -q[6];
}

/* kdl::GD as a plain per-thread loop (used by the RBS frontier).  q: user order in/out.  returns converged */
__device__ __forceinline__ bool gd_project(const ckc_kin_dev& k, const gd_ref& s, double* q, int max_iter, int* iters) {
    gd_init(s, q);
    bool conv = false;
    int it = 0;
#pragma unroll 1
    for (; it < max_iter; it++) {
        if (gd_iterate(k, s)) { conv = true; it++; break; }
    }
    *iters = it;
    if (!conv) return false;
    gd_finish(s, q);
    return true;
}

/* check_relax_constraint StateValidityCheckerRX.cpp:105-135 */
__device__ __forceinline__ double rx_residual(const ckc_kin_dev& k, const double* q) {
    double qc[12], R[9], p[3];
#pragma unroll
    for (int i = 0; i < 6; i++) { qc[i] = q[i]; qc[6 + i] = q[11 - i]; }
    qc[11] = -qc[11];
    chain_fk<false>(k, qc, R, p, nullptr);
    const double dx = p[0] - k.D, dy = p[1], dz = p[2];
    const double d = dx * dx + dy * dy + dz * dz;
    const double roll = atan2(R[3], R[0]);
    const double pitch = atan2(-R[6], sqrt(R[7] * R[7] + R[8] * R[8]));
    const double yaw = atan2(R[7], R[8]);
    const double a = (yaw * yaw + pitch * pitch + roll * roll);
    const double Ka = (0.3 * 0.3) / (100.0 * 100.0);                           /* RX.h:230-233 */
    return sqrt(Ka * d + a);
}

}  // namespace ckc
#endif
