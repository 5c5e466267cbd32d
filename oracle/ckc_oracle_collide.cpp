/*
 * ckc_oracle_collide.cpp -- CPU ORACLE (test infrastructure, NOT product code): collision_state.
 * Restates validity_checkers/collisionDetection.cpp:31-494 (link frames, pair table, quirks) and the
 * semantics of the external PQP v1.3 tolerance query (closer_than_tolerance <=> exists a triangle pair with
 * TriDist <= tol; TriDist/SegPoints restated from the published PQP algorithm).  The bounding hierarchy is
 * the oracle's own (PCA-fitted OBB tree with a separating-axis lower bound): PQP's RSS tree only prunes
 * conservatively, so the boolean does not depend on the tree.  Pinned against the PQP
 * copy compiled into tests/trbspcs (oracle/refshim.cpp): per-pair flags, masks and TriDist values.
 */
#include "ckc_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

/* ------------------------------------------------------------------------------------------------ */
/* 3x3 helpers with the operand association of simulator/MatVec.h (MxM :231-264, MxV :331-346)       */
/* ------------------------------------------------------------------------------------------------ */
static void MxM(double Mr[9], const double A[9], const double B[9]) {
    double R[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[3 * i + j] = (A[3 * i + 0] * B[0 + j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]);
    memcpy(Mr, R, sizeof(R));
}
/* MxV writes Vr component by component: when Vr aliases V (collisionDetection.cpp:103,207) later components
 * read the already-overwritten ones.  Reproduced on purpose. */
static void MxV_inplace_alias(double V[3], const double M[9]) {
    V[0] = (M[0] * V[0] + M[1] * V[1] + M[2] * V[2]);
    V[1] = (M[3] * V[0] + M[4] * V[1] + M[5] * V[2]);
    V[2] = (M[6] * V[0] + M[7] * V[1] + M[8] * V[2]);
}
static void MxV(double Vr[3], const double M[9], const double V[3]) {
    double r0 = (M[0] * V[0] + M[1] * V[1] + M[2] * V[2]);
    double r1 = (M[3] * V[0] + M[4] * V[1] + M[5] * V[2]);
    double r2 = (M[6] * V[0] + M[7] * V[1] + M[8] * V[2]);
    Vr[0] = r0; Vr[1] = r1; Vr[2] = r2;
}
static void MRotZ(double M[9], double t) { double c = cos(t), s = sin(t); M[0] = c; M[3] = s; M[1] = -s; M[4] = c; M[6] = M[7] = 0; M[2] = M[5] = 0; M[8] = 1; }
static void MRotX(double M[9], double t) { double c = cos(t), s = sin(t); M[4] = c; M[7] = s; M[5] = -s; M[8] = c; M[1] = M[2] = 0; M[3] = M[6] = 0; M[0] = 1; }
static void MRotY(double M[9], double t) { double c = cos(t), s = sin(t); M[8] = c; M[2] = s; M[6] = -s; M[0] = c; M[5] = M[3] = 0; M[7] = M[1] = 0; M[4] = 1; }

static inline void v_sub(double r[3], const double a[3], const double b[3]) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v_add(double r[3], const double a[3], const double b[3]) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline double v_dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v_cross(double r[3], const double a[3], const double b[3]) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline void v_cpy(double r[3], const double a[3]) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void v_axpy(double r[3], const double a[3], const double b[3], double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }

/* ------------------------------------------------------------------------------------------------ */
/* PQP v1.3 SegPoints / TriDist (external library src/TriDist.cpp; restated, validated vs the binary)  */
/* ------------------------------------------------------------------------------------------------ */
/* closest points X on segment P+tA and Y on segment Q+uB (t,u in [0,1]); VEC = a vector along Y-X that stays
 * meaningful when the points coincide */
static void seg_points(double VEC[3], double X[3], double Y[3], const double P[3], const double A[3], const double Q[3], const double B[3]) {
    double T[3], TMP[3];
    v_sub(T, Q, P);
    const double A_dot_A = v_dot(A, A), B_dot_B = v_dot(B, B), A_dot_B = v_dot(A, B);
    const double A_dot_T = v_dot(A, T), B_dot_T = v_dot(B, T);
    const double denom = A_dot_A * B_dot_B - A_dot_B * A_dot_B;
    double t = (A_dot_T * B_dot_B - B_dot_T * A_dot_B) / denom;
    if ((t < 0) || isnan(t)) t = 0; else if (t > 1) t = 1;
    double u = (t * A_dot_B - B_dot_T) / B_dot_B;
    if ((u <= 0) || isnan(u)) {
        v_cpy(Y, Q);
        t = A_dot_T / A_dot_A;
        if ((t <= 0) || isnan(t)) { v_cpy(X, P); v_sub(VEC, Q, P); }
        else if (t >= 1) { v_add(X, P, A); v_sub(VEC, Q, X); }
        else { v_axpy(X, P, A, t); v_cross(TMP, T, A); v_cross(VEC, A, TMP); }
    } else if (u >= 1) {
        v_add(Y, Q, B);
        t = (A_dot_B + A_dot_T) / A_dot_A;
        if ((t <= 0) || isnan(t)) { v_cpy(X, P); v_sub(VEC, Y, P); }
        else if (t >= 1) { v_add(X, P, A); v_sub(VEC, Y, X); }
        else { v_axpy(X, P, A, t); v_sub(T, Y, P); v_cross(TMP, T, A); v_cross(VEC, A, TMP); }
    } else {
        v_axpy(Y, Q, B, u);
        if ((t <= 0) || isnan(t)) { v_cpy(X, P); v_cross(TMP, T, B); v_cross(VEC, B, TMP); }
        else if (t >= 1) { v_add(X, P, A); v_sub(T, Q, X); v_cross(TMP, T, B); v_cross(VEC, B, TMP); }
        else {
            v_axpy(X, P, A, t);
            v_cross(VEC, A, B);
            if (v_dot(VEC, T) < 0) { VEC[0] = -VEC[0]; VEC[1] = -VEC[1]; VEC[2] = -VEC[2]; }
        }
    }
}

extern "C" double ckco_tri_dist(double P[3], double Q[3], const double S[3][3], const double T[3][3]) {
    double Sv[3][3], Tv[3][3], VEC[3];
    v_sub(Sv[0], S[1], S[0]); v_sub(Sv[1], S[2], S[1]); v_sub(Sv[2], S[0], S[2]);
    v_sub(Tv[0], T[1], T[0]); v_sub(Tv[1], T[2], T[1]); v_sub(Tv[2], T[0], T[2]);
    double V[3], Z[3], minP[3] = {0, 0, 0}, minQ[3] = {0, 0, 0}, mindd;
    int shown_disjoint = 0;
    { double d0[3]; v_sub(d0, S[0], T[0]); mindd = v_dot(d0, d0) + 1; }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            seg_points(VEC, P, Q, S[i], Sv[i], T[j], Tv[j]);
            v_sub(V, Q, P);
            const double dd = v_dot(V, V);
            if (dd <= mindd) {
                v_cpy(minP, P); v_cpy(minQ, Q); mindd = dd;
                v_sub(Z, S[(i + 2) % 3], P); double a = v_dot(Z, VEC);
                v_sub(Z, T[(j + 2) % 3], Q); double b = v_dot(Z, VEC);
                if ((a <= 0) && (b >= 0)) return sqrt(dd);
                const double p = v_dot(V, VEC);
                if (a < 0) a = 0;
                if (b > 0) b = 0;
                if ((p - a + b) > 0) shown_disjoint = 1;
            }
        }
    }
    /* vertex-face cases */
    double Sn[3]; v_cross(Sn, Sv[0], Sv[1]);
    const double Snl = v_dot(Sn, Sn);
    if (Snl > 1e-15) {
        double Tp[3];
        v_sub(V, S[0], T[0]); Tp[0] = v_dot(V, Sn);
        v_sub(V, S[0], T[1]); Tp[1] = v_dot(V, Sn);
        v_sub(V, S[0], T[2]); Tp[2] = v_dot(V, Sn);
        int point = -1;
        if ((Tp[0] > 0) && (Tp[1] > 0) && (Tp[2] > 0)) {
            point = (Tp[0] < Tp[1]) ? 0 : 1;
            if (Tp[2] < Tp[point]) point = 2;
        } else if ((Tp[0] < 0) && (Tp[1] < 0) && (Tp[2] < 0)) {
            point = (Tp[0] > Tp[1]) ? 0 : 1;
            if (Tp[2] > Tp[point]) point = 2;
        }
        if (point >= 0) {
            shown_disjoint = 1;
            v_sub(V, T[point], S[0]); v_cross(Z, Sn, Sv[0]);
            if (v_dot(V, Z) > 0) {
                v_sub(V, T[point], S[1]); v_cross(Z, Sn, Sv[1]);
                if (v_dot(V, Z) > 0) {
                    v_sub(V, T[point], S[2]); v_cross(Z, Sn, Sv[2]);
                    if (v_dot(V, Z) > 0) {
                        v_axpy(P, T[point], Sn, Tp[point] / Snl);
                        v_cpy(Q, T[point]);
                        double d[3]; v_sub(d, P, Q);
                        return sqrt(v_dot(d, d));
                    }
                }
            }
        }
    }
    double Tn[3]; v_cross(Tn, Tv[0], Tv[1]);
    const double Tnl = v_dot(Tn, Tn);
    if (Tnl > 1e-15) {
        double Sp[3];
        v_sub(V, T[0], S[0]); Sp[0] = v_dot(V, Tn);
        v_sub(V, T[0], S[1]); Sp[1] = v_dot(V, Tn);
        v_sub(V, T[0], S[2]); Sp[2] = v_dot(V, Tn);
        int point = -1;
        if ((Sp[0] > 0) && (Sp[1] > 0) && (Sp[2] > 0)) {
            point = (Sp[0] < Sp[1]) ? 0 : 1;
            if (Sp[2] < Sp[point]) point = 2;
        } else if ((Sp[0] < 0) && (Sp[1] < 0) && (Sp[2] < 0)) {
            point = (Sp[0] > Sp[1]) ? 0 : 1;
            if (Sp[2] > Sp[point]) point = 2;
        }
        if (point >= 0) {
            shown_disjoint = 1;
            v_sub(V, S[point], T[0]); v_cross(Z, Tn, Tv[0]);
            if (v_dot(V, Z) > 0) {
                v_sub(V, S[point], T[1]); v_cross(Z, Tn, Tv[1]);
                if (v_dot(V, Z) > 0) {
                    v_sub(V, S[point], T[2]); v_cross(Z, Tn, Tv[2]);
                    if (v_dot(V, Z) > 0) {
                        v_cpy(P, S[point]);
                        v_axpy(Q, S[point], Tn, Sp[point] / Tnl);
                        double d[3]; v_sub(d, P, Q);
                        return sqrt(v_dot(d, d));
                    }
                }
            }
        }
    }
    if (shown_disjoint) { v_cpy(P, minP); v_cpy(Q, minQ); return sqrt(mindd); }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* meshes + the oracle's own OBB hierarchy                                                            */
/* ------------------------------------------------------------------------------------------------ */
struct obb_node {
    double c[3];        /* centre, mesh frame */
    double ax[9];       /* rows = box axes, mesh frame */
    double h[3];        /* half extents */
    int left, right;    /* children, or -1 */
    int tri;            /* leaf: triangle index */
};
struct mesh {
    std::vector<double> tris;       /* ntris * 9 */
    std::vector<obb_node> nodes;    /* root = 0 */
    int ntris() const { return (int)(tris.size() / 9); }
};
struct ckco_world {
    int env;
    double D, L;
    mesh meshes[CKCO_NUM_MESHES];
    std::vector<ckco_pair> pairs;
    double tol;
};

static void jacobi_eig3(double A[9], double V[9]) {     /* symmetric 3x3 -> eigenvectors in columns of V */
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (off < 1e-30) break;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 3; q++) {
                double apq = A[3 * p + q];
                if (fabs(apq) < 1e-300) continue;
                double app = A[3 * p + p], aqq = A[3 * q + q];
                double theta = (aqq - app) / (2 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq; A[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk; A[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq; V[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
}

static void fit_obb(const mesh& m, const std::vector<int>& idx, obb_node& n) {
    double mean[3] = {0, 0, 0};
    const int nv = (int)idx.size() * 3;
    for (int t : idx) for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) mean[k] += m.tris[9 * t + 3 * v + k];
    for (int k = 0; k < 3; k++) mean[k] /= nv;
    double C[9] = {0};
    for (int t : idx) for (int v = 0; v < 3; v++) {
        double d[3]; for (int k = 0; k < 3; k++) d[k] = m.tris[9 * t + 3 * v + k] - mean[k];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[3 * i + j] += d[i] * d[j];
    }
    double V[9]; jacobi_eig3(C, V);
    /* orthonormal right-handed axes = columns of V */
    double a0[3] = { V[0], V[3], V[6] }, a1[3] = { V[1], V[4], V[7] }, a2[3];
    double l0 = sqrt(v_dot(a0, a0)); for (int k = 0; k < 3; k++) a0[k] /= l0;
    double d01 = v_dot(a0, a1); for (int k = 0; k < 3; k++) a1[k] -= d01 * a0[k];
    double l1 = sqrt(v_dot(a1, a1));
    if (l1 < 1e-12) { /* degenerate: pick any perpendicular */
        double e[3] = { fabs(a0[0]) < 0.9 ? 1.0 : 0.0, fabs(a0[0]) < 0.9 ? 0.0 : 1.0, 0 };
        v_cross(a1, a0, e); l1 = sqrt(v_dot(a1, a1));
    }
    for (int k = 0; k < 3; k++) a1[k] /= l1;
    v_cross(a2, a0, a1);
    const double* ax[3] = { a0, a1, a2 };
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int t : idx) for (int v = 0; v < 3; v++)
        for (int i = 0; i < 3; i++) {
            double pr = v_dot(&m.tris[9 * t + 3 * v], ax[i]);
            lo[i] = std::min(lo[i], pr); hi[i] = std::max(hi[i], pr);
        }
    for (int i = 0; i < 3; i++) { n.h[i] = 0.5 * (hi[i] - lo[i]); for (int k = 0; k < 3; k++) n.ax[3 * i + k] = ax[i][k]; }
    for (int k = 0; k < 3; k++) n.c[k] = 0.5 * (lo[0] + hi[0]) * a0[k] + 0.5 * (lo[1] + hi[1]) * a1[k] + 0.5 * (lo[2] + hi[2]) * a2[k];
}

static int build_node(mesh& m, std::vector<int>& idx) {
    obb_node n; n.left = n.right = -1; n.tri = -1;
    fit_obb(m, idx, n);
    const int me = (int)m.nodes.size();
    m.nodes.push_back(n);
    if (idx.size() == 1) { m.nodes[me].tri = idx[0]; return me; }
    /* split along the longest box axis at the median triangle centroid */
    int axis = 0; if (n.h[1] > n.h[axis]) axis = 1; if (n.h[2] > n.h[axis]) axis = 2;
    const double* a = &m.nodes[me].ax[3 * axis];
    double aa[3] = { a[0], a[1], a[2] };
    std::sort(idx.begin(), idx.end(), [&](int x, int y) {
        double cx = 0, cy = 0;
        for (int v = 0; v < 3; v++) { cx += v_dot(&m.tris[9 * x + 3 * v], aa); cy += v_dot(&m.tris[9 * y + 3 * v], aa); }
        if (cx != cy) return cx < cy;
        return x < y;
    });
    const size_t half = idx.size() / 2;
    std::vector<int> L(idx.begin(), idx.begin() + half), R(idx.begin() + half, idx.end());
    int l = build_node(m, L);
    int r = build_node(m, R);
    m.nodes[me].left = l; m.nodes[me].right = r;
    return me;
}

static void build_mesh(mesh& m) {
    m.nodes.clear();
    if (m.ntris() == 0) return;
    std::vector<int> idx(m.ntris());
    for (int i = 0; i < m.ntris(); i++) idx[i] = i;
    build_node(m, idx);
}

static const char* const k_mesh_files[CKCO_NUM_MESHES] = {       /* collisionDetection.cpp:502-923 */
    "Base_r.tris", "Link1_r.tris", "Link2_r.tris", "Link3_r.tris", "Link4_r2.tris", "Link5_r.tris",
    "Link6.tris", "EE_r.tris", "table.tris", "obs.tris", "cone.tris", NULL };

static int read_tris(const char* path, std::vector<double>& out) {   /* format: collisionDetection.cpp:502-518 */
    FILE* fp = fopen(path, "r");
    if (!fp) return -1;
    int ntris = 0;
    if (fscanf(fp, "%d", &ntris) != 1) { fclose(fp); return -1; }
    out.resize((size_t)ntris * 9);
    for (int i = 0; i < ntris * 9; i++)
        if (fscanf(fp, "%lf", &out[i]) != 1) { fclose(fp); return -1; }
    fclose(fp);
    return ntris;
}

/* the rod: setP (StateValidityCheckerPCS.h:150-158) + collision_state :50-60: points (20 i,0,0), i=0..L/20,
 * grouped three by three into degenerate triangles */
static void make_rod(mesh& m, double L) {
    const int dl = 20;
    const int n = (int)(L / dl);
    std::vector<double> pts;
    for (int i = 0; i <= n; i++) { pts.push_back((double)i * dl); pts.push_back(0); pts.push_back(0); }
    const int ntri = (n + 1) / 3;
    m.tris.assign(pts.begin(), pts.begin() + (size_t)ntri * 9);
}

static void add_pair(ckco_world* w, int ma, int fa, int mb, int fb) { ckco_pair p = { ma, fa, mb, fb }; w->pairs.push_back(p); }

static void build_pairs(ckco_world* w) {
    enum { B = CKCO_M_BASE, L1 = CKCO_M_LINK1, L2 = CKCO_M_LINK2, L3 = CKCO_M_LINK3, L4 = CKCO_M_LINK4, L5 = CKCO_M_LINK5,
           L6 = CKCO_M_LINK6, EE = CKCO_M_EE, TB = CKCO_M_TABLE, ROD = CKCO_M_ROD };
    const int linkmesh[8] = { B, L1, L2, L3, L4, L5, L6, EE };
    /* frames: robot 1 = 0..7 (base, link1, link2, link3, link4, link5, link6, EE/rod), robot 2 = 8..15 */
    for (int r = 0; r < 2; r++) {                 /* res[0..12], res[13..25]  (:267-295) */
        const int o = 8 * r;
        for (int b = 4; b <= 6; b++) for (int a = 0; a <= 2; a++) add_pair(w, linkmesh[a], o + a, linkmesh[b], o + b);
        add_pair(w, B, o + 0, L3, o + 3);
        for (int a = 0; a <= 2; a++) add_pair(w, linkmesh[a], o + a, EE, o + 7);
    }
    for (int a = 3; a <= 6; a++) for (int b = 3; b <= 6; b++) add_pair(w, linkmesh[a], a, linkmesh[b], 8 + b);   /* res[26..41] :298-313 */
    for (int a = 3; a <= 6; a++) add_pair(w, linkmesh[a], a, EE, 15);                                             /* res[42..45] */
    for (int b = 3; b <= 6; b++) add_pair(w, EE, 7, linkmesh[b], 8 + b);                                          /* res[46..49] */
    for (int r = 0; r < 2; r++) for (int b = 0; b <= 6; b++) add_pair(w, ROD, 7, linkmesh[b], 8 * r + b);         /* res[50..63] :326-339 */
    for (int r = 0; r < 2; r++) for (int b = 2; b <= 7; b++) add_pair(w, TB, 0, linkmesh[b], 8 * r + b);          /* res[64..75] :342-353 */
    add_pair(w, TB, 0, ROD, 7);                                                                                    /* res[76] :356 */
    const int nobs = (w->env == 1) ? 3 : 2;
    const int obsmesh = (w->env == 1) ? CKCO_M_OBS : CKCO_M_CONE;
    for (int k = 0; k < nobs; k++) {               /* :372-420 (env 1), :443-472 (env 2) */
        const int fo = 18 + k;
        add_pair(w, obsmesh, fo, ROD, 7);
        /* quirk 3: obs1 x link2 uses (R22, T2_t) = frame 16; obs2/obs3 use the true robot-1 link2 frame */
        add_pair(w, obsmesh, fo, L2, (k == 0) ? 16 : 2);
        for (int b = 3; b <= 7; b++) add_pair(w, obsmesh, fo, linkmesh[b], b);
        /* quirk 3: obs2/obs3 x link22 use (R2, T2_t2) = frame 17; obs1 uses the true robot-2 link2 frame */
        add_pair(w, obsmesh, fo, L2, (k == 0) ? 10 : 17);
        for (int b = 3; b <= 7; b++) add_pair(w, obsmesh, fo, linkmesh[b], 8 + b);
    }
}

static ckco_world* world_finish(ckco_world* w) {
    make_rod(w->meshes[CKCO_M_ROD], w->L);
    for (int i = 0; i < CKCO_NUM_MESHES; i++) build_mesh(w->meshes[i]);
    build_pairs(w);
    w->tol = 8.0;                                  /* collisionDetection.cpp:263 */
    return w;
}

extern "C" ckco_world* ckco_world_create_from_tris(int env, const char* tris_dir) {
    ckco_world* w = new ckco_world();
    w->env = env; w->D = (env == 1) ? 900. : 1200.; w->L = (env == 1) ? 300. : 500.;
    for (int i = 0; i < CKCO_NUM_MESHES; i++) {
        if (!k_mesh_files[i]) continue;
        char path[1024];
        snprintf(path, sizeof(path), "%s/%s", tris_dir, k_mesh_files[i]);
        if (read_tris(path, w->meshes[i].tris) < 0) { fprintf(stderr, "ckc_oracle: cannot read %s\n", path); delete w; return NULL; }
    }
    return world_finish(w);
}

/* packed mesh file (tools/pack_meshes.py): "CKCMESH1", int32 nmesh, then per mesh: int32 id, int32 ntris, ntris*9 f64 */
extern "C" ckco_world* ckco_world_create_from_pack(int env, const char* pack_path) {
    FILE* fp = fopen(pack_path, "rb");
    if (!fp) { fprintf(stderr, "ckc_oracle: cannot open %s\n", pack_path); return NULL; }
    char magic[8]; int32_t nmesh = 0;
    if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "CKCMESH1", 8) != 0 || fread(&nmesh, 4, 1, fp) != 1) { fclose(fp); return NULL; }
    ckco_world* w = new ckco_world();
    w->env = env; w->D = (env == 1) ? 900. : 1200.; w->L = (env == 1) ? 300. : 500.;
    for (int m = 0; m < nmesh; m++) {
        int32_t id, ntris;
        if (fread(&id, 4, 1, fp) != 1 || fread(&ntris, 4, 1, fp) != 1 || id < 0 || id >= CKCO_NUM_MESHES) { fclose(fp); delete w; return NULL; }
        w->meshes[id].tris.resize((size_t)ntris * 9);
        if (fread(w->meshes[id].tris.data(), 8, (size_t)ntris * 9, fp) != (size_t)ntris * 9) { fclose(fp); delete w; return NULL; }
    }
    fclose(fp);
    return world_finish(w);
}

extern "C" void ckco_world_destroy(ckco_world* w) { delete w; }
extern "C" int ckco_world_num_pairs(const ckco_world* w) { return (int)w->pairs.size(); }
extern "C" int ckco_world_mesh_ntris(const ckco_world* w, int m) { return w->meshes[m].ntris(); }
extern "C" const double* ckco_world_mesh_tris(const ckco_world* w, int m) { return w->meshes[m].tris.data(); }
extern "C" const ckco_pair* ckco_world_pairs(const ckco_world* w) { return w->pairs.data(); }

/* ------------------------------------------------------------------------------------------------ */
/* link frames: collisionDetection.cpp:72-260                                                         */
/* ------------------------------------------------------------------------------------------------ */
static void robot_frames(double base_rot_half, double offX, const double q[6], double fr[8][12]) {
    double M[9], R[7][9], T[8][3], v[3], tmp[3];
    /* base: M0 = Rz(angle), R0 = M0*M0  (:72-73 robot 1 with angle 0; :177-178 robot 2 with 3.14159265/2) */
    MRotZ(M, base_rot_half); MxM(R[0], M, M);
    MRotZ(M, q[0]); MxM(R[1], R[0], M);            /* :75-76 */
    MRotY(M, q[1]); MxM(R[2], R[1], M);            /* :78-79 */
    MRotY(M, q[2]); MxM(R[3], R[2], M);            /* :81-82 */
    MRotX(M, q[3]); MxM(R[4], R[3], M);            /* :84-85 */
    MRotY(M, q[4]); MxM(R[5], R[4], M);            /* :87-88 */
    MRotX(M, q[5]); MxM(R[6], R[5], M);            /* :90-91 */
    /* T0 = (-offsetX/2, offsetY/2, offsetZ/2) then the aliased MxV(T0,R0,T0)  (:99-103) */
    T[0][0] = -offX / 2; T[0][1] = 0.0 / 2; T[0][2] = 0.0 / 2;
    MxV_inplace_alias(T[0], R[0]);
    v[0] = 0; v[1] = 0; v[2] = 145 * 2;            MxV(tmp, R[0], v); v_add(T[1], T[0], tmp);      /* :105-110 */
    v[0] = 0; v[1] = 0; v[2] = 0;                  MxV(tmp, R[1], v); v_add(T[2], T[1], tmp);      /* :112-117 (T2_t) */
    v[0] = 0; v[1] = 0; v[2] = 270;                MxV(tmp, R[2], v); v_add(T[3], T[2], tmp);      /* :132-137 */
    v[0] = 134; v[1] = 0; v[2] = 70;               MxV(tmp, R[3], v); v_add(T[4], tmp, T[3]);      /* :140-145 */
    v[0] = 168; v[1] = 0; v[2] = 0;                MxV(tmp, R[4], v); v_add(T[5], T[4], tmp);      /* :147-152 */
    v[0] = 72 - 13 / 2; v[1] = 0; v[2] = 0;        MxV(tmp, R[5], v); v_add(T[6], T[5], tmp);      /* :154-159  (integer 13/2 = 6) */
    v[0] = 13 / 2 + 60; v[1] = 0; v[2] = 0;        MxV(tmp, R[6], v); v_add(T[7], T[6], tmp);      /* :161-166 */
    for (int f = 0; f < 7; f++) { memcpy(fr[f], R[f], 9 * sizeof(double)); memcpy(fr[f] + 9, T[f], 3 * sizeof(double)); }
    memcpy(fr[7], R[6], 9 * sizeof(double)); memcpy(fr[7] + 9, T[7], 3 * sizeof(double));           /* EE / rod pose = (R6, T7) */
}

extern "C" void ckco_link_frames(const ckco_world* w, const double q[12], double frames[CKCO_NUM_FRAMES][12]) {
    robot_frames(0.0, w->D, q, frames);                                   /* robot 1, offsetRot = 0 */
    robot_frames(3.14159265 / 2 + 0.0, w->D, q + 6, frames + 8);          /* robot 2 (:177) */
    /* quirk frames */
    memcpy(frames[16], frames[10], 9 * sizeof(double)); memcpy(frames[16] + 9, frames[2] + 9, 3 * sizeof(double));   /* (R22, T2_t)  :373 */
    memcpy(frames[17], frames[2], 9 * sizeof(double));  memcpy(frames[17] + 9, frames[10] + 9, 3 * sizeof(double));  /* (R2, T2_t2)  :397,:415 */
    /* static obstacle poses */
    double Robs[9], Mobs[9];
    for (int k = 0; k < 3; k++) { memset(frames[18 + k], 0, 12 * sizeof(double)); frames[18 + k][0] = frames[18 + k][4] = frames[18 + k][8] = 1; }
    if (w->env == 1) {                                                    /* :363-406 */
        MRotZ(Robs, 0); MxM(Mobs, Robs, Robs);
        const double dObs = 250;
        for (int k = 0; k < 3; k++) memcpy(frames[18 + k], Mobs, sizeof(Mobs));
        frames[19][9] = dObs;  frames[19][10] = dObs;
        frames[20][9] = -dObs; frames[20][10] = -dObs;
    } else {                                                              /* :431-456 */
        MRotZ(Mobs, 0); memcpy(frames[18], Mobs, sizeof(Mobs));
        MRotX(Mobs, 3.14); memcpy(frames[19], Mobs, sizeof(Mobs));
        frames[19][11] = 790;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* tolerance query                                                                                    */
/* ------------------------------------------------------------------------------------------------ */
/* Lower bound on the distance between two boxes: max over the 15 SAT axes of (gap / |axis|).
 * A: box a in frame A; B: box b, with (R,T) taking frame-B coordinates to frame-A coordinates. */
static double obb_gap_lower_bound(const obb_node& a, const obb_node& b, const double R[9], const double T[3]) {
    /* box b centre and axes in frame A */
    double cb[3], bax[9];
    MxV(cb, R, b.c); v_add(cb, cb, T);
    for (int j = 0; j < 3; j++) { double t[3]; MxV(t, R, &b.ax[3 * j]); bax[3 * j] = t[0]; bax[3 * j + 1] = t[1]; bax[3 * j + 2] = t[2]; }
    double d[3]; v_sub(d, cb, a.c);
    double t[3], Rab[3][3], AR[3][3];
    for (int i = 0; i < 3; i++) {
        t[i] = v_dot(d, &a.ax[3 * i]);
        for (int j = 0; j < 3; j++) { Rab[i][j] = v_dot(&a.ax[3 * i], &bax[3 * j]); AR[i][j] = fabs(Rab[i][j]); }
    }
    double best = -1e300;
    for (int i = 0; i < 3; i++) {
        double g = fabs(t[i]) - (a.h[i] + b.h[0] * AR[i][0] + b.h[1] * AR[i][1] + b.h[2] * AR[i][2]);
        best = std::max(best, g);
    }
    for (int j = 0; j < 3; j++) {
        double g = fabs(t[0] * Rab[0][j] + t[1] * Rab[1][j] + t[2] * Rab[2][j]) - (b.h[j] + a.h[0] * AR[0][j] + a.h[1] * AR[1][j] + a.h[2] * AR[2][j]);
        best = std::max(best, g);
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            double len2 = 1.0 - Rab[i][j] * Rab[i][j];
            if (len2 < 1e-12) continue;
            double tl = fabs(t[i2] * Rab[i1][j] - t[i1] * Rab[i2][j]);
            double ra = a.h[i1] * AR[i2][j] + a.h[i2] * AR[i1][j];
            double rb = b.h[j1] * AR[i][j2] + b.h[j2] * AR[i][j1];
            best = std::max(best, (tl - ra - rb) / sqrt(len2));
        }
    return best;
}

struct query {
    const mesh* ma; const mesh* mb;
    double R[9], T[3];      /* frame B -> frame A, as PQP_Tolerance computes (R = Ra^T Rb, T = Ra^T (Tb - Ta)) */
    double tol;
    int hit;
    int64_t bv_tests, tri_tests;
    double best;            /* for distance queries */
};

static double leaf_tri_dist(const query& Q, int ta, int tb) {
    /* PQP TriDistance: triangle of model 2 is transformed into model 1's frame, then TriDist */
    double S[3][3], Tt[3][3], P[3], Qp[3];
    for (int v = 0; v < 3; v++) {
        v_cpy(S[v], &Q.ma->tris[9 * ta + 3 * v]);
        double t[3]; MxV(t, Q.R, &Q.mb->tris[9 * tb + 3 * v]);
        Tt[v][0] = t[0] + Q.T[0]; Tt[v][1] = t[1] + Q.T[1]; Tt[v][2] = t[2] + Q.T[2];
    }
    return ckco_tri_dist(P, Qp, S, Tt);
}

static double box_size(const obb_node& n) { return n.h[0] * n.h[0] + n.h[1] * n.h[1] + n.h[2] * n.h[2]; }

static void tol_recurse(query& Q, int na, int nb) {
    const obb_node& a = Q.ma->nodes[na];
    const obb_node& b = Q.mb->nodes[nb];
    if (a.tri >= 0 && b.tri >= 0) {
        Q.tri_tests++;
        if (leaf_tri_dist(Q, a.tri, b.tri) <= Q.tol) Q.hit = 1;
        return;
    }
    int ca[2], cb[2];
    if (b.tri >= 0 || (a.tri < 0 && box_size(a) > box_size(b))) { ca[0] = a.left; ca[1] = a.right; cb[0] = cb[1] = nb; }
    else { ca[0] = ca[1] = na; cb[0] = b.left; cb[1] = b.right; }
    double g[2];
    for (int k = 0; k < 2; k++) { Q.bv_tests++; g[k] = obb_gap_lower_bound(Q.ma->nodes[ca[k]], Q.mb->nodes[cb[k]], Q.R, Q.T); }
    const int first = (g[1] < g[0]) ? 1 : 0;
    for (int s = 0; s < 2; s++) {
        const int k = s == 0 ? first : 1 - first;
        if (g[k] > Q.tol) continue;
        tol_recurse(Q, ca[k], cb[k]);
        if (Q.hit) return;
    }
}

static void dist_recurse(query& Q, int na, int nb) {
    const obb_node& a = Q.ma->nodes[na];
    const obb_node& b = Q.mb->nodes[nb];
    if (a.tri >= 0 && b.tri >= 0) {
        double d = leaf_tri_dist(Q, a.tri, b.tri);
        if (d < Q.best) Q.best = d;
        return;
    }
    int ca[2], cb[2];
    if (b.tri >= 0 || (a.tri < 0 && box_size(a) > box_size(b))) { ca[0] = a.left; ca[1] = a.right; cb[0] = cb[1] = nb; }
    else { ca[0] = ca[1] = na; cb[0] = b.left; cb[1] = b.right; }
    double g[2];
    for (int k = 0; k < 2; k++) g[k] = obb_gap_lower_bound(Q.ma->nodes[ca[k]], Q.mb->nodes[cb[k]], Q.R, Q.T);
    const int first = (g[1] < g[0]) ? 1 : 0;
    for (int s = 0; s < 2; s++) {
        const int k = s == 0 ? first : 1 - first;
        if (g[k] >= Q.best) continue;
        dist_recurse(Q, ca[k], cb[k]);
    }
}

static void setup_query(query& Q, const ckco_world* w, const ckco_pair& p, const double frames[CKCO_NUM_FRAMES][12]) {
    Q.ma = &w->meshes[p.mesh_a]; Q.mb = &w->meshes[p.mesh_b];
    const double* Ra = frames[p.frame_a]; const double* Ta = Ra + 9;
    const double* Rb = frames[p.frame_b]; const double* Tb = Rb + 9;
    /* PQP_Tolerance: MTxM(res->R,R1,R2); VmV(Ttemp,T2,T1); MTxV(res->T,R1,Ttemp) */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            Q.R[3 * i + j] = (Ra[0 + i] * Rb[0 + j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j]);
    double tt[3]; v_sub(tt, Tb, Ta);
    for (int i = 0; i < 3; i++) Q.T[i] = (Ra[0 + i] * tt[0] + Ra[3 + i] * tt[1] + Ra[6 + i] * tt[2]);
    Q.tol = w->tol; Q.hit = 0; Q.bv_tests = 0; Q.tri_tests = 0; Q.best = 1e300;
}

extern "C" int ckco_collision_state(const ckco_world* w, const double q[12], int early_exit, ckco_collide_stats* stats, uint8_t* per_pair) {
    double frames[CKCO_NUM_FRAMES][12];
    ckco_link_frames(w, q, frames);
    int any = 0;
    /* env 2 pre-check (collisionDetection.cpp:427-428): T7z > 800 or T72z > 800 => collision */
    if (w->env == 2 && (frames[7][11] > 800 || frames[15][11] > 800)) {
        if (per_pair) memset(per_pair, 0, w->pairs.size());
        return 1;
    }
    const bool all = (!early_exit) || stats || per_pair;
    for (size_t i = 0; i < w->pairs.size(); i++) {
        query Q; setup_query(Q, w, w->pairs[i], frames);
        if (!Q.ma->nodes.empty() && !Q.mb->nodes.empty()) {
            if (stats) stats->bv_tests++;
            if (obb_gap_lower_bound(Q.ma->nodes[0], Q.mb->nodes[0], Q.R, Q.T) <= Q.tol) tol_recurse(Q, 0, 0);
        }
        if (stats) { stats->bv_tests += Q.bv_tests; stats->tri_tests += Q.tri_tests; stats->pairs_hit += Q.hit; }
        if (per_pair) per_pair[i] = (uint8_t)Q.hit;
        if (Q.hit) { any = 1; if (!all) return 1; }
    }
    return any;
}

extern "C" int ckco_collision_state_brute(const ckco_world* w, const double q[12], uint8_t* per_pair) {
    double frames[CKCO_NUM_FRAMES][12];
    ckco_link_frames(w, q, frames);
    if (w->env == 2 && (frames[7][11] > 800 || frames[15][11] > 800)) { if (per_pair) memset(per_pair, 0, w->pairs.size()); return 1; }
    int any = 0;
    for (size_t i = 0; i < w->pairs.size(); i++) {
        query Q; setup_query(Q, w, w->pairs[i], frames);
        int hit = 0;
        for (int ta = 0; ta < Q.ma->ntris() && !hit; ta++)
            for (int tb = 0; tb < Q.mb->ntris(); tb++)
                if (leaf_tri_dist(Q, ta, tb) <= Q.tol) { hit = 1; break; }
        if (per_pair) per_pair[i] = (uint8_t)hit;
        any |= hit;
    }
    return any;
}

extern "C" double ckco_min_pair_distance(const ckco_world* w, const double q[12]) {
    double frames[CKCO_NUM_FRAMES][12];
    ckco_link_frames(w, q, frames);
    double best = 1e300;
    for (size_t i = 0; i < w->pairs.size(); i++) {
        query Q; setup_query(Q, w, w->pairs[i], frames);
        if (Q.ma->nodes.empty() || Q.mb->nodes.empty()) continue;
        Q.best = best;
        if (obb_gap_lower_bound(Q.ma->nodes[0], Q.mb->nodes[0], Q.R, Q.T) < Q.best) dist_recurse(Q, 0, 0);
        best = std::min(best, Q.best);
    }
    return best;
}
