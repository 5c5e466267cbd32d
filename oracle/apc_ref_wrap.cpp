/*
 * apc_ref_wrap.cpp -- ORACLE INFRASTRUCTURE (test/baseline only, NOT product code).
 * extern "C" batch wrapper around the reference's OWN proj_classes/apc_class.cpp, which oracle/Makefile compiles from
 * where it lies under /root/reference into oracle/_ref/libapc_ref.so (no reference source is copied into the repo).
 * Used to pin the oracle restatement of FK / IK / closure projection from source.
 */
#include "apc_class.h"      /* -I/root/reference/proj_classes */
#include <unistd.h>
#include <fcntl.h>

static two_robots* make(int env) {
    const double D = env == 1 ? 900. : 1200., L = env == 1 ? 300. : 500.;
    fflush(stdout); int so = dup(1); int dn = open("/dev/null", O_WRONLY); dup2(dn, 1);
    two_robots* r = new two_robots({ -D / 2, 0, 0 }, { D / 2, 0, PI_ }, L);     /* StateValidityCheckerPCS.h:37 */
    fflush(stdout); dup2(so, 1); close(dn); close(so);
    return r;
}
static two_robots* g[3];
static two_robots* get(int env) { if (!g[env]) g[env] = make(env); return g[env]; }

extern "C" void apcref_fk(int env, long n, const double* q6, const signed char* robot, double* T16) {
    two_robots* r = get(env);
    for (long i = 0; i < n; i++) {
        State q(q6 + 6 * i, q6 + 6 * i + 6);
        r->FKsolve_rob(q, robot[i]);
        Matrix T = robot[i] == 1 ? r->get_FK_solution_T1() : r->get_FK_solution_T2();
        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) T16[16 * i + 4 * a + b] = T[a][b];
    }
}

/* IKproject(q1,q2,ac,ik) of StateValidityCheckerPCS.cpp:140-162 issued on the reference's two_robots */
extern "C" void apcref_project(int env, long n, const double* q12, const signed char* ac, const signed char* ik, unsigned char* ok, double* q_out) {
    two_robots* r = get(env);
    const double L = env == 1 ? 300. : 500.;
    Matrix Q = { {1, 0, 0, L}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1} };
    for (long i = 0; i < n; i++) {
        State q1(q12 + 12 * i, q12 + 12 * i + 6), q2(q12 + 12 * i + 6, q12 + 12 * i + 12);
        bool v;
        if (!ac[i]) { v = r->calc_specific_IK_solution_R1(Q, q1, ik[i]); if (v) q2 = r->get_IK_solution_q2(); }
        else { v = r->calc_specific_IK_solution_R2(Q, q2, ik[i]); if (v) q1 = r->get_IK_solution_q1(); }
        ok[i] = v;
        for (int j = 0; j < 6; j++) { q_out[12 * i + j] = q1[j]; q_out[12 * i + 6 + j] = q2[j]; }
    }
}

extern "C" void apcref_limits(int env, long n, const double* q12, unsigned char* ok) {
    two_robots* r = get(env);
    for (long i = 0; i < n; i++) { State q(q12 + 12 * i, q12 + 12 * i + 12); ok[i] = r->check_angle_limits(q); }
}
