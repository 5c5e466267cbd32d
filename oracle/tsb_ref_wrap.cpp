/*
 * tsb_ref_wrap.cpp -- ORACLE INFRASTRUCTURE (test/baseline only, NOT product code).
 * extern "C" batch wrapper around the reference's OWN closure check
 *     verification_class::Tsb_matrix     validity_checkers/verification_class.cpp:141-170
 *     verification_class::test_constraint validity_checkers/verification_class.cpp:116-139   (0.1 rotation / 0.5 mm)
 * (the same expression and tolerances test_gd_kdl.cpp:61-75 applies to the output of kdl::GD), which oracle/Makefile
 * compiles from where it lies under /root/reference into oracle/_ref/libtsb_ref_env{1,2}.so together with the reference's
 * proj_classes/apc_class.cpp (no reference source is copied into the repo).  verification_class derives from the
 * OMPL-dependent StateValidityChecker; it only uses that base for initMatrix / get_robots_properties / printVector /
 * printMatrix / normDistance / get_RBS_tol, so the stand-in below derives from the reference's two_robots and adds the one
 * missing member.  ROBOTS_DISTANCE / ROD_LENGTH (verification_class.h:55-59) come from the compiler command line, one
 * library per environment.
 *
 * This is the reference-held pin of the GD flavour: every configuration the device reports as projected must pass it.
 */
#include "apc_class.h"      /* -I/root/reference/proj_classes */
#include <fstream>
#include <unistd.h>
#include <fcntl.h>

#ifndef ROBOTS_DISTANCE
#error "compile with -DROBOTS_DISTANCE=... -DROD_LENGTH=..."
#endif

class StateValidityChecker : public two_robots {
public:
    StateValidityChecker() : two_robots({ -ROBOTS_DISTANCE / 2., 0, 0 }, { ROBOTS_DISTANCE / 2., 0, PI_ }, ROD_LENGTH) {}
    double get_RBS_tol() { return 0.05; }      /* StateValidityCheckerPCS.h:218 */
};

/* -I/root/reference/validity_checkers ; the .cpp includes verification_class.h (no include guard there), which with no
 * PCS / PGD / HB macro defined includes nothing of OMPL */
#include "verification_class.cpp"

static verification_class* g_v;
static verification_class* get() {
    if (!g_v) {
        fflush(stdout); int so = dup(1); int dn = open("/dev/null", O_WRONLY); dup2(dn, 1);
        g_v = new verification_class();
        fflush(stdout); dup2(so, 1); close(dn); close(so);
    }
    return g_v;
}

/* Tsb[i] = 3 x 4 (row-major) of the reference's Tsb_matrix(q); pass[i] = test_constraint's rule evaluated silently */
extern "C" void tsbref_check(long n, const double* q12, double* Tsb12, unsigned char* pass) {
    verification_class* v = get();
    const double Tsd[3][4] = { { 1, 0, 0, (double)ROBOTS_DISTANCE }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 } };    /* verification_class.h:45 */
    for (long i = 0; i < n; i++) {
        State q(q12 + 12 * i, q12 + 12 * i + 12);
        Matrix T = v->Tsb_matrix(q);
        bool ok = true;
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 4; k++) {
                if (Tsb12) Tsb12[12 * i + 4 * j + k] = T[j][k];
                /* verification_class.cpp:122 (negated so that a NaN entry fails) */
                if (!((k < 3 && fabs(T[j][k] - Tsd[j][k]) <= 0.1) || (k == 3 && fabs(T[j][k] - Tsd[j][k]) <= 0.5))) ok = false;
            }
        pass[i] = ok ? 1 : 0;
    }
}

/* the reference's own member, prints on failure (used once by the tests to show that the silent rule above is the same) */
extern "C" int tsbref_test_constraint(const double* q12) {
    State q(q12, q12 + 12);
    return get()->test_constraint(q) ? 1 : 0;
}
