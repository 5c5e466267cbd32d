/*
 * surface_check.cpp -- compile-time proof for tests/test_abi_cpu.py (g++ -fsyntax-only, once per flavour): a class that inherits the
 * drop-in StateValidityChecker publicly, the way every reference planner does (planners/CBiRRT_PCS.h:66, SBL_PCS.h:87, ...), and calls
 * each member the reference's planners and run/ drivers call (tests/golden/dropin_surface.json, produced by
 * tools/scan_dropin_surface.py from planners/ *_<flavour>*.cpp and run/plan_<flavour>*.cpp) with the argument types used there.
 * Select the flavour with -DSURFACE_PCS / _GD / _HB / _RX / _SG.
 */
#if defined(SURFACE_PCS)
#include "StateValidityCheckerPCS.h"
#elif defined(SURFACE_GD)
#include "StateValidityCheckerGD.h"
#elif defined(SURFACE_HB)
#include "StateValidityCheckerHB.h"
#elif defined(SURFACE_RX)
#include "StateValidityCheckerRX.h"
#elif defined(SURFACE_SG)
#include "StateValidityCheckerRSS.h"
#else
#error "define SURFACE_<flavour>"
#endif

class Planner : public StateValidityChecker {
public:
    Planner(const ob::SpaceInformationPtr& si, int env) : StateValidityChecker(si, env) {}
    void calls(ob::State* st, const ob::State* a, const ob::State* b);
};

void Planner::calls(ob::State* st, const ob::State* a, const ob::State* b) {
    State q(12), q1(6), q2(6), ik(2);
    Matrix M;
    bool ok = true; double d = 0; int n = 0;
    defaultSettings(); initiate_log_parameters(); LogPerf2file(); printStateVector(a);
#if defined(SURFACE_PCS) || defined(SURFACE_SG) || defined(SURFACE_HB)
    retrieveStateVector(a, q1, q2);
    ok &= collision_state(getPMatrix(), q1, q2) != 0;
    ik = identify_state_ik(a);
    ok &= checkMotionRBS(a, b, 0, 3);
    ok &= reconstructRBS(a, b, M, 0, 3);
#endif
#if defined(SURFACE_PCS) || defined(SURFACE_SG)
    retrieveStateVector(a, q);
    updateStateVector(st, q); updateStateVector(st, q1, q2);
    ok &= calc_specific_IK_solution_R1(getQ(), q1, 2);
    q2 = get_IK_solution_q2();
    ik = identify_state_ik(q1, q2);
    q = join_Vectors(q1, q2);
    ok &= sample_q(st);
    ok &= reconstructRBS(q1, q2, q1, q2, 0, 3, M, 0, 1, 1);
#endif
#if defined(SURFACE_PCS)
    ok &= calc_specific_IK_solution_R2(getQ(), q2, 2);
    q1 = get_IK_solution_q1();
    ik = identify_state_ik(q1, q2, ik);
    d += get_IK_time() + get_collisionCheck_time() + get_RBS_tol() + get_iden() + stateDistance(a, b);
    n += get_IK_counter() + get_collisionCheck_counter() + get_isValid_counter();
#endif
#if defined(SURFACE_SG)
    ok &= sampleSingular(st);
    ok &= checkMotionRBS(a, b, 1, 0, 3);
    ok &= reconstructRBS(a, b, M, 1, 0, 3);
#endif
#if defined(SURFACE_HB)
    ok &= IKproject(a);
    d += get_iden();
#endif
#if defined(SURFACE_GD) || defined(SURFACE_RX)
    retrieveStateVector(a, q);
    updateStateVector(st, q);
    ok &= checkMotionRBS(a, b);
    ok &= reconstructRBS(a, b, M);
    ok &= reconstructRBS(q, q, M, 0, 1, 1);
    n += get_n();
#endif
#if defined(SURFACE_GD)
    ok &= IKproject(a); ok &= IKproject(a, false);
    ok &= reconstructSew(a, b, M);
    q = sample_q();
    d += get_IK_time() + get_collisionCheck_time() + get_RBS_tol() + stateDistance(a, b);
    n += get_IK_counter() + get_collisionCheck_counter() + get_isValid_counter();
#endif
#if defined(SURFACE_RX)
    ok &= check_project(a);
    ok &= sample_q(q);
    set_epsilon(0.5);
#endif
    (void)ok; (void)d; (void)n;
}
