/*
 * test_dropin_pcs.cpp -- exercises the drop-in StateValidityChecker (PCS flavour) the way the reference's own Monte-Carlo
 * driver tests/test_rbs_pcs.cpp does (sample_q / IKproject(ac, ik) / identify_state_ik / checkMotionRBS), but on inputs
 * read from a file so that tests/test_gpu_dropin.py can compare every answer with the CPU oracle.
 * usage: test_dropin_pcs <in.txt> <out.txt>     in: n then n lines "q[12] ik qb[12]"
 */
#include "StateValidityCheckerPCS.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    StateValidityChecker svc(1);
    std::ifstream in(argv[1]); std::ofstream out(argv[2]); out.precision(17);
    int n; in >> n;
    for (int i = 0; i < n; i++) {
        State q1(6), q2(6), b1(6), b2(6); int ik;
        for (int j = 0; j < 6; j++) in >> q1[j];
        for (int j = 0; j < 6; j++) in >> q2[j];
        in >> ik;
        for (int j = 0; j < 6; j++) in >> b1[j];
        for (int j = 0; j < 6; j++) in >> b2[j];
        State a1 = q1, a2 = q2;
        bool proj = svc.IKproject(a1, a2, 0, ik);                           /* test_rbs_pcs.cpp:47 */
        int coll = proj ? svc.collision_state(svc.getPMatrix(), a1, a2) : -1;
        State id = proj ? svc.identify_state_ik(a1, a2) : State{ -1, -1 };   /* :54-57 */
        State v1 = q1, v2 = q2;
        bool valid = svc.isValidRBS(v1, v2, 0, ik);
        bool rbs = false;
        State e1 = b1, e2 = b2;
        if (valid && svc.isValidRBS(e1, e2, 0, ik)) rbs = svc.checkMotionRBS(v1, v2, e1, e2, 0, ik, 0, 0);   /* :63-67 */
        out << proj << " " << coll << " " << id[0] << " " << id[1] << " " << valid << " " << rbs;
        for (int j = 0; j < 6; j++) out << " " << a2[j];
        /* the OMPL-facing isValid(state) on the projected configuration and on a slightly opened one */
        ob::RealVectorStateSpace::StateType st(12);
        svc.updateStateVector(&st, a1, a2);
        out << " " << svc.isValid(&st);
        State o2 = a2; o2[1] += 0.004;
        svc.updateStateVector(&st, a1, o2);
        out << " " << svc.isValid(&st);
        out << "\n";
    }
    /* sampler + batch forms */
    srand(7);
    State s = svc.sample_q();
    State s1(s.begin(), s.begin() + 6), s2(s.begin() + 6, s.end());
    out << "sample " << svc.collision_state(svc.getPMatrix(), s1, s2) << " " << svc.check_angle_limits(s);
    for (double v : s) out << " " << v;
    out << "\n";
    svc.LogPerf2file();
    return 0;
}
