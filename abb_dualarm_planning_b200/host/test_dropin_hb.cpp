/*
 * test_dropin_hb.cpp -- exercises the hybrid flavour (validity_checkers/StateValidityCheckerHB: GD projection for sampling,
 * APC bisection for the local connection, as planners/CBiRRT_HB.cpp uses it): for every input configuration
 *   IKproject(q)                       kdl::GD + limits + collision          (HB.cpp:109-147)
 *   identify_state_ik(q1, q2)          branches of the projected point       (shared with PCS)
 *   checkMotionRBS(q, q', ac, ik)      APC local connection to a second projected point
 * usage: test_dropin_hb <in.txt> <out.txt>     in: n then n lines "qa[12] qb[12]"
 * out: per line  va vb  ik0 ik1  ika0 ikb0(second point)  rbs   qa_projected[12]  qb_projected[12]
 */
#include "StateValidityCheckerHB.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    StateValidityChecker svc(1);
    std::ifstream in(argv[1]); std::ofstream out(argv[2]); out.precision(17);
    int n; in >> n;
    for (int i = 0; i < n; i++) {
        State a(12), b(12);
        for (int j = 0; j < 12; j++) in >> a[j];
        for (int j = 0; j < 12; j++) in >> b[j];
        const bool va = svc.IKproject(a), vb = svc.IKproject(b);
        State ida = { -1, -1 }, idb = { -1, -1 };
        bool rbs = false;
        if (va && vb) {
            State a1(a.begin(), a.begin() + 6), a2(a.begin() + 6, a.end()), b1(b.begin(), b.begin() + 6), b2(b.begin() + 6, b.end());
            ida = svc.identify_state_ik(a1, a2); idb = svc.identify_state_ik(b1, b2);
            if (ida[0] >= 0 && ida[0] == idb[0]) rbs = svc.checkMotionRBS(a1, a2, b1, b2, 0, (int)ida[0], 0, 0);
        }
        out << va << " " << vb << " " << ida[0] << " " << ida[1] << " " << idb[0] << " " << idb[1] << " " << rbs;
        for (int j = 0; j < 12; j++) out << " " << a[j];
        for (int j = 0; j < 12; j++) out << " " << b[j];
        out << "\n";
    }
    srand(3);
    State s = svc.sample_q_gd();
    out << "sample";
    for (double v : s) out << " " << v;
    out << "\n";
    return 0;
}
