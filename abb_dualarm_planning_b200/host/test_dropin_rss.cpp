/*
 * test_dropin_rss.cpp -- RSS + HB members of the drop-in class: sampleSingular, the singular RBS variant, GD projection.
 * usage: test_dropin_rss <out.txt>
 */
#include "StateValidityCheckerRSS.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    StateValidityChecker svc(1);
    std::ofstream out(argv[1]); out.precision(17);
    srand(11);
    ob::SpaceInformation si;
    for (int k = 0; k < 12; k++) {
        ob::State* a = si.allocState(); ob::State* b = si.allocState();
        svc.sampleSingular(a);
        State b12 = svc.sample_q();
        State b1(b12.begin(), b12.begin() + 6), b2(b12.begin() + 6, b12.end());
        svc.updateStateVector(b, b1, b2);
        State a1(6), a2(6); svc.retrieveStateVector(a, a1, a2);
        State ida = svc.identify_state_ik(a), idb = svc.identify_state_ik(b);
        /* connect with arm 2 active on a's branch, as planners/CBiRRT_SG.cpp does for singular nodes */
        State pb1 = b1, pb2 = b2;
        bool proj = svc.IKproject(pb1, pb2, 1, (int)ida[1]) && !svc.collision_state(svc.getPMatrix(), pb1, pb2);
        bool rbs = false;
        if (proj) rbs = svc.checkMotionRBS(a1, a2, pb1, pb2, a1, a2, 1, (int)ida[1], 0, 0);
        out << ida[0] << " " << ida[1] << " " << proj << " " << rbs;
        for (double v : a1) out << " " << v; for (double v : a2) out << " " << v;
        for (double v : pb1) out << " " << v; for (double v : pb2) out << " " << v;
        out << "\n";
        si.freeState(a); si.freeState(b);
    }
    return 0;
}
