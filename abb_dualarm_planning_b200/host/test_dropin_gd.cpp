/*
 * test_dropin_gd.cpp -- exercises the drop-in StateValidityChecker (GD flavour) like tests/test_rbs_gd.cpp of the reference
 * (IKproject(q) / isValidRBS / checkMotionRBS) on inputs from a file.  usage: test_dropin_gd <in.txt> <out.txt>
 */
#include "StateValidityCheckerGD.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    StateValidityChecker svc(1);
    std::ifstream in(argv[1]); std::ofstream out(argv[2]); out.precision(17);
    int n; in >> n;
    std::vector<State> all;
    for (int i = 0; i < n; i++) {
        State q(12); for (int j = 0; j < 12; j++) in >> q[j];
        all.push_back(q);
        State p = q;
        bool proj = svc.GD(p);
        State r = svc.get_GD_result();
        State v = q;
        bool valid = svc.isValidRBS(v);
        out << proj << " " << valid;
        for (double x : r) out << " " << x;
        out << "\n";
    }
    std::vector<unsigned char> valid;
    svc.isValidBatch(all, valid);
    out << "batch";
    for (unsigned char v : valid) out << " " << (int)v;
    out << "\n";
    return 0;
}
