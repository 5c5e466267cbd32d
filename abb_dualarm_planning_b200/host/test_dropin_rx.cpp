/*
 * test_dropin_rx.cpp -- exercises the relaxed-constraint flavour (validity_checkers/StateValidityCheckerRX.{h,cpp}) of the
 * drop-in GD class: check_relax_constraint (:105-135), isValidRBS (:201-214) and the host-side checkMotionRBS recursion.
 * usage: test_dropin_rx <eps> <in.txt> <out.txt>     in: n then n lines "qa[12] qb[12]"
 * out: per line  relax(qa) valid(qa) valid(qb) rbs(qa, qb)
 */
#include "StateValidityCheckerGD.h"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    StateValidityChecker svc(1);
    svc.set_epsilon(atof(argv[1]));
    std::ifstream in(argv[2]); std::ofstream out(argv[3]);
    int n; in >> n;
    for (int i = 0; i < n; i++) {
        State a(12), b(12);
        for (int j = 0; j < 12; j++) in >> a[j];
        for (int j = 0; j < 12; j++) in >> b[j];
        const bool ra = svc.check_relax_constraint(a), va = svc.isValidRBS(a), vb = svc.isValidRBS(b);
        const bool rbs = va && vb && svc.checkMotionRBS(a, b, 0, 0);
        out << ra << " " << va << " " << vb << " " << rbs << "\n";
    }
    return 0;
}
