/*
 * test_dropin_members.cpp -- exercises the members of the drop-in StateValidityChecker classes that the first round left untested
 * (checkMotion, reconstructRBS, GD isValid(const ob::State*), the 5- and 3-argument IKproject, the samplers with many
 * draws, FKsolve_rob / IKsolve_rob / calc_all_IK_solutions, close_chain, kdl::FK, sewing, the log writers) on inputs from a file, one
 * labelled output line per answer, so that tests/test_gpu_dropin.py can compare each with the CPU oracle / the reference binary's goldens.
 * Build once per flavour: (none) = PCS, -DCKC_FLAVOUR_GD (the GD header), -DCKC_FLAVOUR_RX.
 * usage: test_dropin_members <in.txt> <out.txt>   in: "nE nC" | nE lines qa[12] qb[12] ik | nC lines q[12] ac nn0 nn1
 */
#if defined(MEMBERS_GD) || defined(CKC_FLAVOUR_RX)
#include "StateValidityCheckerGD.h"
#else
#include "StateValidityCheckerPCS.h"
#endif

static void set_state(ob::RealVectorStateSpace::StateType& st, const State& q) { for (int j = 0; j < 12; j++) st.values[j] = q[j]; }
static void put_rows(std::ofstream& out, const Matrix& M) { out << " " << M.size(); for (auto& r : M) for (double v : r) out << " " << v; }

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    StateValidityChecker svc(1);
    std::ifstream in(argv[1]); std::ofstream out(argv[2]); out.precision(17);
    int nE, nC; in >> nE >> nC;
    std::vector<State> A(nE, State(12)), B(nE, State(12)), Cq(nC, State(12)); std::vector<int> K(nE), Cac(nC), Cn0(nC), Cn1(nC);
    for (int i = 0; i < nE; i++) { for (double& v : A[i]) in >> v; for (double& v : B[i]) in >> v; in >> K[i]; }
    for (int i = 0; i < nC; i++) { for (double& v : Cq[i]) in >> v; in >> Cac[i] >> Cn0[i] >> Cn1[i]; }
    ob::RealVectorStateSpace::StateType sa(12), sb(12);
    srand(11);
#if defined(MEMBERS_GD) || defined(CKC_FLAVOUR_RX)
    for (int i = 0; i < nE; i++) {
        set_state(sa, A[i]); set_state(sb, B[i]);
        Matrix M;
        const bool r = svc.reconstructRBS(&sa, &sb, M);                       /* GD.cpp:246-296 / RX.cpp:263-310 */
        out << "recon " << i << " " << r; put_rows(out, M); out << "\n";
        out << "cm " << i << " " << svc.checkMotion(&sa, &sb) << " " << svc.checkMotionRBS(&sa, &sb) << " " << svc.stateDistance(&sa, &sb) << "\n";
    }
    for (int i = 0; i < nC; i++) {
        set_state(sa, Cq[i]);
        const bool v = svc.isValid(&sa);                                      /* GD.cpp:115-134: projects the state in place */
        out << "isvalid " << i << " " << v; for (int j = 0; j < 12; j++) out << " " << sa.values[j]; out << "\n";
        svc.FK(Cq[i]); Matrix T = svc.get_FK_solution();                      /* kdl_class.cpp:251-285 */
        out << "fk " << i; for (auto& r : T) for (double x : r) out << " " << x; out << "\n";
    }
#ifdef CKC_FLAVOUR_RX
    svc.set_epsilon(0.5);
    int acc = 0;
    for (int i = 0; i < 20000 && acc < 5; i++) { State q(12); if (svc.sample_q(q)) { acc++; out << "sample " << acc; for (double x : q) out << " " << x; out << "\n"; } }
    out << "counts " << svc.sampling_counter[0] << " " << svc.sampling_counter[1] << " " << svc.get_epsilon() << "\n";
#else
    for (int k = 0; k < 20; k++) { State q = svc.sample_q(); out << "sample " << k; for (double x : q) out << " " << x; out << "\n"; }
    if (nE > 0) {                                                              /* sewing reconstruction of the first edge (literal GD.cpp:352-376) */
        Matrix M; const bool r = svc.reconstructSew(A[0], B[0], M);
        out << "sew " << r; put_rows(out, M); out << "\n";
    }
#endif
    out << "n " << svc.get_n() << " " << svc.get_RBS_tol() << "\n";
    svc.total_runtime = 1.5; svc.final_solved = true; svc.nodes_in_path = 7;
    svc.LogPerf2file();
#else
    for (int i = 0; i < nE; i++) {
        set_state(sa, A[i]); set_state(sb, B[i]);
        Matrix M;
        const bool r = svc.reconstructRBS(&sa, &sb, M, 0, K[i]);              /* PCS.cpp:531-577 */
        out << "recon " << i << " " << r; put_rows(out, M); out << "\n";
        out << "cm " << i << " " << svc.checkMotion(&sa, &sb, 0, K[i]) << " " << svc.stateDistance(&sa, &sb) << "\n";   /* PCS.cpp:389-432 */
    }
    for (int i = 0; i < nC; i++) {
        State q1(Cq[i].begin(), Cq[i].begin() + 6), q2(Cq[i].begin() + 6, Cq[i].end()), ik(2);
        int ac = Cac[i];
        const bool v = svc.IKproject(q1, q2, ik, ac, { (double)Cn0[i], (double)Cn1[i] });      /* PCS.cpp:94-138 */
        out << "ikp5 " << i << " " << v << " " << ac << " " << ik[0] << " " << ik[1]; for (double x : q1) out << " " << x; for (double x : q2) out << " " << x; out << "\n";
        State p1(Cq[i].begin(), Cq[i].begin() + 6), p2(Cq[i].begin() + 6, Cq[i].end());
        const bool w = svc.IKproject(p1, p2, Cac[i]);                                         /* PCS.cpp:165-193 (random branch order) */
        out << "ikp3 " << i << " " << w; for (double x : p1) out << " " << x; for (double x : p2) out << " " << x; out << "\n";
        if (i < 40) {
            State a1(Cq[i].begin(), Cq[i].begin() + 6);
            svc.FKsolve_rob(a1, 1 + (i & 1)); Matrix T = (i & 1) ? svc.get_FK_solution_T2() : svc.get_FK_solution_T1();   /* apc_class.cpp:51-76 */
            out << "fk " << i; for (auto& r : T) for (double x : r) out << " " << x; out << "\n";
            const int cnt = (i & 1) ? svc.calc_all_IK_solutions_2(T) : svc.calc_all_IK_solutions_1(T);                    /* apc_class.cpp:249-277 */
            out << "ikall " << i << " " << cnt;
            for (int c = 0; c < cnt; c++) {
                out << " " << ((i & 1) ? svc.get_valid_IK_solutions_indices_2(c) : svc.get_valid_IK_solutions_indices_1(c));
                State s = (i & 1) ? svc.get_all_IK_solutions_2(c) : svc.get_all_IK_solutions_1(c);
                for (double x : s) out << " " << x;
            }
            out << "\n";
            out << "feas " << i << " " << svc.IsRobotsFeasible_R1(svc.getQ(), a1) << " " << svc.get_countSolutions() << "\n";
        }
    }
    for (int k = 0; k < 50; k++) { State q = svc.sample_q(); out << "sample " << k; for (double x : q) out << " " << x; out << "\n"; }
    Matrix Qi; svc.InvertMatrix(svc.getQ(), Qi);
    out << "qinv"; for (auto& r : svc.MatricesMult(svc.getQ(), Qi)) for (double x : r) out << " " << x; out << "\n";
    out << "iden " << (svc.get_iden() >= 0) << " " << svc.get_RBS_tol() << "\n";
    svc.total_runtime = 1.5; svc.final_solved = true; svc.nodes_in_path = 7;
    svc.LogPerf2file();
#endif
    return 0;
}
