/*
 * StateValidityCheckerPCS.h -- drop-in for validity_checkers/StateValidityCheckerPCS.{h,cpp} of the reference
 * (class name, public members and signatures kept so that the planners (planners/ *_PCS.cpp) compile unchanged).  Every member that
 * computes anything forwards to the batched C ABI of include/ckc.h (sm_100a kernels); the extra *Batch members expose
 * the batch form directly.  Members cite the reference line they replace.
 */
#ifndef CHECKER_H_
#define CHECKER_H_

#include "ckc_checker_common.h"

class StateValidityChecker {
public:
    /** Constructors  (StateValidityCheckerPCS.h:35-51) */
    StateValidityChecker(const ob::SpaceInformationPtr& si, int env = 1) : mysi_(si.get()), ckc_(env) { init(env); }
    StateValidityChecker(int env = 1) : mysi_(nullptr), ckc_(env) { init(env); }

    /* ---- two_robots (proj_classes/apc_class.h:47-119), every public member ---- */
    void FKsolve_rob(State q, int robot_num);                                    /* apc_class.cpp:51-76 */
    Matrix get_FK_solution_T1() { return T_fk_solution_1; }
    State get_FK_solution_p1() { return p_fk_solution_1; }
    Matrix get_FK_solution_T2() { return T_fk_solution_2; }
    State get_FK_solution_p2() { return p_fk_solution_2; }
    bool IKsolve_rob(Matrix T, int robot_num, int solution_num);                 /* apc_class.cpp:114-245 */
    bool calc_specific_IK_solution_R1(Matrix T, State q1, int IKsol);            /* apc_class.cpp:356-371 */
    bool calc_specific_IK_solution_R2(Matrix T, State q2, int IKsol);            /* apc_class.cpp:373-389 */
    State get_IK_solution_q1() { return q_IK_solution_1; }
    State get_IK_solution_q2() { return q_IK_solution_2; }
    bool IsRobotsFeasible_R1(Matrix T, State q1);                                /* apc_class.cpp:392-407 */
    bool IsRobotsFeasible_R2(Matrix T, State q2);                                /* apc_class.cpp:409-424 */
    int calc_all_IK_solutions_1(Matrix T);                                       /* apc_class.cpp:249-262: the 8 branches in one batch */
    int calc_all_IK_solutions_2(Matrix T);                                       /* apc_class.cpp:264-277 */
    State get_all_IK_solutions_1(int i) { return Q_IK_solutions_1[i]; }
    State get_all_IK_solutions_2(int i) { return Q_IK_solutions_2[i]; }
    int get_valid_IK_solutions_indices_1(int i) { return (int)valid_IK_solutions_indices_1[i]; }
    int get_valid_IK_solutions_indices_2(int i) { return (int)valid_IK_solutions_indices_2[i]; }
    int get_IK_number() { return IK_number; }
    Matrix get_T2() { return T2_; }
    bool check_angle_limits(State q);                                            /* apc_class.cpp:304-334 */
    void initVector(State& V, int n) { V.resize(n); }
    void initMatrix(Matrix& M, int n, int m) { M.resize(n); for (int i = 0; i < n; ++i) M[i].resize(m); }
    double deg2rad(double deg) { return deg * PI_ / 180.0; }
    void printMatrix(Matrix M);
    void printVector(State p);
    Matrix MatricesMult(Matrix M1, Matrix M2);                                   /* apc_class.cpp:427-436 */
    bool InvertMatrix(Matrix M, Matrix& Minv);                                   /* apc_class.cpp:439-575 (4x4 cofactor inverse) */
    void clearMatrix(Matrix& M) { for (auto& r : M) for (double& v : r) v = 0; }
    State rand_q_ambient();                                                      /* apc_class.cpp:643-661 */
    State get_robots_properties() { return { 166, 124, 270, 70, 149.6, 152.4, 72 - 13. / 2, 60 + 13. / 2 }; }   /* apc_class.cpp:5-12 */
    bool include_joint_limits = true;
    int get_IK_counter() { return IK_counter; }
    double get_IK_time() { return IK_time; }
    int IK_counter; double IK_time;

    /* ---- collisionDetection (validity_checkers/collisionDetection.h) ---- */
    int collision_state(Matrix M, State q1, State q2);                           /* collisionDetection.cpp:31-494 */
    int get_collisionCheck_counter() { return collisionCheck_counter; }
    double get_collisionCheck_time() { return collisionCheck_time; }
    int collisionCheck_counter; double collisionCheck_time;

    /* ---- StateValidityChecker (StateValidityCheckerPCS.h:53-125) ---- */
    bool isValid(const ob::State*);                                              /* PCS.cpp:323-357 */
    bool isValid(const ob::State*, int, int);                                    /* PCS.cpp:360-387 */
    bool isValidRBS(State&, State&, int, int);                                   /* PCS.cpp:455-466 */
    bool checkMotion(const ob::State*, const ob::State*, int, int);              /* PCS.cpp:389-432 */
    bool checkMotionRBS(const ob::State*, const ob::State*, int, int);           /* PCS.cpp:470-483 */
    bool checkMotionRBS(State, State, State, State, int, int, int, int);         /* PCS.cpp:486-511 */
    bool reconstructRBS(const ob::State*, const ob::State*, Matrix&, int, int);  /* PCS.cpp:531-543 */
    bool reconstructRBS(State, State, State, State, int, int, Matrix&, int, int, int);   /* PCS.cpp:545-577 */
    double normDistanceDuo(State, State, State, State);
    void midpoint(State, State, State, State, State&, State&);
    void retrieveStateVector(const ob::State*, State&);                          /* 12-vector form (HB.h:83) */
    void retrieveStateVector(const ob::State*, State&, State&);
    void updateStateVector(const ob::State*, State);                             /* 12-vector form (HB.h:87) */
    void updateStateVector(const ob::State*, State, State);
    void printStateVector(const ob::State*);
    void defaultSettings() {}
    double normDistance(State, State);
    double stateDistance(const ob::State*, const ob::State*);
    State sample_q();                                                            /* PCS.cpp:234-259 */
    bool sample_q(ob::State* st);                                                /* PCS.cpp:196-231 */
    bool close_chain(const ob::State*, int);                                     /* PCS.cpp:59-92 */
    bool IKproject(State&, State&, int, int);                                    /* PCS.cpp:140-162 */
    bool IKproject(State&, State&, int);                                         /* PCS.cpp:165-193: first of the 8 branches (random order) that projects and is identified */
    bool IKproject(State&, State&, State&, int&, State);                         /* PCS.cpp:94-138 */
    State identify_state_ik(const ob::State*, State = {-1, -1});                 /* PCS.cpp:261-273 */
    State identify_state_ik(State, State, State = {-1, -1});                     /* PCS.cpp:275-316 */
    State join_Vectors(State, State);
    void seperate_Vector(State, State&, State&);
    void log_q(ob::State*);                                                      /* PCS.cpp:602-608 */
    void log_q(State, State);                                                    /* PCS.cpp:610-626 */
    int get_valid_solution_index() { return valid_solution_index; }
    double iden = 0;
    double get_iden() { return iden; }                                           /* PCS.h:166-167: time spent in identify_state_ik */
    void setQ();                                                                 /* PCS.h:136-147 */
    void setP();                                                                 /* PCS.h:150-158 */
    Matrix getPMatrix() { return P; }
    Matrix getQ() { return Q; }
    double get_RBS_tol() { return RBS_tol; }
    int get_isValid_counter() { return isValid_counter; }
    int isValid_counter;

    /* performance log (StateValidityCheckerPCS.h:171-204, .cpp:628-651) */
    double total_runtime; clock_t startTime, endTime; int nodes_in_path, nodes_in_trees; double PlanDistance; bool final_solved;
    double local_connection_time; int local_connection_count, local_connection_success_count; double sampling_time; State sampling_counter;
    void initiate_log_parameters();
    void LogPerf2file();

#ifdef CKC_FLAVOUR_HB
    /* ---- hybrid checker (validity_checkers/StateValidityCheckerHB.{h,cpp}): sampling / projection by the KDL Newton
     *      projection, local connection by the APC bisection ---- */
    bool GD(State q);                                                            /* kdl_class.cpp:68-140 */
    State get_GD_result() { return q_solution; }
    void FK(State q);                                                            /* kdl_class.cpp:251-278 */
    Matrix get_FK_solution() { return T_fk; }
    Matrix T_fk;
    bool IKproject(const ob::State*, bool = true);                               /* HB.cpp:95-106 */
    bool IKproject(State&, bool = true);                                         /* HB.cpp:109-122: GD projection of all 12 joints (+ collision) */
    State sample_q_gd();                                                         /* HB.cpp:73-93 sample_q: GD from uniform seeds */
    bool checkMotion(const ob::State*, const ob::State*);                        /* HB.cpp:233-271: serial grid, each point through isValid (GD) */
    bool isValidSew(State&);                                                     /* HB.cpp:405-421 */
    bool checkMotionSew(const ob::State*, const ob::State*);                     /* HB.cpp:423-441 */
    bool reconstructSew(const ob::State*, const ob::State*, Matrix&);            /* HB.cpp:446-453 */
    bool reconstructSew(State, State, Matrix&);                                  /* HB.cpp:455-480 */
    double maxDistance(State, State);
    double MaxAngleDistance(State, State);
    void Join_States(State&, State, State);
    int get_n() { return 12; }
#endif
#ifdef CKC_FLAVOUR_RSS
    /* ---- random singular sampling checker (validity_checkers/StateValidityCheckerRSS.{h,cpp}) ---- */
    bool sampleSingular(ob::State*);                                             /* RSS.cpp:255-306 */
    bool checkMotionRBS(const ob::State*, const ob::State*, int sg, int active_chain, int ik_sol);          /* RSS.cpp:566-583 */
    bool checkMotionRBS(State, State, State, State, State, State, int, int, int, int);                      /* RSS.cpp:586-613 */
    bool reconstructRBS(const ob::State*, const ob::State*, Matrix&, int sg, int active_chain, int ik_sol); /* RSS.cpp:684-700 */
    bool reconstructRBS(State, State, State, State, State, State, int, int, Matrix&, int, int, int);        /* RSS.cpp:702-741 */
#endif
    int get_countSolutions() { return countSolutions; }                          /* apc_class.h:62 */

    /* ---- batch forms (new): what a planner calls to submit whole batches ---- */
    /** n configurations (q[i] = 12 joints), per-sample active chain / IK branch; returns valid flags, projects q in place */
    void isValidBatch(std::vector<State>& q, const std::vector<int>& ac, const std::vector<int>& ik, std::vector<unsigned char>& valid);
    /** n edges at once: ok[i] = checkMotionRBS(a[i], b[i], ac[i], ik[i]) */
    void checkMotionRBSBatch(const std::vector<State>& a, const std::vector<State>& b, const std::vector<int>& ac, const std::vector<int>& ik,
                             std::vector<unsigned char>& ok);
    ckc_ctx* ckc() { return ckc_.ctx(); }

private:
    void init(int env);
    ob::SpaceInformation* mysi_;
    CkcContext ckc_;
    State q_IK_solution_1, q_IK_solution_2, q_solution;
    Matrix T_fk_solution_1, T_fk_solution_2, Q_IK_solutions_1, Q_IK_solutions_2, T2_;
    State p_fk_solution_1, p_fk_solution_2, valid_IK_solutions_indices_1, valid_IK_solutions_indices_2;
    int IK_number = 0, valid_solution_index = 0;
    int calc_all(const Matrix& T, int robot, Matrix& sols, State& idx);
    void ckc_check(int rc);                                                      /* a failed device call ends the process like the reference's exit(-1) */
    int countSolutions = 0;
    double L;
    Matrix Q, P;
    bool withObs = true;
    double RBS_tol = 0.05;
    int RBS_max_depth = 150;
};

#endif
