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

    /* ---- two_robots (proj_classes/apc_class.h) ---- */
    bool calc_specific_IK_solution_R1(Matrix T, State q1, int IKsol);            /* apc_class.cpp:356-371 */
    bool calc_specific_IK_solution_R2(Matrix T, State q2, int IKsol);            /* apc_class.cpp:373-389 */
    State get_IK_solution_q1() { return q_IK_solution_1; }
    State get_IK_solution_q2() { return q_IK_solution_2; }
    bool check_angle_limits(State q);                                            /* apc_class.cpp:304-334 */
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
    void retrieveStateVector(const ob::State*, State&, State&);
    void updateStateVector(const ob::State*, State, State);
    void printStateVector(const ob::State*);
    void defaultSettings() {}
    double normDistance(State, State);
    double stateDistance(const ob::State*, const ob::State*);
    State sample_q();                                                            /* PCS.cpp:234-259 */
    bool sample_q(ob::State* st);                                                /* PCS.cpp:196-231 */
    bool IKproject(State&, State&, int, int);                                    /* PCS.cpp:140-162 */
    bool IKproject(State&, State&, State&, int&, State);                         /* PCS.cpp:94-138 */
    State identify_state_ik(const ob::State*, State = {-1, -1});                 /* PCS.cpp:261-273 */
    State identify_state_ik(State, State, State = {-1, -1});                     /* PCS.cpp:275-316 */
    State join_Vectors(State, State);
    void seperate_Vector(State, State&, State&);
    void log_q(State, State);
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
    bool IKproject(const ob::State*);                                            /* HB.cpp:109-122 */
    bool IKproject(State&);                                                      /* GD projection of all 12 joints + collision */
    State sample_q_gd();                                                         /* HB.cpp sample_q: GD from uniform seeds */
#endif
#ifdef CKC_FLAVOUR_RSS
    /* ---- random singular sampling checker (validity_checkers/StateValidityCheckerRSS.{h,cpp}) ---- */
    bool sampleSingular(ob::State*);                                             /* RSS.cpp:255-306 */
    bool checkMotionRBS(const ob::State*, const ob::State*, int sg, int active_chain, int ik_sol);          /* RSS.cpp:566-583 */
    bool checkMotionRBS(State, State, State, State, State, State, int, int, int, int);                      /* RSS.cpp:586-613 */
    int get_countSolutions() { return countSolutions; }
#endif

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
    int countSolutions = 0;
    double L;
    Matrix Q, P;
    bool withObs = true;
    double RBS_tol = 0.05;
    int RBS_max_depth = 150;
};

#endif
