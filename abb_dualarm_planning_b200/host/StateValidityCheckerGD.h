/*
 * StateValidityCheckerGD.h -- drop-in for validity_checkers/StateValidityCheckerGD.{h,cpp} (+ the kdl class it inherits,
 * proj_classes/kdl_class.{h,cpp}) of the reference.  Same class name and member signatures; compute goes through the
 * batched C ABI (include/ckc.h).  -DCKC_FLAVOUR_RX selects the relaxed-constraint checker
 * (validity_checkers/StateValidityCheckerRX.{h,cpp}) which shares this interface.
 */
#ifndef CHECKER_H_
#define CHECKER_H_

#include "ckc_checker_common.h"

class StateValidityChecker {
public:
    StateValidityChecker(const ob::SpaceInformationPtr& si, int env = 1) : mysi_(si.get()), ckc_(env) { init(env); }   /* GD.h:35-43 */
    StateValidityChecker(int env = 1) : mysi_(nullptr), ckc_(env) { init(env); }                                          /* GD.h:44-51 */

    /* ---- kdl (proj_classes/kdl_class.h:33-108).  Not provided: GD_JL (ChainIkSolverPos_NR_JL; only tests/test_kdl_JL.cpp calls
     *      it) and the public KDL:: data members (chain, jointpositions, ...: the Orocos objects do not exist here) ---- */
    bool GD(State q);                                              /* kdl_class.cpp:68-140 */
    State get_GD_result() { return q_solution; }
    bool check_angle_limits(State q);                              /* kdl_class.cpp:213-242 */
    void FK(State q);                                              /* kdl_class.cpp:251-278 */
    Matrix get_FK_solution() { return T_fk; }
    Matrix T_fk, T_fk_temp;
    void initVector(State& V, int n) { V.resize(n); }
    void initMatrix(Matrix& M, int n, int m) { M.resize(n); for (int i = 0; i < n; ++i) M[i].resize(m); }
    double deg2rad(double deg) { return deg * PI / 180.0; }
    void printMatrix(Matrix M);
    void printVector(State p);
    void clearMatrix(Matrix& M) { for (auto& r : M) for (double& v : r) v = 0; }
    void log_q(State q);                                           /* kdl_class.cpp: one configuration to ../paths/path.txt */
    State get_robots_properties() { return { 166, 124, 270, 70, 149.6, 152.4, (double)(72 - 13 / 2), (double)(60 + 13 / 2) }; }   /* kdl_class.cpp:5-12 (integer division) */
    Matrix get_Tpose() { return T_pose; }
    bool include_joint_limits = true;
    int get_IK_counter() { return IK_counter; }
    double get_IK_time() { return IK_time; }
    int IK_counter; double IK_time;

    /* ---- collisionDetection ---- */
    int collision_state(Matrix M, State q1, State q2);             /* collisionDetection.cpp:31-494 */
    int get_collisionCheck_counter() { return collisionCheck_counter; }
    double get_collisionCheck_time() { return collisionCheck_time; }
    int collisionCheck_counter; double collisionCheck_time;

    /* ---- StateValidityChecker ---- */
    bool isValid(const ob::State*);                                /* GD.cpp:115-134 / RX.cpp:140-155 */
#ifdef CKC_FLAVOUR_RX
    bool isValidRBS(State q);                                      /* RX.cpp:201-214 */
    bool check_project(const ob::State*);                          /* RX.cpp:90-103 */
    bool check_relax_constraint(State q);                          /* RX.cpp:105-135 */
    void set_epsilon(double n) { epsilon = n; }
    double get_epsilon() { return epsilon; }
#else
    bool isValidRBS(State& q);                                     /* GD.cpp:180-196 */
    bool IKproject(const ob::State*, bool = true);                 /* GD.cpp:82-94 */
    bool IKproject(State&, bool = true);                           /* GD.cpp:96-110 */
#endif
    bool checkMotion(const ob::State*, const ob::State*);          /* GD.cpp:136-175 / RX.cpp:157-196 */
    bool checkMotionRBS(const ob::State*, const ob::State*);       /* GD.cpp:199-211 */
    bool checkMotionRBS(State, State, int, int);                   /* GD.cpp:214-241 */
    bool reconstructRBS(const ob::State*, const ob::State*, Matrix&);          /* GD.cpp:246-256 / RX.cpp:263-273 */
    bool reconstructRBS(State, State, Matrix&, int, int, int);                 /* GD.cpp:258-296 / RX.cpp:275-310 */
#ifndef CKC_FLAVOUR_RX
    bool isValidSew(State&);                                       /* GD.cpp:301-318 */
    bool checkMotionSew(const ob::State*, const ob::State*);       /* GD.cpp:320-338 */
    bool reconstructSew(const ob::State*, const ob::State*, Matrix&);          /* GD.cpp:342-350 */
    bool reconstructSew(State, State, Matrix&);                    /* GD.cpp:352-376 */
#else
    bool sample_q(State&);                                         /* RX.cpp:73-88: one draw, may fail */
    bool IKproject(const ob::State*, bool = true);                 /* declared by RX.h:86-87, defined nowhere in the reference: here the */
    bool IKproject(State&, bool = true);                           /* relaxed check of the state as it is (nothing is projected in RX) */
#endif
    State sample_q();                                              /* GD.cpp:53-80 / RX.cpp:53-70 */
    void retrieveStateVector(const ob::State*, State&);
    void retrieveStateVector(const ob::State*, State&, State&);
    void updateStateVector(const ob::State*, State);
    void seperate_Vector(State, State&, State&);
    State join_Vectors(State, State);
    void Join_States(State&, State, State);
    void printStateVector(const ob::State*);                       /* GD.cpp:41-51 */
    double normDistance(State, State);
    double stateDistance(const ob::State*, const ob::State*);      /* GD.cpp:382-389 */
    double maxDistance(State, State);
    double MaxAngleDistance(State, State);
    void defaultSettings() {}
    int get_valid_solution_index() { return valid_solution_index; }
    Matrix getPMatrix() { return P; }
    Matrix getQ() { return Q; }
    void setQ();
    void setP();
    double iden = 0;
    double get_iden() { return iden; }
    int get_n() { return n; }
    double get_RBS_tol() { return RBS_tol; }
    void set_q_prev(State q) { for (size_t i = 0; i < q.size(); i++) q_prev[i] = q[i]; }
    void log_q(State q, bool New);                                 /* GD.cpp:430-445 */
    int get_isValid_counter() { return isValid_counter; }
    int isValid_counter;
    /* performance log (GD.h:171-204, GD.cpp:447-470) */
    double total_runtime; clock_t startTime, endTime; int nodes_in_path, nodes_in_trees; double PlanDistance; bool final_solved;
    double local_connection_time; int local_connection_count, local_connection_success_count;
    State sampling_counter; double sampling_time;
    void initiate_log_parameters();
    void LogPerf2file();

    /* ---- batch forms ---- */
    void isValidBatch(std::vector<State>& q, std::vector<unsigned char>& valid);
    void checkMotionRBSBatch(const std::vector<State>& a, const std::vector<State>& b, std::vector<unsigned char>& ok);
    ckc_ctx* ckc() { return ckc_.ctx(); }

private:
    void init(int env);
    ob::SpaceInformation* mysi_;
    CkcContext ckc_;
    void ckc_check(int rc);
    State q_solution, q_prev;
    int valid_solution_index = 0;
    double L;
    Matrix P, Q, T_pose;
    double dq = 0.05;
    int n = 12;
    bool withObs = true;
    double RBS_tol = 0.05;
    int RBS_max_depth = 150;
    double epsilon = 0.3;
};

#endif
