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

    /* ---- kdl (proj_classes/kdl_class.h) ---- */
    bool GD(State q);                                              /* kdl_class.cpp:68-140 */
    State get_GD_result() { return q_solution; }
    bool check_angle_limits(State q);                              /* kdl_class.cpp:213-242 */
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
    State sample_q();                                              /* GD.cpp:53-80 / RX.cpp:53-70 */
    void retrieveStateVector(const ob::State*, State&);
    void retrieveStateVector(const ob::State*, State&, State&);
    void updateStateVector(const ob::State*, State);
    void seperate_Vector(State, State&, State&);
    State join_Vectors(State, State);
    double normDistance(State, State);
    void defaultSettings() {}
    Matrix getPMatrix() { return P; }
    double get_RBS_tol() { return RBS_tol; }
    int get_isValid_counter() { return isValid_counter; }
    int isValid_counter;
    State sampling_counter; double sampling_time;
    void initiate_log_parameters();

    /* ---- batch forms ---- */
    void isValidBatch(std::vector<State>& q, std::vector<unsigned char>& valid);
    void checkMotionRBSBatch(const std::vector<State>& a, const std::vector<State>& b, std::vector<unsigned char>& ok);
    ckc_ctx* ckc() { return ckc_.ctx(); }

private:
    void init(int env);
    ob::SpaceInformation* mysi_;
    CkcContext ckc_;
    State q_solution, q_prev;
    double L;
    Matrix P;
    int n = 12;
    bool withObs = true;
    double RBS_tol = 0.05;
    int RBS_max_depth = 150;
    double epsilon = 0.3;
};

#endif
