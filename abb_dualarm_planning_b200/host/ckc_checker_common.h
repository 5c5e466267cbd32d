/*
 * ckc_checker_common.h -- shared plumbing of the drop-in StateValidityChecker classes: owns one libckc context and
 * offers the reference's vector types.  Host language = the reference's (C++); all compute goes through include/ckc.h.
 */
#ifndef CKC_CHECKER_COMMON_H_
#define CKC_CHECKER_COMMON_H_

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "../../include/ckc.h"
#include "ckc_ompl_shim.h"

#define PI_ 3.1416                      /* proj_classes/apc_class.h:10 */
#define PI 3.1416                       /* proj_classes/kdl_class.h:24 */
#define ROBOTS_DISTANCE_ENV_I 900.      /* validity_checkers/StateValidityCheckerPCS.h:23-26 */
#define ROD_LENGTH_ENV_I 300.
#define ROBOTS_DISTANCE_ENV_II 1200.
#define ROD_LENGTH_ENV_II 500.

typedef std::vector<double> State;
typedef std::vector<std::vector<double> > Matrix;
namespace ob = ompl::base;

class CkcContext {
public:
    explicit CkcContext(int env) : env_(env), ctx_(nullptr) {
        /* the reference loads "./simulator/" relative to the cwd (collisionDetection.cpp:3); CKC_MESH_PATH overrides */
        const char* p = getenv("CKC_MESH_PATH");
        const char* dev = getenv("CKC_DEVICE");
        std::string path = p ? p : "./simulator";
        /* CKC_DEVICES=0,1,2,3: a multi-device context -- every batched member (sample_q's candidate batches, checkMotionRBSBatch,
         * isValidBatch ...) is then split over those GPUs inside the library; the planner code does not change */
        const char* devs = getenv("CKC_DEVICES");
        int rc;
        if (devs && strchr(devs, ',')) {
            std::vector<int> ids;
            for (const char* c = devs; *c; ) { ids.push_back(atoi(c)); c = strchr(c, ','); if (!c) break; c++; }
            rc = ckc_create_multi(&ctx_, env, path.c_str(), ids.data(), (int)ids.size());
        } else rc = ckc_create(&ctx_, env, path.c_str(), devs ? atoi(devs) : (dev ? atoi(dev) : 0));
        if (rc != CKC_OK) {           /* the reference exit(-1)s when a mesh cannot be opened (collisionDetection.cpp:503) */
            fprintf(stderr, "ckc_create(%s) failed: %d\n", path.c_str(), rc);
            exit(-1);
        }
    }
    ~CkcContext() { ckc_destroy(ctx_); }
    CkcContext(const CkcContext&) = delete;
    CkcContext& operator=(const CkcContext&) = delete;
    ckc_ctx* ctx() const { return ctx_; }
    int env() const { return env_; }
private:
    int env_;
    ckc_ctx* ctx_;
};

inline void ckc_pack(const State& q1, const State& q2, double* q) { for (int i = 0; i < 6; i++) { q[i] = q1[i]; q[6 + i] = q2[i]; } }
inline void ckc_unpack(const double* q, State& q1, State& q2) { q1.resize(6); q2.resize(6); for (int i = 0; i < 6; i++) { q1[i] = q[i]; q2[i] = q[6 + i]; } }

#endif
