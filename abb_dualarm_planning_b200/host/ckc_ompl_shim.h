/*
 * ckc_ompl_shim.h -- the handful of OMPL types the reference's StateValidityChecker signatures mention
 * (ob::State, ob::RealVectorStateSpace::StateType::values, ob::SpaceInformationPtr), for builds WITHOUT OMPL
 * (this image has none).  With OMPL installed, compile with -DCKC_HAVE_OMPL and the real headers are used instead.
 */
#ifndef CKC_OMPL_SHIM_H_
#define CKC_OMPL_SHIM_H_
#ifdef CKC_HAVE_OMPL
#include <ompl/base/SpaceInformation.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/base/State.h>
#else
#include <memory>
namespace ompl { namespace base {
class State {
public:
    virtual ~State() {}
    template <class T> const T* as() const { return static_cast<const T*>(this); }
    template <class T> T* as() { return static_cast<T*>(this); }
};
class RealVectorStateSpace {
public:
    class StateType : public State {
    public:
        explicit StateType(unsigned n = 12) : values(new double[n]()) {}
        ~StateType() override { delete[] values; }
        double* values;
    };
};
class StateSpace {};
class SpaceInformation {
public:
    State* allocState() const { return new RealVectorStateSpace::StateType(12); }
    void freeState(State* s) const { delete s; }
};
typedef std::shared_ptr<SpaceInformation> SpaceInformationPtr;
} }
#endif
#endif
