#ifndef OPENMM_COMPAT_NONBONDEDFORCE_H_
#define OPENMM_COMPAT_NONBONDEDFORCE_H_
#include "openmm/Force.h"
namespace OpenMM {
// Only the two settings the MPID plugin reads when it asks for automatic PME parameters.
class OPENMM_EXPORT NonbondedForce : public Force {
public:
    NonbondedForce() : cutoff(1.0), ewaldTol(5e-4) {}
    double getCutoffDistance() const { return cutoff; }
    void setCutoffDistance(double d) { cutoff = d; }
    double getEwaldErrorTolerance() const { return ewaldTol; }
    void setEwaldErrorTolerance(double t) { ewaldTol = t; }
protected:
    ForceImpl* createImpl() const { return 0; }
private:
    double cutoff, ewaldTol;
};
} // namespace OpenMM
#endif
