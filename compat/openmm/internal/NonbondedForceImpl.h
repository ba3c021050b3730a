#ifndef OPENMM_COMPAT_NONBONDEDFORCEIMPL_H_
#define OPENMM_COMPAT_NONBONDEDFORCEIMPL_H_
#include "openmm/NonbondedForce.h"
#include "openmm/System.h"
namespace OpenMM {
class OPENMM_EXPORT NonbondedForceImpl {
public:
    // OpenMM 7.x rule for lj=false: alpha = sqrt(-ln(2 tol))/rc, n = ceil(2 alpha L / (3 tol^(1/5))), min 6.
    // (unpinned by any reference test; restated from the published formula.)
    static void calcPMEParameters(const System& system, const NonbondedForce& force, double& alpha,
                                  int& xsize, int& ysize, int& zsize, bool lj);
};
} // namespace OpenMM
#endif
