#ifndef OPENMM_COMPAT_VERLETINTEGRATOR_H_
#define OPENMM_COMPAT_VERLETINTEGRATOR_H_
#include "openmm/Integrator.h"
namespace OpenMM {
class OPENMM_EXPORT VerletIntegrator : public Integrator {
public:
    explicit VerletIntegrator(double stepSize) : Integrator(stepSize) {}
};
} // namespace OpenMM
#endif
