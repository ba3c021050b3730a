#ifndef OPENMM_COMPAT_INTEGRATOR_H_
#define OPENMM_COMPAT_INTEGRATOR_H_
#include "openmm/internal/windowsExport.h"
namespace OpenMM {
// The hot path needs no dynamics: integrators only carry their parameters here.
class OPENMM_EXPORT Integrator {
public:
    Integrator(double stepSize = 0.001) : stepSize(stepSize) {}
    virtual ~Integrator() {}
    double getStepSize() const { return stepSize; }
    void setStepSize(double size) { stepSize = size; }
private:
    double stepSize;
};
} // namespace OpenMM
#endif
