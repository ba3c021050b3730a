#ifndef OPENMM_COMPAT_LANGEVININTEGRATOR_H_
#define OPENMM_COMPAT_LANGEVININTEGRATOR_H_
#include "openmm/Integrator.h"
namespace OpenMM {
class OPENMM_EXPORT LangevinIntegrator : public Integrator {
public:
    LangevinIntegrator(double temperature, double frictionCoeff, double stepSize)
        : Integrator(stepSize), temperature(temperature), friction(frictionCoeff) {}
    double getTemperature() const { return temperature; }
    double getFriction() const { return friction; }
private:
    double temperature, friction;
};
} // namespace OpenMM
#endif
