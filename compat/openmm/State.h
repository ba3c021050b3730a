#ifndef OPENMM_COMPAT_STATE_H_
#define OPENMM_COMPAT_STATE_H_
#include "openmm/Vec3.h"
#include <vector>
namespace OpenMM {
class OPENMM_EXPORT State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16 };
    State() : energy(0.0) {}
    const std::vector<Vec3>& getPositions() const { return positions; }
    const std::vector<Vec3>& getForces() const { return forces; }
    double getPotentialEnergy() const { return energy; }
    double getKineticEnergy() const { return 0.0; }
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
private:
    friend class Context;
    std::vector<Vec3> positions, forces;
    double energy;
    Vec3 box[3];
};
} // namespace OpenMM
#endif
