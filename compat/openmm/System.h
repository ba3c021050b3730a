#ifndef OPENMM_COMPAT_SYSTEM_H_
#define OPENMM_COMPAT_SYSTEM_H_
#include "openmm/Force.h"
#include "openmm/Vec3.h"
#include <vector>
namespace OpenMM {
class OPENMM_EXPORT System {
public:
    System();
    ~System();
    int getNumParticles() const { return (int) masses.size(); }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size()-1; }
    double getParticleMass(int index) const { return masses[index]; }
    void setParticleMass(int index, double mass) { masses[index] = mass; }
    int addForce(Force* force) { forces.push_back(force); return (int) forces.size()-1; }  // takes ownership
    int getNumForces() const { return (int) forces.size(); }
    const Force& getForce(int index) const { return *forces[index]; }
    Force& getForce(int index) { return *forces[index]; }
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c);
    bool usesPeriodicBoundaryConditions() const;
private:
    std::vector<double> masses;
    std::vector<Force*> forces;
    Vec3 box[3];
};
} // namespace OpenMM
#endif
