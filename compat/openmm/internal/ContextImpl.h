#ifndef OPENMM_COMPAT_CONTEXTIMPL_H_
#define OPENMM_COMPAT_CONTEXTIMPL_H_
#include "openmm/Context.h"
#include "openmm/internal/ForceImpl.h"
#include <vector>
namespace OpenMM {
// Holds positions/forces/box for a Context on a generic host-memory "platform data"
// block; platforms attach their own data through set/getPlatformData like in OpenMM.
class OPENMM_EXPORT ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform,
                const std::map<std::string, std::string>& properties);
    ~ContextImpl();
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    const void* getPlatformData() const { return platformData; }
    void setPlatformData(void* data) { platformData = data; }
    void getPositions(std::vector<Vec3>& out) const { out = positions; }
    void setPositions(const std::vector<Vec3>& in);
    void getForces(std::vector<Vec3>& out) const { out = forces; }
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c);
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF);
    void systemChanged() {}
    std::vector<ForceImpl*>& getForceImpls() { return forceImpls; }
    // Host-resident state every platform of this compat layer shares.
    std::vector<Vec3> positions, velocities, forces;
    Vec3 box[3];
private:
    friend class Context;
    Context& owner;
    const System& system;
    Integrator& integrator;
    Platform* platform;
    void* platformData;
    std::vector<ForceImpl*> forceImpls;
};
} // namespace OpenMM
#endif
