#ifndef OPENMM_COMPAT_REFERENCEPLATFORM_H_
#define OPENMM_COMPAT_REFERENCEPLATFORM_H_
#include "openmm/Platform.h"
#include "openmm/Vec3.h"
#include <vector>
namespace OpenMM {
class OPENMM_EXPORT ReferencePlatform : public Platform {
public:
    class PlatformData;
    ReferencePlatform() {}
    const std::string& getName() const { static const std::string name = "Reference"; return name; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const;
    void contextDestroyed(ContextImpl& context) const;
};
class ReferencePlatform::PlatformData {
public:
    int numParticles, stepCount;
    double time;
    void* positions;          // std::vector<Vec3>*
    void* velocities;         // std::vector<Vec3>*
    void* forces;             // std::vector<Vec3>*
    void* periodicBoxSize;    // Vec3*
    void* periodicBoxVectors; // Vec3[3]
    Vec3 boxSize;
};
} // namespace OpenMM
#endif
