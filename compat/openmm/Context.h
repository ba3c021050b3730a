#ifndef OPENMM_COMPAT_CONTEXT_H_
#define OPENMM_COMPAT_CONTEXT_H_
#include "openmm/Integrator.h"
#include "openmm/Platform.h"
#include "openmm/State.h"
#include "openmm/System.h"
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class ContextImpl;
class OPENMM_EXPORT Context {
public:
    Context(const System& system, Integrator& integrator);
    Context(const System& system, Integrator& integrator, Platform& platform);
    Context(const System& system, Integrator& integrator, Platform& platform, const std::map<std::string, std::string>& properties);
    ~Context();
    const System& getSystem() const;
    Integrator& getIntegrator() { return integrator; }
    const Platform& getPlatform() const;
    Platform& getPlatform();
    State getState(int types, bool enforcePeriodicBox = false, int groups = 0xFFFFFFFF) const;
    void setPositions(const std::vector<Vec3>& positions);
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c);
    void reinitialize(bool preserveState = false);
private:
    friend class Force;
    friend class Platform;
    Context(const Context&);
    Context& operator=(const Context&);
    ContextImpl* impl;
    Integrator& integrator;
    std::map<std::string, std::string> properties;
};
} // namespace OpenMM
#endif
