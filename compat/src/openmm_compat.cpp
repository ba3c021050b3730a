// Implementation of the MPIDB200 "openmm compat" layer: the small subset of the OpenMM C++
// runtime (System / Context / Platform / Kernel plumbing) that the MPID plugin is written
// against.  It exists because OpenMM itself is not installable in the build container; plugin
// code compiled against these headers compiles unchanged against real OpenMM headers.
#include "openmm/Context.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/NonbondedForceImpl.h"
#include "ReferencePlatform.h"
#include <cmath>

namespace OpenMM {

// ---------------------------------------------------------------- System
System::System() {
    box[0] = Vec3(2, 0, 0); box[1] = Vec3(0, 2, 0); box[2] = Vec3(0, 0, 2);
}
System::~System() {
    for (Force* f : forces) delete f;
}
void System::setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) {
    if (a[1] != 0.0 || a[2] != 0.0) throw OpenMMException("First periodic box vector must be parallel to x.");
    if (b[2] != 0.0) throw OpenMMException("Second periodic box vector must be in the x-y plane.");
    box[0] = a; box[1] = b; box[2] = c;
}
bool System::usesPeriodicBoundaryConditions() const {
    for (const Force* f : forces) if (f->usesPeriodicBoundaryConditions()) return true;
    return false;
}

// ---------------------------------------------------------------- Force
ForceImpl& Force::getImplInContext(Context& context) {
    for (ForceImpl* impl : context.impl->getForceImpls())
        if (&impl->getOwner() == this) return *impl;
    throw OpenMMException("getImplInContext: This Force is not present in the Context");
}
const ForceImpl& Force::getImplInContext(const Context& context) const {
    for (ForceImpl* impl : context.impl->getForceImpls())
        if (&impl->getOwner() == this) return *impl;
    throw OpenMMException("getImplInContext: This Force is not present in the Context");
}
ContextImpl& Force::getContextImpl(Context& context) { return *context.impl; }

// ---------------------------------------------------------------- Platform
static std::vector<Platform*>& platformRegistry() {
    static std::vector<Platform*> platforms;
    return platforms;
}
Platform::~Platform() {
    for (auto& kv : kernelFactories) {
        bool shared = false;  // a factory may be registered under several names
        for (auto& other : kernelFactories) if (&other != &kv && other.second == kv.second && other.first < kv.first) shared = true;
        if (!shared) delete kv.second;
    }
}
const std::string& Platform::getPropertyValue(const Context& context, const std::string& property) const {
    auto it = context.properties.find(property);
    if (it != context.properties.end()) return it->second;
    return getPropertyDefaultValue(property);
}
const std::string& Platform::getPropertyDefaultValue(const std::string& property) const {
    auto it = defaultProperties.find(property);
    if (it == defaultProperties.end()) throw OpenMMException("getPropertyDefaultValue: Illegal property name");
    return it->second;
}
void Platform::setPropertyDefaultValue(const std::string& property, const std::string& value) {
    bool known = false;
    for (const std::string& n : propertyNames) if (n == property) known = true;
    if (!known) throw OpenMMException("setPropertyDefaultValue: Illegal property name");
    defaultProperties[property] = value;
}
void Platform::registerKernelFactory(const std::string& name, KernelFactory* factory) {
    auto it = kernelFactories.find(name);
    if (it != kernelFactories.end() && it->second != factory) delete it->second;
    kernelFactories[name] = factory;
}
bool Platform::supportsKernels(const std::vector<std::string>& kernelNames) const {
    for (const std::string& n : kernelNames) if (kernelFactories.find(n) == kernelFactories.end()) return false;
    return true;
}
Kernel Platform::createKernel(const std::string& name, ContextImpl& context) const {
    auto it = kernelFactories.find(name);
    if (it == kernelFactories.end())
        throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
    return Kernel(it->second->createKernelImpl(name, *this, context));
}
void Platform::registerPlatform(Platform* platform) { platformRegistry().push_back(platform); }
int Platform::getNumPlatforms() { return (int) platformRegistry().size(); }
Platform& Platform::getPlatform(int index) {
    if (index < 0 || index >= getNumPlatforms()) throw OpenMMException("Invalid platform index");
    return *platformRegistry()[index];
}
Platform& Platform::getPlatformByName(const std::string& name) {
    for (Platform* p : platformRegistry()) if (p->getName() == name) return *p;
    throw OpenMMException("There is no registered Platform called \"" + name + "\"");
}

// The compat runtime always offers a host-memory Reference platform.
namespace {
struct RegisterReferencePlatform {
    RegisterReferencePlatform() { Platform::registerPlatform(new ReferencePlatform()); }
} registerReferencePlatformInstance;
}

void ReferencePlatform::contextCreated(ContextImpl& context, const std::map<std::string, std::string>&) const {
    PlatformData* data = new PlatformData();
    data->numParticles = context.getSystem().getNumParticles();
    data->stepCount = 0;
    data->time = 0.0;
    data->positions = &context.positions;
    data->velocities = &context.velocities;
    data->forces = &context.forces;
    data->periodicBoxVectors = context.box;
    data->boxSize = Vec3(context.box[0][0], context.box[1][1], context.box[2][2]);
    data->periodicBoxSize = &data->boxSize;
    context.setPlatformData(data);
}
void ReferencePlatform::contextDestroyed(ContextImpl& context) const {
    delete static_cast<PlatformData*>(context.getPlatformData());
    context.setPlatformData(0);
}

// ---------------------------------------------------------------- ContextImpl / Context
ContextImpl::ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform,
                         const std::map<std::string, std::string>& properties)
    : owner(owner), system(system), integrator(integrator), platform(platform), platformData(0) {
    int n = system.getNumParticles();
    if (n == 0) throw OpenMMException("Cannot create a Context for a System with no particles");
    positions.assign(n, Vec3());
    velocities.assign(n, Vec3());
    forces.assign(n, Vec3());
    system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
    platform->contextCreated(*this, properties);
    for (int i = 0; i < system.getNumForces(); i++) {
        ForceImpl* impl = system.getForce(i).createImpl();
        if (impl) forceImpls.push_back(impl);
    }
    for (ForceImpl* impl : forceImpls) impl->initialize(*this);
}
ContextImpl::~ContextImpl() {
    for (ForceImpl* impl : forceImpls) delete impl;
    platform->contextDestroyed(*this);
}
void ContextImpl::setPositions(const std::vector<Vec3>& in) {
    if ((int) in.size() != system.getNumParticles())
        throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
    positions = in;
}
void ContextImpl::setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) {
    box[0] = a; box[1] = b; box[2] = c;
    if (platform->getName() == "Reference") {
        ReferencePlatform::PlatformData* data = static_cast<ReferencePlatform::PlatformData*>(platformData);
        data->boxSize = Vec3(a[0], b[1], c[2]);
    }
}
double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups) {
    for (Vec3& f : forces) f = Vec3();
    double energy = 0.0;
    for (ForceImpl* impl : forceImpls) energy += impl->calcForcesAndEnergy(*this, includeForces, includeEnergy, groups);
    return energy;
}

Context::Context(const System& system, Integrator& integrator)
    : impl(0), integrator(integrator) {
    impl = new ContextImpl(*this, system, integrator, &Platform::getPlatform(Platform::getNumPlatforms()-1), properties);
}
Context::Context(const System& system, Integrator& integrator, Platform& platform)
    : impl(0), integrator(integrator) {
    impl = new ContextImpl(*this, system, integrator, &platform, properties);
}
Context::Context(const System& system, Integrator& integrator, Platform& platform, const std::map<std::string, std::string>& props)
    : impl(0), integrator(integrator), properties(props) {
    impl = new ContextImpl(*this, system, integrator, &platform, properties);
}
Context::~Context() { delete impl; }
const System& Context::getSystem() const { return impl->getSystem(); }
const Platform& Context::getPlatform() const { return impl->getPlatform(); }
Platform& Context::getPlatform() { return impl->getPlatform(); }
void Context::setPositions(const std::vector<Vec3>& positions) { impl->setPositions(positions); }
void Context::setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { impl->setPeriodicBoxVectors(a, b, c); }
void Context::reinitialize(bool preserveState) {
    std::vector<Vec3> pos = impl->positions;
    Vec3 box[3] = {impl->box[0], impl->box[1], impl->box[2]};
    const System& system = impl->getSystem();
    Platform* platform = &impl->getPlatform();
    delete impl;
    impl = new ContextImpl(*this, system, integrator, platform, properties);
    if (preserveState) {
        impl->setPositions(pos);
        impl->setPeriodicBoxVectors(box[0], box[1], box[2]);
    }
}
State Context::getState(int types, bool, int groups) const {
    State state;
    bool wantForces = (types & State::Forces) != 0;
    bool wantEnergy = (types & State::Energy) != 0;
    if (wantForces || wantEnergy) {
        double e = impl->calcForcesAndEnergy(wantForces, wantEnergy, groups);
        state.energy = e;
        if (wantForces) state.forces = impl->forces;
    }
    if (types & State::Positions) state.positions = impl->positions;
    state.box[0] = impl->box[0]; state.box[1] = impl->box[1]; state.box[2] = impl->box[2];
    return state;
}

// ---------------------------------------------------------------- NonbondedForceImpl
void NonbondedForceImpl::calcPMEParameters(const System& system, const NonbondedForce& force, double& alpha,
                                           int& xsize, int& ysize, int& zsize, bool lj) {
    Vec3 a, b, c;
    system.getDefaultPeriodicBoxVectors(a, b, c);
    double tol = force.getEwaldErrorTolerance();
    alpha = std::sqrt(-std::log(2.0*tol))/force.getCutoffDistance();
    (void) lj;
    double denom = 3.0*std::pow(tol, 0.2);
    xsize = (int) std::ceil(2*alpha*a[0]/denom);
    ysize = (int) std::ceil(2*alpha*b[1]/denom);
    zsize = (int) std::ceil(2*alpha*c[2]/denom);
    xsize = xsize < 6 ? 6 : xsize;
    ysize = ysize < 6 ? 6 : ysize;
    zsize = zsize < 6 ? 6 : zsize;
}

} // namespace OpenMM
