// Plugin entry points of the MPIDB200 platform kernel.  OpenMM's plugin loader dlopens every library in
// lib/plugins and calls registerPlatforms() then registerKernelFactories(); tests link the library and
// call registerMPIDB200KernelFactories() directly -- the same three-function pattern the reference uses
// (platforms/cuda/src/MPIDCudaKernelFactory.cpp:36-66, platforms/reference/src/MPIDReferenceKernelFactory.cpp).
#include "MPIDB200KernelFactory.h"
#include "MPIDB200Kernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/windowsExport.h"

using namespace OpenMM;

MPIDB200Platform::MPIDB200Platform(const std::string& platformName) : name(platformName) {
    std::vector<std::string> names;
    names.push_back(Precision()); names.push_back(DeviceIndex()); names.push_back(Solver());
    platformProperties(names);
    setPropertyDefaultValue(Precision(), "mixed");
    setPropertyDefaultValue(DeviceIndex(), "0");
    setPropertyDefaultValue(Solver(), "DIIS");
}

void MPIDB200Platform::contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {
    PlatformData* data = new PlatformData();
    data->positions = &context.positions;
    data->forces = &context.forces;
    data->box = context.box;
    data->properties = properties;
    context.setPlatformData(data);
}
void MPIDB200Platform::contextDestroyed(ContextImpl& context) const {
    delete static_cast<PlatformData*>(context.getPlatformData());
    context.setPlatformData(0);
}

KernelImpl* MPIDB200KernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name == CalcMPIDForceKernel::Name())
        return new B200CalcMPIDForceKernel(name, platform, context.getSystem(), context);
    throw OpenMMException(("Tried to create kernel with illegal kernel name '" + name + "'").c_str());
}

#ifndef MPIDB200_PLATFORM_NAME
#define MPIDB200_PLATFORM_NAME "MPIDB200"
#endif

// The exported names registerPlatforms / registerKernelFactories are shared by every OpenMM plugin library, so
// inside this library they are never called by name (another plugin's definition could interpose).
static void addPlatform(const char* platformName) {
    try { Platform::getPlatformByName(platformName); }
    catch (...) { Platform::registerPlatform(new MPIDB200Platform(platformName)); }
}
static void addKernelFactory(const char* platformName) {
    try {
        Platform& platform = Platform::getPlatformByName(platformName);
        platform.registerKernelFactory(CalcMPIDForceKernel::Name(), new MPIDB200KernelFactory());
    }
    catch (...) {
        // the platform is not present: nothing to register
    }
}

extern "C" OPENMM_EXPORT void registerPlatforms() { addPlatform(MPIDB200_PLATFORM_NAME); }
extern "C" OPENMM_EXPORT void registerKernelFactories() { addKernelFactory(MPIDB200_PLATFORM_NAME); }

extern "C" OPENMM_EXPORT void registerMPIDB200KernelFactories() {
    addPlatform(MPIDB200_PLATFORM_NAME);
    addKernelFactory(MPIDB200_PLATFORM_NAME);
}

// Same kernel under another platform name.  Used by the parity tests to run the reference's own
// TestCudaMPIDForce.cpp unmodified: that file asks for the platform called "CUDA".
extern "C" OPENMM_EXPORT void registerMPIDB200KernelFactoriesAs(const char* platformName) {
    addPlatform(platformName);
    addKernelFactory(platformName);
}
