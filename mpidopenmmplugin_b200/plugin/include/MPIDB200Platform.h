// MPIDB200 platform object of the OpenMM-facing plugin layer.
//
// In a real OpenMM installation a Platform owns positions / forces / box for each Context (for the
// reference's CUDA plugin that is CudaPlatform::PlatformData -> CudaContext, reference:
// platforms/cuda/src/MPIDCudaKernelFactory.cpp:69-70).  This platform keeps that state in host memory
// (the layout of OpenMM's Reference platform data, reference: platforms/reference/src/
// MPIDReferenceKernels.cpp:43-66) and hands it to the device engine through the C ABI of
// include/mpidb200.h, so the only thing a different host platform has to supply is the three
// accessors of StateAccess below (INTEGRATION.md shows the CudaContext variant).
#ifndef MPIDB200_PLATFORM_H_
#define MPIDB200_PLATFORM_H_

#include "openmm/Platform.h"
#include "openmm/Vec3.h"
#include "openmm/internal/ContextImpl.h"
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class MPIDB200Platform : public Platform {
public:
    // platform properties (OpenMM CUDA platform names: "Precision", "DeviceIndex")
    static const std::string& Precision()   { static const std::string k = "Precision";   return k; }
    static const std::string& DeviceIndex() { static const std::string k = "DeviceIndex"; return k; }
    static const std::string& Solver()      { static const std::string k = "MutualSolver"; return k; }

    explicit MPIDB200Platform(const std::string& platformName = "MPIDB200");
    const std::string& getName() const { return name; }
    double getSpeed() const { return 200.0; }
    bool supportsDoublePrecision() const { return true; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const;
    void contextDestroyed(ContextImpl& context) const;

    // Per-Context state: where the kernel reads positions / box and adds forces.
    struct PlatformData {
        std::vector<Vec3>* positions;
        std::vector<Vec3>* forces;
        Vec3* box;                       // three box vectors
        std::map<std::string, std::string> properties;
    };
private:
    std::string name;
};

} // namespace OpenMM
#endif
