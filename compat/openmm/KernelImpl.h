#ifndef OPENMM_COMPAT_KERNELIMPL_H_
#define OPENMM_COMPAT_KERNELIMPL_H_
#include "openmm/internal/windowsExport.h"
#include <string>
namespace OpenMM {
class Platform;
class OPENMM_EXPORT KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform), referenceCount(0) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    friend class Kernel;
    std::string name;
    const Platform* platform;
    int referenceCount;
};
} // namespace OpenMM
#endif
