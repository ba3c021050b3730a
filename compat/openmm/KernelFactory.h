#ifndef OPENMM_COMPAT_KERNELFACTORY_H_
#define OPENMM_COMPAT_KERNELFACTORY_H_
#include "openmm/KernelImpl.h"
namespace OpenMM {
class ContextImpl;
class OPENMM_EXPORT KernelFactory {
public:
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
    virtual ~KernelFactory() {}
};
} // namespace OpenMM
#endif
