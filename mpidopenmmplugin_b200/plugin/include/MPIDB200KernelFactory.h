// KernelFactory of the MPIDB200 platform (counterpart of MPIDCudaKernelFactory,
// reference: platforms/cuda/src/MPIDCudaKernelFactory.h / .cpp:68-76).
#ifndef MPIDB200_KERNEL_FACTORY_H_
#define MPIDB200_KERNEL_FACTORY_H_

#include "openmm/KernelFactory.h"
#include <string>

namespace OpenMM {

class MPIDB200KernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const;
};

} // namespace OpenMM

extern "C" void registerPlatforms();
extern "C" void registerKernelFactories();
extern "C" void registerMPIDB200KernelFactories();
extern "C" void registerMPIDB200KernelFactoriesAs(const char* platformName);

#endif
