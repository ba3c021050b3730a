// MPIDB200 on OpenMM's CUDA platform: the same engine, bound to the device-resident data of a CudaContext.
//
// Built only with -DMPIDB200_ON_OPENMM_CUDA against a real OpenMM install (it needs CudaContext.h / CudaPlatform.h);
// in this repository it is TYPE-CHECKED against the stand-in compat/openmm/cuda/CudaContext.h (plugin/Makefile, target
// cuda-variant-check) and its device path -- mpidb200_execute_cuda_context with its two conversion kernels -- is
// exercised on a GPU by tests/test_gpu_parity.py::test_cuda_context_entry_matches_the_host_entry.
//
// What it replaces in the reference: platforms/cuda/src/MPIDCudaKernelFactory.cpp:36-76 (registration on the "CUDA"
// platform, contexts[0] only) and the data plumbing of CudaCalcMPIDForceKernel (platforms/cuda/src/MPIDCudaKernels.cpp:
// 61-104 ForceInfo, 216 posq, 659-661 addForce, 1089 atom order).  Positions are read from cu.getPosq() in the context's
// reordered atom order, forces are added to cu.getForce() in its 64-bit fixed-point layout, everything runs on
// cu.getCurrentStream(); nothing but the energy crosses the PCIe bus.
#ifdef MPIDB200_ON_OPENMM_CUDA
#include "MPIDB200Kernels.h"
#include "openmm/MPIDForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/windowsExport.h"
#include "CudaContext.h"
#include "CudaPlatform.h"
#include "CudaForceInfo.h"
#include <cstring>

using namespace OpenMM;

namespace {

// Which particles the CudaContext may treat as interchangeable when it reorders molecules: same rule as the reference
// (MPIDCudaKernels.cpp:61-104) -- identical multipole parameters, and the covalent maps as particle groups.
class MPIDB200ForceInfo : public CudaForceInfo {
public:
    explicit MPIDB200ForceInfo(const MPIDForce& force) : force(force) {}
    bool areParticlesIdentical(int p1, int p2) {
        double c1, c2, t1, t2;
        int a1, a2, z1, z2, x1, x2, y1, y2;
        std::vector<double> d1, d2, q1, q2, o1, o2, al1, al2;
        force.getMultipoleParameters(p1, c1, d1, q1, o1, a1, z1, x1, y1, t1, al1);
        force.getMultipoleParameters(p2, c2, d2, q2, o2, a2, z2, x2, y2, t2, al2);
        return c1 == c2 && t1 == t2 && a1 == a2 && al1 == al2 && d1 == d2 && q1 == q2 && o1 == o2;
    }
    int getNumParticleGroups() { return 7*force.getNumMultipoles(); }
    void getParticlesInGroup(int index, std::vector<int>& particles) {
        const int particle = index/7, type = index - 7*particle;
        force.getCovalentMap(particle, MPIDForce::CovalentType(type), particles);
    }
    bool areGroupsIdentical(int g1, int g2) { return (g1 % 7) == (g2 % 7); }
private:
    const MPIDForce& force;
};

class CudaB200CalcMPIDForceKernel : public B200CalcMPIDForceKernel {
public:
    CudaB200CalcMPIDForceKernel(std::string name, const Platform& platform, CudaContext& cu, const System& system, ContextImpl& context)
        : B200CalcMPIDForceKernel(name, platform, system, context), cu(cu) {}
    void initialize(const System& system, const MPIDForce& force) {
        cu.setAsCurrent();
        // precision and device follow the CudaContext, not platform properties of our own
        initializeOn(system, force, cu.getUseDoublePrecision() ? MPIDB200_DOUBLE : MPIDB200_MIXED, cu.getDeviceIndex(), MPIDB200_SOLVER_DIIS);
        cu.addForce(new MPIDB200ForceInfo(force));
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
        cu.setAsCurrent();
        Vec3 a, b, c;
        cu.getPeriodicBoxVectors(a, b, c);
        syncBoxVectors(a, b, c);
        check(mpidb200_set_stream(engineHandle(), (void*) cu.getCurrentStream()));
        double energy = 0.0;
        const bool dbl = cu.getUseDoublePrecision();
        const void* correction = (!dbl && cu.getUseMixedPrecision()) ? (const void*) cu.getPosqCorrection().getDevicePointer() : 0;
        check(mpidb200_execute_cuda_context(engineHandle(), (const void*) cu.getPosq().getDevicePointer(), dbl ? 1 : 0, correction,
                                            (const int*) cu.getAtomIndexArray().getDevicePointer(), cu.getPaddedNumAtoms(),
                                            includeForces ? 1 : 0, includeEnergy ? 1 : 0, &energy,
                                            includeForces ? (void*) cu.getForce().getDevicePointer() : 0));
        return energy;
    }
private:
    CudaContext& cu;
};

class MPIDB200CudaKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
        CudaPlatform::PlatformData& data = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData());
        CudaContext& cu = *data.contexts[0];                 // like the reference: the first device context only
        if (name == CalcMPIDForceKernel::Name())
            return new CudaB200CalcMPIDForceKernel(name, platform, cu, context.getSystem(), context);
        throw OpenMMException(("Tried to create kernel with illegal kernel name '" + name + "'").c_str());
    }
};

} // namespace

// registers CalcMPIDForce on the existing "CUDA" platform (the library is then used INSTEAD of libMPIDPluginCUDA.so)
extern "C" OPENMM_EXPORT void registerMPIDB200OnCudaPlatform() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        platform.registerKernelFactory(CalcMPIDForceKernel::Name(), new MPIDB200CudaKernelFactory());
    }
    catch (...) {
        // no CUDA platform in this OpenMM: nothing to register
    }
}
#endif
