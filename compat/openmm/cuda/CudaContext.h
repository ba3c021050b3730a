// Stand-in for the handful of OpenMM CUDA-platform declarations the MPIDB200 CUDA-platform binding touches
// (openmm/platforms/cuda/include: CudaContext.h, CudaArray.h, CudaPlatform.h, CudaForceInfo.h -- OpenMM 7.x/8.x).
// COMPILE-CHECK ONLY: it lets mpidopenmmplugin_b200/plugin/src/MPIDB200CudaPlatformKernel.cpp be type-checked in a
// container without OpenMM; nothing here runs.  Signatures follow the ones the reference's own CUDA platform uses
// (platforms/cuda/src/MPIDCudaKernels.cpp:216, 269, 659-661, 1089; MPIDCudaKernelFactory.cpp:68-75).
#ifndef OPENMM_COMPAT_CUDACONTEXT_H_
#define OPENMM_COMPAT_CUDACONTEXT_H_
#include "openmm/Platform.h"
#include "openmm/Vec3.h"
#include <string>
#include <vector>
typedef unsigned long long CUdeviceptr;
typedef struct CUstream_st* CUstream;
namespace OpenMM {
class CudaArray {
public:
    CUdeviceptr& getDevicePointer() { return ptr; }
    int getSize() const { return 0; }
private:
    CUdeviceptr ptr = 0;
};
class CudaForceInfo {
public:
    virtual ~CudaForceInfo() {}
    virtual bool areParticlesIdentical(int, int) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int, std::vector<int>&) {}
    virtual bool areGroupsIdentical(int, int) { return true; }
};
class CudaContext {
public:
    void setAsCurrent() {}
    int getNumAtoms() const { return 0; }
    int getPaddedNumAtoms() const { return 0; }
    int getDeviceIndex() const { return 0; }
    bool getUseDoublePrecision() const { return false; }
    bool getUseMixedPrecision() const { return true; }
    CudaArray& getPosq() { return posq; }
    CudaArray& getPosqCorrection() { return posqCorrection; }
    CudaArray& getAtomIndexArray() { return atomIndex; }
    CudaArray& getForce() { return force; }
    const std::vector<int>& getAtomIndex() const { return order; }
    CUstream getCurrentStream() { return 0; }
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = Vec3(); b = Vec3(); c = Vec3(); }
    void addForce(CudaForceInfo* info) { delete info; }
private:
    CudaArray posq, posqCorrection, atomIndex, force;
    std::vector<int> order;
};
class CudaPlatform : public Platform {
public:
    class PlatformData { public: std::vector<CudaContext*> contexts; };
};
} // namespace OpenMM
#endif
