// B200CalcMPIDForceKernel: the MPIDB200 implementation of the plugin's kernel contract
//   class CalcMPIDForceKernel   (reference: openmmapi/include/openmm/mpidKernels.h:50-104)
// It is created by MPIDForceImpl::initialize through Platform::createKernel("CalcMPIDForce")
// (reference: openmmapi/src/MPIDForceImpl.cpp:151-152) exactly like the reference's
// ReferenceCalcMPIDForceKernel / CudaCalcMPIDForceKernel, and forwards every call to the device engine
// through the C ABI in include/mpidb200.h.  No arithmetic lives in this file's translation unit.
#ifndef MPIDB200_KERNELS_H_
#define MPIDB200_KERNELS_H_

#include "openmm/mpidKernels.h"
#include "openmm/System.h"
#include "MPIDB200Platform.h"
#include "mpidb200.h"
#include <vector>

namespace OpenMM {

class B200CalcMPIDForceKernel : public CalcMPIDForceKernel {
public:
    B200CalcMPIDForceKernel(std::string name, const Platform& platform, const System& system, ContextImpl& context);
    ~B200CalcMPIDForceKernel();
    void initialize(const System& system, const MPIDForce& force);
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy);
    void getLabFramePermanentDipoles(ContextImpl& context, std::vector<Vec3>& dipoles);
    void getInducedDipoles(ContextImpl& context, std::vector<Vec3>& dipoles);
    void getTotalDipoles(ContextImpl& context, std::vector<Vec3>& dipoles);
    void getElectrostaticPotential(ContextImpl& context, const std::vector<Vec3>& inputGrid,
                                   std::vector<double>& outputElectrostaticPotential);
    void getSystemMultipoleMoments(ContextImpl& context, std::vector<double>& outputMultipoleMoments);
    void copyParametersToContext(ContextImpl& context, const MPIDForce& force);
    void getPMEParameters(double& alpha, int& nx, int& ny, int& nz) const;

    // engine statistics of the last execute (iterations of the mutual solver, final epsilon)
    void getSolverStatistics(int& iterations, double& epsilon) const;
protected:
    // shared with the CUDA-platform binding (src/MPIDB200CudaPlatformKernel.cpp)
    void check(int status) const;                       // C-ABI status -> OpenMMException
    void initializeOn(const System& system, const MPIDForce& force, int precision, int device, int solver);
    void syncBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c);
    mpidb200_handle engineHandle() const { return engine; }
private:
    void uploadParticles(const MPIDForce& force);
    void syncBox(ContextImpl& context);
    void pinIfMoved(std::vector<double>& buffer, const double*& pinned, size_t& pinnedSize);
    const double* flatPositions(ContextImpl& context);
    void dipoleQuery(ContextImpl& context, int which, std::vector<Vec3>& out);

    const System& system;
    ContextImpl& owner;
    mpidb200_handle engine;
    int numMultipoles;
    bool usePme;
    double lastBox[9];
    bool haveBox;
    std::vector<double> posFlat, forceFlat;
    const double* pinnedPos; const double* pinnedForce; size_t pinnedPosSize, pinnedForceSize;
};

} // namespace OpenMM
#endif
