// TEST INFRASTRUCTURE ONLY -- parity oracle driver.
//
// Flat C entry points around the reference's own, unmodified stack
//   MPIDForce -> MPIDForceImpl -> ReferenceCalcMPIDForceKernel -> MPIDReference[Pme]Force
// (reference: openmmapi/src/MPIDForce.cpp, openmmapi/src/MPIDForceImpl.cpp:51-159,
//  platforms/reference/src/MPIDReferenceKernels.cpp:84-258), hosted by the repo's openmm-compat
// runtime.  Built by oracle/Makefile into oracle/_ref/libmpidref.so.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may load it.
#include "openmm/Context.h"
#include "openmm/MPIDForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/VerletIntegrator.h"
#include <cstring>
#include <string>
#include <vector>

using namespace OpenMM;

extern "C" void registerMPIDReferenceKernelFactories();

namespace {
std::string lastError;
struct RefHandle {
    System system;
    MPIDForce* force;      // owned by system
    VerletIntegrator integrator;
    Context* context;
    int n;
    RefHandle() : force(0), integrator(0.001), context(0), n(0) {}
    ~RefHandle() { delete context; }
};
bool registered = false;
}

extern "C" {

const char* mpidref_last_error() { return lastError.c_str(); }

// Covalent maps arrive as CSR: for type t (0..7, MPIDForce::CovalentType order) atom i owns
// cov_indices[cov_offsets[t*(n+1)+i] .. cov_offsets[t*(n+1)+i+1]).
int mpidref_create(int n,
                   const double* charges, const double* dipoles, const double* quadrupoles, const double* octopoles,
                   const int* axisTypes, const int* atomZ, const int* atomX, const int* atomY,
                   const double* tholes, const double* alphas,
                   const int* cov_offsets, const int* cov_indices,
                   int method, int polarization, double cutoff, double ewaldAlpha, int nx, int ny, int nz,
                   double ewaldTol, double defaultThole, double scale14, int maxIter, double epsilon,
                   int ncoef, const double* coefs, const double* box9, void** out) {
    try {
        if (!registered) { registerMPIDReferenceKernelFactories(); registered = true; }
        RefHandle* h = new RefHandle();
        h->n = n;
        h->force = new MPIDForce();
        for (int i = 0; i < n; i++) {
            std::vector<double> d(dipoles+3*i, dipoles+3*i+3), q(quadrupoles+6*i, quadrupoles+6*i+6),
                                o(octopoles+10*i, octopoles+10*i+10), a(alphas+3*i, alphas+3*i+3);
            h->force->addMultipole(charges[i], d, q, o, axisTypes[i], atomZ[i], atomX[i], atomY[i], tholes[i], a);
            h->system.addParticle(1.0);
        }
        for (int t = 0; t < 8; t++)
            for (int i = 0; i < n; i++) {
                int b = cov_offsets[t*(n+1)+i], e = cov_offsets[t*(n+1)+i+1];
                if (e > b) {
                    std::vector<int> lst(cov_indices+b, cov_indices+e);
                    h->force->setCovalentMap(i, (MPIDForce::CovalentType) t, lst);
                }
            }
        h->force->setNonbondedMethod((MPIDForce::NonbondedMethod) method);
        h->force->setPolarizationType((MPIDForce::PolarizationType) polarization);
        h->force->setCutoffDistance(cutoff);
        h->force->setPMEParameters(ewaldAlpha, nx, ny, nz);
        h->force->setEwaldErrorTolerance(ewaldTol);
        h->force->setDefaultTholeWidth(defaultThole);
        h->force->set14ScaleFactor(scale14);
        h->force->setMutualInducedMaxIterations(maxIter);
        h->force->setMutualInducedTargetEpsilon(epsilon);
        if (ncoef > 0) h->force->setExtrapolationCoefficients(std::vector<double>(coefs, coefs+ncoef));
        h->system.setDefaultPeriodicBoxVectors(Vec3(box9[0], box9[1], box9[2]), Vec3(box9[3], box9[4], box9[5]),
                                               Vec3(box9[6], box9[7], box9[8]));
        h->system.addForce(h->force);
        h->context = new Context(h->system, h->integrator, Platform::getPlatformByName("Reference"));
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

int mpidref_set_box(void* handle, const double* box9) {
    try {
        RefHandle* h = static_cast<RefHandle*>(handle);
        h->context->setPeriodicBoxVectors(Vec3(box9[0], box9[1], box9[2]), Vec3(box9[3], box9[4], box9[5]),
                                          Vec3(box9[6], box9[7], box9[8]));
        return 0;
    } catch (const std::exception& e) { lastError = e.what(); return 1; }
}

static void loadPositions(RefHandle* h, const double* pos) {
    std::vector<Vec3> p(h->n);
    for (int i = 0; i < h->n; i++) p[i] = Vec3(pos[3*i], pos[3*i+1], pos[3*i+2]);
    h->context->setPositions(p);
}

int mpidref_execute(void* handle, const double* pos, double* energy, double* forces) {
    try {
        RefHandle* h = static_cast<RefHandle*>(handle);
        loadPositions(h, pos);
        State s = h->context->getState(State::Forces | State::Energy);
        *energy = s.getPotentialEnergy();
        if (forces)
            for (int i = 0; i < h->n; i++)
                for (int k = 0; k < 3; k++) forces[3*i+k] = s.getForces()[i][k];
        return 0;
    } catch (const std::exception& e) { lastError = e.what(); return 1; }
}

// which: 0 induced, 1 lab-frame permanent, 2 total
int mpidref_get_dipoles(void* handle, const double* pos, int which, double* out) {
    try {
        RefHandle* h = static_cast<RefHandle*>(handle);
        loadPositions(h, pos);
        std::vector<Vec3> d;
        if (which == 0) h->force->getInducedDipoles(*h->context, d);
        else if (which == 1) h->force->getLabFramePermanentDipoles(*h->context, d);
        else h->force->getTotalDipoles(*h->context, d);
        for (int i = 0; i < h->n; i++)
            for (int k = 0; k < 3; k++) out[3*i+k] = d[i][k];
        return 0;
    } catch (const std::exception& e) { lastError = e.what(); return 1; }
}

int mpidref_get_pme_parameters(void* handle, double* alpha, int* nx, int* ny, int* nz) {
    try {
        RefHandle* h = static_cast<RefHandle*>(handle);
        h->force->getPMEParametersInContext(*h->context, *alpha, *nx, *ny, *nz);
        return 0;
    } catch (const std::exception& e) { lastError = e.what(); return 1; }
}

int mpidref_get_system_multipole_moments(void* handle, const double* pos, double* out13) {
    try {
        RefHandle* h = static_cast<RefHandle*>(handle);
        loadPositions(h, pos);
        std::vector<double> m;
        h->force->getSystemMultipoleMoments(*h->context, m);
        for (size_t i = 0; i < m.size() && i < 13; i++) out13[i] = m[i];
        return 0;
    } catch (const std::exception& e) { lastError = e.what(); return 1; }
}

void mpidref_destroy(void* handle) { delete static_cast<RefHandle*>(handle); }

} // extern "C"
