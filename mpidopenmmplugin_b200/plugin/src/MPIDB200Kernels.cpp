// Host side of the MPIDB200 platform kernel: flattens an MPIDForce into the C-ABI arrays and moves
// positions / forces between the Context and the engine.  Behaviour follows the reference's platform
// glue (platforms/reference/src/MPIDReferenceKernels.cpp:84-388, platforms/cuda/src/MPIDCudaKernels.cpp:
// 200-460 for what is uploaded once and what per call); the arithmetic is all behind mpidb200_*.
#include "MPIDB200Kernels.h"
#include "openmm/MPIDForce.h"
#include "openmm/OpenMMException.h"
#include <cstdlib>
#include <cstring>
#include <sstream>

using namespace OpenMM;

namespace {

MPIDB200Platform::PlatformData& stateOf(ContextImpl& context) {
    return *static_cast<MPIDB200Platform::PlatformData*>(context.getPlatformData());
}

int precisionFromProperty(const std::string& value) {
    // OpenMM's CUDA platform accepts "single", "mixed" and "double"; the engine's FP32-pair-math mode
    // already accumulates in FP64 / 64-bit fixed point, so "single" and "mixed" select the same path.
    if (value == "double") return MPIDB200_DOUBLE;
    if (value == "single" || value == "mixed") return MPIDB200_MIXED;
    throw OpenMMException("Illegal value for Precision: " + value);
}

} // namespace

B200CalcMPIDForceKernel::B200CalcMPIDForceKernel(std::string name, const Platform& platform, const System& system, ContextImpl& context)
    : CalcMPIDForceKernel(name, platform), system(system), owner(context), engine(0), numMultipoles(0), usePme(false), haveBox(false),
      pinnedPos(0), pinnedForce(0), pinnedPosSize(0), pinnedForceSize(0) {
    memset(lastBox, 0, sizeof(lastBox));
}

B200CalcMPIDForceKernel::~B200CalcMPIDForceKernel() {
    if (engine) mpidb200_destroy(engine);
}

void B200CalcMPIDForceKernel::check(int status) const {
    if (status != 0) throw OpenMMException(mpidb200_last_error());
}

void B200CalcMPIDForceKernel::initialize(const System& sys, const MPIDForce& force) {
    const MPIDB200Platform::PlatformData& state = stateOf(owner);
    auto prop = [&](const std::string& key) -> std::string {
        std::map<std::string, std::string>::const_iterator it = state.properties.find(key);
        return it != state.properties.end() ? it->second : getPlatform().getPropertyDefaultValue(key);
    };
    initializeOn(sys, force, precisionFromProperty(prop(MPIDB200Platform::Precision())), atoi(prop(MPIDB200Platform::DeviceIndex()).c_str()),
                 prop(MPIDB200Platform::Solver()) == "CG" ? MPIDB200_SOLVER_CG : MPIDB200_SOLVER_DIIS);
}

void B200CalcMPIDForceKernel::initializeOn(const System& sys, const MPIDForce& force, int precision, int device, int solver) {
    numMultipoles = force.getNumMultipoles();
    if (numMultipoles != sys.getNumParticles())
        throw OpenMMException("MPIDForce must have exactly as many particles as the System it belongs to.");

    mpidb200_config cfg;
    mpidb200_default_config(&cfg);
    cfg.num_particles = numMultipoles;
    cfg.nonbonded_method = force.getNonbondedMethod() == MPIDForce::PME ? MPIDB200_PME : MPIDB200_NOCUTOFF;
    switch (force.getPolarizationType()) {
        case MPIDForce::Mutual:       cfg.polarization_type = MPIDB200_MUTUAL; break;
        case MPIDForce::Direct:       cfg.polarization_type = MPIDB200_DIRECT; break;
        case MPIDForce::Extrapolated: cfg.polarization_type = MPIDB200_EXTRAPOLATED; break;
        default: throw OpenMMException("MPIDForce: unknown polarization type");
    }
    usePme = cfg.nonbonded_method == MPIDB200_PME;
    cfg.cutoff = force.getCutoffDistance();
    // explicit PME parameters when the user set them, otherwise 0 -> the engine applies the rule the
    // reference delegates to NonbondedForceImpl::calcPMEParameters (MPIDReferenceKernels.cpp:161-170)
    force.getPMEParameters(cfg.ewald_alpha, cfg.grid[0], cfg.grid[1], cfg.grid[2]);
    if (cfg.grid[0] == 0 || cfg.ewald_alpha == 0.0) { cfg.ewald_alpha = 0.0; cfg.grid[0] = cfg.grid[1] = cfg.grid[2] = 0; }
    cfg.ewald_tolerance = force.getEwaldErrorTolerance();
    cfg.default_thole_width = force.getDefaultTholeWidth();
    cfg.scale14 = force.get14ScaleFactor();
    cfg.max_iterations = force.getMutualInducedMaxIterations();
    cfg.target_epsilon = force.getMutualInducedTargetEpsilon();
    const std::vector<double>& coefs = force.getExtrapolationCoefficients();
    if (coefs.size() > 8) throw OpenMMException("MPIDForce: at most 8 extrapolation coefficients are supported");
    cfg.num_extrapolation_coefficients = (int) coefs.size();
    for (size_t k = 0; k < coefs.size(); k++) cfg.extrapolation_coefficients[k] = coefs[k];

    cfg.precision = precision;
    cfg.device = device;
    cfg.solver = solver;

    check(mpidb200_create(&cfg, &engine));
    uploadParticles(force);

    // covalent maps -> CSR per CovalentType (MPIDForce::getCovalentMaps, MPIDForce.h:308)
    const int numTypes = (int) MPIDForce::CovalentEnd;
    std::vector<std::vector<std::vector<int> > > maps(numMultipoles);
    for (int i = 0; i < numMultipoles; i++) force.getCovalentMaps(i, maps[i]);
    std::vector<int> offsets((size_t) numTypes*(numMultipoles + 1), 0), indices;
    for (int t = 0; t < numTypes; t++) {
        for (int i = 0; i < numMultipoles; i++) {
            offsets[(size_t) t*(numMultipoles + 1) + i] = (int) indices.size();
            if (t < (int) maps[i].size()) indices.insert(indices.end(), maps[i][t].begin(), maps[i][t].end());
        }
        offsets[(size_t) t*(numMultipoles + 1) + numMultipoles] = (int) indices.size();
    }
    if (indices.empty()) indices.push_back(0);
    check(mpidb200_set_covalent_maps(engine, offsets.data(), indices.data()));
}

void B200CalcMPIDForceKernel::uploadParticles(const MPIDForce& force) {
    const int n = numMultipoles;
    std::vector<double> charges(n), dipoles(3*(size_t) n), quadrupoles(6*(size_t) n), octopoles(10*(size_t) n), tholes(n), alphas(3*(size_t) n);
    std::vector<int> axis(n), az(n), ax(n), ay(n);
    std::vector<double> d, q, o, a;
    for (int i = 0; i < n; i++) {
        force.getMultipoleParameters(i, charges[i], d, q, o, axis[i], az[i], ax[i], ay[i], tholes[i], a);
        for (int k = 0; k < 3; k++)  dipoles[3*(size_t) i + k] = d[k];
        for (int k = 0; k < 6; k++)  quadrupoles[6*(size_t) i + k] = q[k];
        for (int k = 0; k < 10; k++) octopoles[10*(size_t) i + k] = o[k];
        for (int k = 0; k < 3; k++)  alphas[3*(size_t) i + k] = a[k];
    }
    check(mpidb200_set_particles(engine, charges.data(), dipoles.data(), quadrupoles.data(), octopoles.data(),
                                 axis.data(), az.data(), ax.data(), ay.data(), tholes.data(), alphas.data()));
}

// Box vectors are read from the Context on every call (they change under a barostat); the engine only
// rebuilds its reciprocal-space tables when they actually differ (MPIDReferenceKernels.cpp:188-198).
void B200CalcMPIDForceKernel::syncBox(ContextImpl& context) {
    if (!usePme) return;
    const Vec3* box = stateOf(context).box;
    syncBoxVectors(box[0], box[1], box[2]);
}
void B200CalcMPIDForceKernel::syncBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) {
    if (!usePme) return;
    const Vec3 box[3] = {a, b, c};
    double flat[9];
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) flat[3*r + cc] = box[r][cc];
    if (haveBox && memcmp(flat, lastBox, sizeof(flat)) == 0) return;
    check(mpidb200_set_box(engine, flat, flat + 3, flat + 6));
    memcpy(lastBox, flat, sizeof(flat));
    haveBox = true;
}

const double* B200CalcMPIDForceKernel::flatPositions(ContextImpl& context) {
    const std::vector<Vec3>& pos = *stateOf(context).positions;
    posFlat.resize(3*pos.size());
    for (size_t i = 0; i < pos.size(); i++) { posFlat[3*i] = pos[i][0]; posFlat[3*i+1] = pos[i][1]; posFlat[3*i+2] = pos[i][2]; }
    return posFlat.data();
}

double B200CalcMPIDForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    syncBox(context);
    const double* pos = flatPositions(context);
    double energy = 0.0;
    std::vector<Vec3>& forces = *stateOf(context).forces;
    if (includeForces) forceFlat.assign(3*forces.size(), 0.0);
    // the two staging vectors live as long as the kernel: page-lock them once so that the engine moves them by DMA and
    // accumulates the forces on the device (mpidb200_pin_host_buffer)
    pinIfMoved(posFlat, pinnedPos, pinnedPosSize);
    if (includeForces) pinIfMoved(forceFlat, pinnedForce, pinnedForceSize);
    check(mpidb200_execute(engine, pos, includeForces ? 1 : 0, includeEnergy ? 1 : 0, &energy, includeForces ? forceFlat.data() : 0));
    if (includeForces)      // forces are accumulated, never overwritten (MPIDReferenceKernels.cpp:229-238)
        for (size_t i = 0; i < forces.size(); i++)
            forces[i] += Vec3(forceFlat[3*i], forceFlat[3*i+1], forceFlat[3*i+2]);
    return energy;
}

void B200CalcMPIDForceKernel::pinIfMoved(std::vector<double>& buffer, const double*& pinned, size_t& pinnedSize) {
    if (buffer.empty() || (pinned == buffer.data() && pinnedSize == buffer.size())) return;
    if (pinned) mpidb200_unpin_host_buffer(engine, (void*) pinned);
    // page-locking can be refused (limits on locked memory): the engine then stages the array itself
    if (mpidb200_pin_host_buffer(engine, buffer.data(), (unsigned long long) (buffer.size()*sizeof(double))) == 0) { pinned = buffer.data(); pinnedSize = buffer.size(); }
    else { pinned = 0; pinnedSize = 0; }
}

void B200CalcMPIDForceKernel::dipoleQuery(ContextImpl& context, int which, std::vector<Vec3>& out) {
    syncBox(context);
    const double* pos = flatPositions(context);
    std::vector<double> flat(3*(size_t) numMultipoles);
    check(mpidb200_get_dipoles(engine, pos, which, flat.data()));
    out.resize(numMultipoles);
    for (int i = 0; i < numMultipoles; i++) out[i] = Vec3(flat[3*i], flat[3*i+1], flat[3*i+2]);
}
void B200CalcMPIDForceKernel::getInducedDipoles(ContextImpl& context, std::vector<Vec3>& dipoles) { dipoleQuery(context, 0, dipoles); }
void B200CalcMPIDForceKernel::getLabFramePermanentDipoles(ContextImpl& context, std::vector<Vec3>& dipoles) { dipoleQuery(context, 1, dipoles); }
void B200CalcMPIDForceKernel::getTotalDipoles(ContextImpl& context, std::vector<Vec3>& dipoles) { dipoleQuery(context, 2, dipoles); }

void B200CalcMPIDForceKernel::getElectrostaticPotential(ContextImpl& context, const std::vector<Vec3>& inputGrid,
                                                        std::vector<double>& outputElectrostaticPotential) {
    syncBox(context);
    const double* pos = flatPositions(context);
    std::vector<double> pts(3*inputGrid.size());
    for (size_t g = 0; g < inputGrid.size(); g++) { pts[3*g] = inputGrid[g][0]; pts[3*g+1] = inputGrid[g][1]; pts[3*g+2] = inputGrid[g][2]; }
    outputElectrostaticPotential.assign(inputGrid.size(), 0.0);
    if (inputGrid.empty()) return;
    check(mpidb200_get_electrostatic_potential(engine, pos, (int) inputGrid.size(), pts.data(), outputElectrostaticPotential.data()));
}

void B200CalcMPIDForceKernel::getSystemMultipoleMoments(ContextImpl& context, std::vector<double>& outputMultipoleMoments) {
    syncBox(context);
    const double* pos = flatPositions(context);
    std::vector<double> masses(numMultipoles);
    for (int i = 0; i < numMultipoles; i++) masses[i] = context.getSystem().getParticleMass(i);
    outputMultipoleMoments.assign(13, 0.0);
    check(mpidb200_get_system_multipole_moments(engine, pos, masses.data(), outputMultipoleMoments.data()));
}

void B200CalcMPIDForceKernel::copyParametersToContext(ContextImpl& context, const MPIDForce& force) {
    if (numMultipoles != force.getNumMultipoles())
        throw OpenMMException("updateParametersInContext: The number of multipoles has changed");
    uploadParticles(force);
}

void B200CalcMPIDForceKernel::getPMEParameters(double& alpha, int& nx, int& ny, int& nz) const {
    if (!usePme) throw OpenMMException("getPMEParametersInContext: This Context is not using PME");
    if (!haveBox) {
        // no evaluation yet: the engine needs the box to size an automatic grid
        const_cast<B200CalcMPIDForceKernel*>(this)->syncBox(owner);
    }
    check(mpidb200_get_pme_parameters(engine, &alpha, &nx, &ny, &nz));
}

void B200CalcMPIDForceKernel::getSolverStatistics(int& iterations, double& epsilon) const {
    check(mpidb200_get_stats(engine, &iterations, &epsilon, 0, 0));
}
