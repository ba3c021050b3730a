// Test of the MPIDB200 platform kernel through the OpenMM object model (System / MPIDForce / Context / State):
// every quantity of the kernel contract is evaluated on the platform "MPIDB200" and on the reference's own
// "Reference" platform in the same process and compared.  Mirrors the structure of the reference's
// platforms/cuda/tests/TestCudaMPIDForce.cpp (System + MPIDForce + Context per case), but the expected values
// come from the Reference platform at run time instead of hard-coded numbers.
//
//   TestB200MPIDForce <waters.txt> [precision]
// waters.txt: "nWaters Lx Ly Lz" followed by 3*nWaters lines "x y z" (nm), O H H per molecule.
#include "openmm/Context.h"
#include "openmm/MPIDForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/System.h"
#include "openmm/VerletIntegrator.h"
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

using namespace OpenMM;

extern "C" void registerMPIDReferenceKernelFactories();
extern "C" void registerMPIDB200KernelFactories();

namespace {

struct Input { int nWaters; double L[3]; std::vector<Vec3> pos; };

// SWM6-MPID water (values of the reference's examples/parameters/swm6.xml:25-39 in API component order)
MPIDForce* makeWaterForce(const Input& in, MPIDForce::PolarizationType pol, double eps, bool anisotropic) {
    MPIDForce* f = new MPIDForce();
    f->setNonbondedMethod(MPIDForce::PME);
    f->setPolarizationType(pol);
    f->setCutoffDistance(0.8);
    f->setPMEParameters(3.2853, 32, 32, 32);
    f->setDefaultTholeWidth(8.0);
    f->setMutualInducedTargetEpsilon(eps);
    f->setMutualInducedMaxIterations(100);
    std::vector<double> zero3(3, 0.0), zero6(6, 0.0), zero10(10, 0.0);
    std::vector<double> dO(3, 0.0), qO(6, 0.0), oO(10, 0.0), aO(3, 0.00088);
    dO[2] = -0.023671684;
    qO[0] = 0.000150963; qO[2] = 0.00008707; qO[5] = -0.000238034;
    oO[4] = 0.000000426; oO[6] = 0.000000853; oO[9] = -0.000001279;
    if (anisotropic) { aO[0] = 0.00100024*0.88; aO[1] = 0.00125025*0.88; aO[2] = 0.00083350*0.88; }
    for (int w = 0; w < in.nWaters; w++) {
        int o = 3*w;
        f->addMultipole(-1.0614, dO, qO, oO, MPIDForce::Bisector, o+1, o+2, -1, 8.0, aO);
        f->addMultipole(0.5307, zero3, zero6, zero10, MPIDForce::ZThenX, o, o+2, -1, 0.0, zero3);
        f->addMultipole(0.5307, zero3, zero6, zero10, MPIDForce::ZThenX, o, o+1, -1, 0.0, zero3);
        std::vector<int> c12o(2), c12h(1, o), c13a(1, o+2), c13b(1, o+1);
        c12o[0] = o+1; c12o[1] = o+2;
        f->setCovalentMap(o, MPIDForce::Covalent12, c12o);
        f->setCovalentMap(o+1, MPIDForce::Covalent12, c12h);
        f->setCovalentMap(o+2, MPIDForce::Covalent12, c12h);
        f->setCovalentMap(o+1, MPIDForce::Covalent13, c13a);
        f->setCovalentMap(o+2, MPIDForce::Covalent13, c13b);
    }
    return f;
}

double relErr(const std::vector<Vec3>& a, const std::vector<Vec3>& b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); i++) { Vec3 d = a[i] - b[i]; num += d.dot(d); den += b[i].dot(b[i]); }
    return den > 0 ? std::sqrt(num/den) : std::sqrt(num);
}

struct Case {
    System system;
    MPIDForce* force;
    VerletIntegrator integrator;
    Context* context;
    Case(const Input& in, MPIDForce::PolarizationType pol, double eps, bool aniso, const std::string& platform, const std::string& precision)
        : integrator(0.001), context(0) {
        for (int i = 0; i < 3*in.nWaters; i++) system.addParticle(i % 3 == 0 ? 15.999 : 1.008);
        system.setDefaultPeriodicBoxVectors(Vec3(in.L[0], 0, 0), Vec3(0, in.L[1], 0), Vec3(0, 0, in.L[2]));
        force = makeWaterForce(in, pol, eps, aniso);
        system.addForce(force);
        Platform& p = Platform::getPlatformByName(platform);
        if (platform == "MPIDB200") {
            std::map<std::string, std::string> props;
            props["Precision"] = precision;
            context = new Context(system, integrator, p, props);
        } else context = new Context(system, integrator, p);
        context->setPositions(in.pos);
    }
    ~Case() { delete context; }
};

int failures = 0;
void report(const char* what, const char* polName, double value, double tol) {
    bool ok = value <= tol;
    printf("%-34s %-13s %.3e  (tol %.1e)  %s\n", what, polName, value, tol, ok ? "ok" : "FAIL");
    if (!ok) failures++;
}

void compare(const Input& in, MPIDForce::PolarizationType pol, const char* polName, double eps, bool aniso, const std::string& precision) {
    Case ref(in, pol, eps, aniso, "Reference", precision), b200(in, pol, eps, aniso, "MPIDB200", precision);
    const double tol = precision == "double" ? 1e-8 : 1e-5;      // BASELINE.json north_star tolerances
    State sr = ref.context->getState(State::Forces | State::Energy);
    State sb = b200.context->getState(State::Forces | State::Energy);
    report("relative force error", polName, relErr(sb.getForces(), sr.getForces()), tol);
    report("relative energy error", polName, std::fabs(sb.getPotentialEnergy() - sr.getPotentialEnergy())/std::fabs(sr.getPotentialEnergy()), tol);
    // forces are added to the Context's (zeroed) force array, so a second evaluation gives the same State -- up to the
    // summation order of the floating-point grid reductions in the PME spread, which is not fixed run to run
    State sb2 = b200.context->getState(State::Forces | State::Energy);
    report("repeat evaluation force change", polName, relErr(sb2.getForces(), sb.getForces()), 0.1*tol);
    std::vector<Vec3> mr, mb;
    ref.force->getInducedDipoles(*ref.context, mr);
    b200.force->getInducedDipoles(*b200.context, mb);
    report("induced dipole error", polName, relErr(mb, mr), tol);
    ref.force->getLabFramePermanentDipoles(*ref.context, mr);
    b200.force->getLabFramePermanentDipoles(*b200.context, mb);
    report("lab permanent dipole error", polName, relErr(mb, mr), 1e-12);
    ref.force->getTotalDipoles(*ref.context, mr);
    b200.force->getTotalDipoles(*b200.context, mb);
    report("total dipole error", polName, relErr(mb, mr), tol);
    double a1, a2; int g1[3], g2[3];
    ref.force->getPMEParametersInContext(*ref.context, a1, g1[0], g1[1], g1[2]);
    b200.force->getPMEParametersInContext(*b200.context, a2, g2[0], g2[1], g2[2]);
    report("PME parameters in context", polName, std::fabs(a1 - a2) + std::abs(g1[0]-g2[0]) + std::abs(g1[1]-g2[1]) + std::abs(g1[2]-g2[2]), 0.0);
    std::vector<double> momR, momB;
    ref.force->getSystemMultipoleMoments(*ref.context, momR);
    b200.force->getSystemMultipoleMoments(*b200.context, momB);
    double dm = 0, nm = 0;
    for (size_t k = 0; k < momR.size(); k++) { dm += (momR[k]-momB[k])*(momR[k]-momB[k]); nm += momR[k]*momR[k]; }
    report("system multipole moments", polName, std::sqrt(dm/nm), 10*tol);
    std::vector<Vec3> grid;
    grid.push_back(Vec3(0.11, 0.23, 0.37)); grid.push_back(Vec3(1.5, 1.4, 1.3)); grid.push_back(Vec3(2.9, 0.2, 3.0));
    std::vector<double> potR, potB;
    ref.force->getElectrostaticPotential(grid, *ref.context, potR);
    b200.force->getElectrostaticPotential(grid, *b200.context, potB);
    double dp = 0, np = 0;
    for (size_t k = 0; k < potR.size(); k++) { dp += (potR[k]-potB[k])*(potR[k]-potB[k]); np += potR[k]*potR[k]; }
    report("electrostatic potential", polName, std::sqrt(dp/np), 10*tol);
}

void errorBehaviour(const Input& in) {
    // MPIDForceImpl's own validation (cutoff vs. box) must fire before the kernel is reached, and the kernel's
    // "not using PME" error must read as the reference's (MPIDReferenceKernels.cpp:381-382)
    {
        System system;
        for (int i = 0; i < 3; i++) system.addParticle(1.0);
        MPIDForce* f = new MPIDForce();
        std::vector<double> z3(3, 0.0), z6(6, 0.0), z10(10, 0.0);
        for (int i = 0; i < 3; i++) f->addMultipole(i == 0 ? -1.0 : 0.5, z3, z6, z10, MPIDForce::NoAxisType, -1, -1, -1, 0.0, z3);
        system.addForce(f);
        VerletIntegrator integ(0.001);
        Context context(system, integ, Platform::getPlatformByName("MPIDB200"));
        std::vector<Vec3> pos(3);
        pos[0] = Vec3(0, 0, 0); pos[1] = Vec3(0.1, 0, 0); pos[2] = Vec3(0, 0.1, 0);
        context.setPositions(pos);
        bool threw = false; std::string msg;
        try { double a; int x, y, z; f->getPMEParametersInContext(context, a, x, y, z); }
        catch (const OpenMMException& e) { threw = true; msg = e.what(); }
        bool ok = threw && msg == "getPMEParametersInContext: This Context is not using PME";
        printf("%-34s %-13s %s\n", "error: PME query without PME", "", ok ? "ok" : "FAIL");
        if (!ok) failures++;
    }
    {
        // non-convergence must surface as an OpenMMException like the Reference platform's (MPIDReferenceForce.cpp:1222-1228)
        Input small = in;
        Case c(small, MPIDForce::Mutual, 1e-30, false, "MPIDB200", "mixed");
        c.force->setMutualInducedMaxIterations(3);
        c.context->reinitialize(true);
        bool threw = false; std::string msg;
        try { c.context->getState(State::Energy); }
        catch (const OpenMMException& e) { threw = true; msg = e.what(); }
        bool ok = threw && msg.find("Induced dipoles did not converge") == 0;
        printf("%-34s %-13s %s\n", "error: solver non-convergence", "", ok ? "ok" : "FAIL");
        if (!ok) failures++;
    }
}

void updateParameters(const Input& in, const std::string& precision) {
    // MPIDForce::updateParametersInContext -> copyParametersToContext (mpidKernels.h:96): scale every charge by 0.9
    // and compare with a fresh Reference-platform Context built from the scaled force.  (The Reference platform's
    // own copyParametersToContext is not used as an oracle: it writes dipoles and quadrupoles into the octopole
    // array, MPIDReferenceKernels.cpp:371-376.)
    Case b200(in, MPIDForce::Direct, 1e-5, false, "MPIDB200", precision);
    b200.context->getState(State::Energy);
    Case ref(in, MPIDForce::Direct, 1e-5, false, "Reference", precision);
    for (int i = 0; i < 3*in.nWaters; i++) {
        double q, th; int ax, z, x, y; std::vector<double> d, qd, o, a;
        b200.force->getMultipoleParameters(i, q, d, qd, o, ax, z, x, y, th, a);
        b200.force->setMultipoleParameters(i, 0.9*q, d, qd, o, ax, z, x, y, th, a);
        ref.force->setMultipoleParameters(i, 0.9*q, d, qd, o, ax, z, x, y, th, a);
    }
    b200.force->updateParametersInContext(*b200.context);
    ref.context->reinitialize(true);
    State sb = b200.context->getState(State::Forces | State::Energy);
    State sr = ref.context->getState(State::Forces | State::Energy);
    report("updateParametersInContext forces", "Direct", relErr(sb.getForces(), sr.getForces()), precision == "double" ? 1e-8 : 1e-5);
}

} // namespace

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s waters.txt [mixed|double]\n", argv[0]); return 2; }
    std::string precision = argc > 2 ? argv[2] : "mixed";
    try {
        Input in;
        std::ifstream f(argv[1]);
        if (!f) throw OpenMMException(std::string("cannot open ") + argv[1]);
        f >> in.nWaters >> in.L[0] >> in.L[1] >> in.L[2];
        in.pos.resize(3*in.nWaters);
        for (Vec3& p : in.pos) f >> p[0] >> p[1] >> p[2];
        registerMPIDReferenceKernelFactories();
        registerMPIDB200KernelFactories();
        printf("TestB200MPIDForce: %d waters, precision %s\n", in.nWaters, precision.c_str());
        compare(in, MPIDForce::Direct, "Direct", 1e-5, false, precision);
        compare(in, MPIDForce::Extrapolated, "Extrapolated", 1e-5, false, precision);
        compare(in, MPIDForce::Mutual, "Mutual", 1e-9, false, precision);
        compare(in, MPIDForce::Mutual, "Mutual-aniso", 1e-9, true, precision);
        updateParameters(in, precision);
        errorBehaviour(in);
    } catch (const std::exception& e) {
        printf("exception: %s\nFAIL\n", e.what());
        return 1;
    }
    printf(failures == 0 ? "Done\n" : "FAIL (%d checks)\n", failures);
    return failures == 0 ? 0 : 1;
}
