// TEST INFRASTRUCTURE ONLY -- cell-list driver around the reference's own PME pair functions.
//
// The Reference platform visits every (i, j) pair of the system in three O(N^2) loops
//   MPIDReferenceForce::calculateFixedMultipoleField          platforms/reference/src/SimTKReference/MPIDReferenceForce.cpp:910-933
//   MPIDReferencePmeForce::calculateInducedDipoleFields       :4073-4141  (pair loop :4084-4088)
//   MPIDReferencePmeForce::calculateElectrostatic             :4922-4990  (pair loop :4932-4946)
// and lets the pair functions (calculateFixedMultipoleFieldPairIxn :2812, calculateDirectInducedDipolePairIxns :4161,
// calculatePmeDirectElectrostaticPairIxn :4335) reject the pairs beyond the cutoff.  That makes the 95,616-atom and
// 1,024,884-atom boxes of BASELINE.json unreachable (~25 min and ~2 days per evaluation).  This file derives a class
// from MPIDReferencePmeForce that overrides those three virtual functions and calls the SAME pair functions (and the same
// reciprocal-space / self-term members, in the same order) for a candidate list that is a superset of the in-cutoff
// pairs, enumerated in the stock loops' order (i ascending, j > i ascending).  Every pair the stock loops would have
// passed the cutoff test is visited, in the same order, by the same code: with one thread the result is bit-identical
// to the stock loops (tests/test_oracle_cell.py checks that against oracle/_ref/libmpidref.so).  With T threads the
// rows are split into T ranges with private accumulators that are summed in range order (deterministic; differs from
// the stock order by round-off only).
//
// Nothing is copied from the reference: its header is included where it lies, with `private` opened to `protected`
// so that the derived class can reach the PME members (the reference declares them private).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the resulting library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include "openmm/Vec3.h"
#include "openmm/OpenMMException.h"
#include "openmm/MPIDForce.h"
#include "fftpack.h"
#define private protected
#include "MPIDReferenceForce.h"
#undef private

using namespace OpenMM;
using std::vector;

namespace {

double nowSeconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename F> void parallelRanges(int threads, const vector<unsigned>& bounds, F body) {
    if (threads <= 1) { body(0, bounds[0], bounds[1]); return; }
    vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back([&, t] { body(t, bounds[t], bounds[t+1]); });
    for (auto& th : pool) th.join();
}

class CellListPmeForce : public MPIDReferencePmeForce {
public:
    int threads = 1;
    // candidate pairs: for row i the partners j > i, ascending, whose minimum-image distance (the reference's own
    // formula, :2671-2676) is within the cutoff enlarged by 1e-6 relative -- the pair functions apply the exact test
    vector<size_t> rowStart;
    vector<unsigned> partner;
    vector<unsigned> bounds;        // row ranges of the threads, balanced by candidate count
    double secondsBuild = 0, secondsFixed = 0, secondsInduced = 0, secondsElectrostatics = 0;
    long long inducedFieldCalls = 0;
    vector<CellListPmeForce*> helpers;      // per-thread owners of a private _fixedMultipoleField

    ~CellListPmeForce() { for (auto* h : helpers) delete h; }

    void periodicDelta(double* d) const {   // MPIDReferencePmeForce::getPeriodicDelta (:2671-2676), same operation order
        double s = floor(d[2]*_recipBoxVectors[2][2] + 0.5);
        for (int k = 0; k < 3; k++) d[k] -= _periodicBoxVectors[2][k]*s;
        s = floor(d[1]*_recipBoxVectors[1][1] + 0.5);
        for (int k = 0; k < 3; k++) d[k] -= _periodicBoxVectors[1][k]*s;
        s = floor(d[0]*_recipBoxVectors[0][0] + 0.5);
        for (int k = 0; k < 3; k++) d[k] -= _periodicBoxVectors[0][k]*s;
    }

    void buildCandidates(const vector<MultipoleParticleData>& pd) {
        const double t0 = nowSeconds();
        const unsigned n = pd.size();
        const double rc = _cutoffDistance*(1.0 + 1e-6);
        const double rc2 = rc*rc;
        // fractional coordinates (reduced triclinic cell: lower-triangular box, upper... recip as the reference builds it)
        int nc[3], reach[3];
        for (int d = 0; d < 3; d++) {
            // perpendicular width of the cell along d = 1/|d-th column of recip|
            double col2 = 0;
            for (int k = 0; k < 3; k++) col2 += _recipBoxVectors[k][d]*_recipBoxVectors[k][d];
            const double width = 1.0/sqrt(col2);
            int half = (int) floor(width/(0.5*rc));
            int full = (int) floor(width/rc);
            if (half >= 5) { nc[d] = half; reach[d] = 2; }
            else if (full >= 3) { nc[d] = full; reach[d] = 1; }
            else { nc[d] = 1; reach[d] = 0; }
        }
        const int ncell = nc[0]*nc[1]*nc[2];
        vector<int> cellOf(n);
        vector<unsigned> cellCount(ncell + 1, 0);
        for (unsigned i = 0; i < n; i++) {
            const Vec3& p = pd[i].position;
            double f[3] = {p[0]*_recipBoxVectors[0][0] + p[1]*_recipBoxVectors[1][0] + p[2]*_recipBoxVectors[2][0],
                           p[1]*_recipBoxVectors[1][1] + p[2]*_recipBoxVectors[2][1],
                           p[2]*_recipBoxVectors[2][2]};
            int c[3];
            for (int d = 0; d < 3; d++) {
                double w = f[d] - floor(f[d]);
                c[d] = std::min(nc[d] - 1, std::max(0, (int) (w*nc[d])));
            }
            cellOf[i] = (c[0]*nc[1] + c[1])*nc[2] + c[2];
            cellCount[cellOf[i] + 1]++;
        }
        for (int c = 0; c < ncell; c++) cellCount[c+1] += cellCount[c];
        vector<unsigned> cellAtoms(n), fill(cellCount.begin(), cellCount.end() - 1);
        for (unsigned i = 0; i < n; i++) cellAtoms[fill[cellOf[i]]++] = i;      // ascending atom index inside a cell
        // neighbour cells of every cell (deduplicated when the cell grid is small)
        auto neighbourCells = [&](int cell, vector<int>& out) {
            out.clear();
            int cz = cell % nc[2], cy = (cell/nc[2]) % nc[1], cx = cell/(nc[1]*nc[2]);
            for (int dx = -reach[0]; dx <= reach[0]; dx++)
                for (int dy = -reach[1]; dy <= reach[1]; dy++)
                    for (int dz = -reach[2]; dz <= reach[2]; dz++) {
                        int x = ((cx + dx) % nc[0] + nc[0]) % nc[0], y = ((cy + dy) % nc[1] + nc[1]) % nc[1], z = ((cz + dz) % nc[2] + nc[2]) % nc[2];
                        out.push_back((x*nc[1] + y)*nc[2] + z);
                    }
            std::sort(out.begin(), out.end());
            out.erase(std::unique(out.begin(), out.end()), out.end());
        };
        // two passes (count, fill) over row ranges in parallel
        const int T = std::max(1, threads);
        vector<unsigned> even(T + 1);
        for (int t = 0; t <= T; t++) even[t] = (unsigned) ((unsigned long long) n*t/T);
        vector<vector<unsigned> > rowsOf(T);
        vector<vector<unsigned> > countOf(T);
        parallelRanges(T, even, [&](int t, unsigned b, unsigned e) {
            vector<int> cells;
            vector<unsigned> found;
            vector<unsigned>& out = rowsOf[t];
            vector<unsigned>& cnt = countOf[t];
            cnt.assign(e - b, 0);
            int lastCell = -1;
            for (unsigned i = b; i < e; i++) {
                if (cellOf[i] != lastCell) { neighbourCells(cellOf[i], cells); lastCell = cellOf[i]; }
                found.clear();
                const Vec3& pi = pd[i].position;
                for (int c : cells)
                    for (unsigned q = cellCount[c]; q < cellCount[c+1]; q++) {
                        unsigned j = cellAtoms[q];
                        if (j <= i) continue;
                        double d[3] = {pd[j].position[0] - pi[0], pd[j].position[1] - pi[1], pd[j].position[2] - pi[2]};
                        periodicDelta(d);
                        if (d[0]*d[0] + d[1]*d[1] + d[2]*d[2] <= rc2) found.push_back(j);
                    }
                std::sort(found.begin(), found.end());
                cnt[i - b] = found.size();
                out.insert(out.end(), found.begin(), found.end());
            }
        });
        rowStart.assign(n + 1, 0);
        for (int t = 0; t < T; t++)
            for (unsigned i = even[t]; i < even[t+1]; i++) rowStart[i+1] = rowStart[i] + countOf[t][i - even[t]];
        partner.resize(rowStart[n]);
        for (int t = 0; t < T; t++) std::copy(rowsOf[t].begin(), rowsOf[t].end(), partner.begin() + rowStart[even[t]]);
        // thread row ranges balanced by candidate count
        bounds.assign(T + 1, n);
        bounds[0] = 0;
        unsigned row = 0;
        for (int t = 1; t < T; t++) {
            const size_t want = rowStart[n]*t/T;
            while (row < n && rowStart[row] < want) row++;
            bounds[t] = row;
        }
        secondsBuild += nowSeconds() - t0;
    }

    // ---- fixed field: MPIDReferencePmeForce::calculateFixedMultipoleField (:2922-2949) with the pair loop of
    // MPIDReferenceForce::calculateFixedMultipoleField (:910-933) walking the candidate list
    void calculateFixedMultipoleField(const vector<MultipoleParticleData>& particleData) override {
        buildCandidates(particleData);
        const double t0 = nowSeconds();
        resizePmeArrays();
        computeMPIDBsplines(particleData);
        initializePmeGrid();
        spreadFixedMultipolesOntoGrid(particleData);
        fftpack_exec_3d(_fftplan, FFTPACK_FORWARD, _pmeGrid, _pmeGrid);
        performMPIDReciprocalConvolution();
        fftpack_exec_3d(_fftplan, FFTPACK_BACKWARD, _pmeGrid, _pmeGrid);
        computeFixedPotentialFromGrid();
        recordFixedMultipoleField();
        const double selfTerm = (4.0/3.0)*(_alphaEwald*_alphaEwald*_alphaEwald)/SQRT_PI;
        for (unsigned j = 0; j < _numParticles; j++) _fixedMultipoleField[j] += particleData[j].dipole*selfTerm;
        const int T = std::max(1, threads);
        if (T == 1) {
            fixedRows(this, particleData, 0, _numParticles);
        } else {
            while ((int) helpers.size() < T) helpers.push_back(new CellListPmeForce());
            parallelRanges(T, bounds, [&](int t, unsigned b, unsigned e) {
                CellListPmeForce* h = helpers[t];
                h->_alphaEwald = _alphaEwald; h->_cutoffDistance = _cutoffDistance; h->_cutoffDistanceSquared = _cutoffDistanceSquared;
                for (int k = 0; k < 3; k++) { h->_periodicBoxVectors[k] = _periodicBoxVectors[k]; h->_recipBoxVectors[k] = _recipBoxVectors[k]; }
                h->setDefaultTholeWidth(getDefaultTholeWidth());
                h->_numParticles = _numParticles;
                h->_fixedMultipoleField.assign(_numParticles, Vec3(0, 0, 0));
                fixedRows(h, particleData, b, e);
            });
            for (int t = 0; t < T; t++)
                for (unsigned j = 0; j < _numParticles; j++) _fixedMultipoleField[j] += helpers[t]->_fixedMultipoleField[j];
        }
        secondsFixed += nowSeconds() - t0;
    }
    void fixedRows(CellListPmeForce* target, const vector<MultipoleParticleData>& pd, unsigned b, unsigned e) {
        for (unsigned ii = b; ii < e; ii++)
            for (size_t q = rowStart[ii]; q < rowStart[ii+1]; q++) {
                const unsigned jj = partner[q];
                double dScale = 1.0, pScale = 1.0;
                if (jj <= _maxScaleIndex[ii]) getDScaleAndPScale(ii, jj, dScale, pScale);
                target->MPIDReferencePmeForce::calculateFixedMultipoleFieldPairIxn(pd[ii], pd[jj], dScale, pScale);
            }
    }

    // ---- induced field: MPIDReferencePmeForce::calculateInducedDipoleFields (:4073-4141)
    void calculateInducedDipoleFields(const vector<MultipoleParticleData>& particleData,
                                      vector<UpdateInducedDipoleFieldStruct>& fields) override {
        const double t0 = nowSeconds();
        inducedFieldCalls++;
        const Vec3 zero(0.0, 0.0, 0.0);
        for (auto& f : fields) std::fill(f.inducedDipoleField.begin(), f.inducedDipoleField.end(), zero);
        const int T = std::max(1, threads);
        if (T == 1) {
            inducedRows(particleData, fields, 0, particleData.size());
        } else {
            // private copies of the accumulators (field and, for the extrapolated algorithm, the field gradient)
            vector<vector<UpdateInducedDipoleFieldStruct> > mine(T, fields);
            for (int t = 0; t < T; t++)
                for (auto& f : mine[t])
                    for (auto& g : f.inducedDipoleFieldGradient) std::fill(g.begin(), g.end(), 0.0);
            parallelRanges(T, bounds, [&](int t, unsigned b, unsigned e) { inducedRows(particleData, mine[t], b, e); });
            for (int t = 0; t < T; t++)
                for (size_t k = 0; k < fields.size(); k++) {
                    for (size_t j = 0; j < fields[k].inducedDipoleField.size(); j++) fields[k].inducedDipoleField[j] += mine[t][k].inducedDipoleField[j];
                    for (size_t j = 0; j < fields[k].inducedDipoleFieldGradient.size(); j++)
                        for (size_t c = 0; c < fields[k].inducedDipoleFieldGradient[j].size(); c++)
                            fields[k].inducedDipoleFieldGradient[j][c] += mine[t][k].inducedDipoleFieldGradient[j][c];
                }
        }
        calculateReciprocalSpaceInducedDipoleField(fields);
        if (getPolarizationType() == MPIDReferenceForce::Extrapolated) {
            // reciprocal-space field gradient, fractional -> Cartesian, subtracted from the real-space one (:4094-4129)
            double f2c[3][3];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) f2c[i][j] = _pmeGridDimensions[j]*_recipBoxVectors[i][j];
            static const int comp[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {0, 2}, {1, 2}};      // xx yy zz xy xz yz
            static const int phiIndex[3][3] = {{4, 7, 8}, {7, 5, 9}, {8, 9, 6}};
            for (unsigned i = 0; i < _numParticles; i++) {
                double acc[6] = {0, 0, 0, 0, 0, 0};
                for (int k = 0; k < 3; k++)
                    for (int l = 0; l < 3; l++) {
                        const double m = _phidp[35*i + phiIndex[k][l]];
                        for (int c = 0; c < 6; c++) acc[c] += f2c[comp[c][0]][k]*m*f2c[comp[c][1]][l];
                    }
                for (int c = 0; c < 6; c++) fields[0].inducedDipoleFieldGradient[i][c] -= acc[c];
            }
        }
        const double selfTerm = (4.0/3.0)*(_alphaEwald*_alphaEwald*_alphaEwald)/SQRT_PI;
        for (auto& f : fields) {
            vector<Vec3>& mu = *f.inducedDipoles;
            for (unsigned j = 0; j < particleData.size(); j++) f.inducedDipoleField[j] += mu[j]*selfTerm;
        }
        secondsInduced += nowSeconds() - t0;
    }
    void inducedRows(const vector<MultipoleParticleData>& pd, vector<UpdateInducedDipoleFieldStruct>& fields, unsigned b, unsigned e) {
        for (unsigned ii = b; ii < e; ii++)
            for (size_t q = rowStart[ii]; q < rowStart[ii+1]; q++)
                calculateDirectInducedDipolePairIxns(pd[ii], pd[partner[q]], fields);
    }

    // ---- energy, forces, torques: MPIDReferencePmeForce::calculateElectrostatic (:4922-4990)
    double calculateElectrostatic(const vector<MultipoleParticleData>& particleData, vector<Vec3>& torques, vector<Vec3>& forces) override {
        const double t0 = nowSeconds();
        double energy = 0.0;
        const int T = std::max(1, threads);
        if (T == 1) {
            energy += electrostaticRows(particleData, forces, torques, 0, particleData.size());
        } else {
            vector<vector<Vec3> > f(T, vector<Vec3>(forces.size(), Vec3(0, 0, 0))), tq(T, vector<Vec3>(torques.size(), Vec3(0, 0, 0)));
            vector<double> en(T, 0.0);
            parallelRanges(T, bounds, [&](int t, unsigned b, unsigned e) { en[t] = electrostaticRows(particleData, f[t], tq[t], b, e); });
            for (int t = 0; t < T; t++) {
                energy += en[t];
                for (size_t j = 0; j < forces.size(); j++) forces[j] += f[t][j];
                for (size_t j = 0; j < torques.size(); j++) torques[j] += tq[t][j];
            }
        }
        calculatePmeSelfTorque(particleData, torques);
        energy += computeReciprocalSpaceInducedDipoleForceAndEnergy(getPolarizationType(), particleData, forces, torques);
        energy += computeReciprocalSpaceFixedMultipoleForceAndEnergy(particleData, forces, torques);
        energy += calculatePmeSelfEnergy(particleData);
        if (getPolarizationType() == MPIDReferenceForce::Extrapolated) {
            // dipole response force / torque of the extrapolated algorithm (:4956-4984)
            const double prefac = _electric/_dielectric;
            static const int row[3][3] = {{0, 3, 4}, {3, 1, 5}, {4, 5, 2}};       // xx yy zz xy xz yz storage
            for (unsigned i = 0; i < _numParticles; i++)
                for (int l = 0; l < _maxPTOrder - 1; ++l)
                    for (int m = 0; m < _maxPTOrder - 1 - l; ++m) {
                        const double p = _extPartCoefficients[l+m+1];
                        if (std::fabs(p) < 1e-6) continue;
                        const Vec3& mu = _ptDipoleD[l][i];
                        const double* g = &_ptDipoleFieldGradientD[m][6*i];
                        const double* ef = &_ptDipoleFieldD[m][3*i];
                        for (int a = 0; a < 3; a++) forces[i][a] += p*prefac*(mu[0]*g[row[a][0]] + mu[1]*g[row[a][1]] + mu[2]*g[row[a][2]]);
                        if (particleData[i].isAnisotropic) {
                            torques[i][0] += p*prefac*(mu[1]*ef[2] - mu[2]*ef[1]);
                            torques[i][1] += p*prefac*(mu[2]*ef[0] - mu[0]*ef[2]);
                            torques[i][2] += p*prefac*(mu[0]*ef[1] - mu[1]*ef[0]);
                        }
                    }
        }
        secondsElectrostatics += nowSeconds() - t0;
        return energy;
    }
    double electrostaticRows(const vector<MultipoleParticleData>& pd, vector<Vec3>& forces, vector<Vec3>& torques, unsigned b, unsigned e) const {
        double energy = 0.0;
        vector<double> scale(LAST_SCALE_TYPE_INDEX, 1.0);
        for (unsigned ii = b; ii < e; ii++)
            for (size_t q = rowStart[ii]; q < rowStart[ii+1]; q++) {
                const unsigned jj = partner[q];
                const bool scaled = jj <= _maxScaleIndex[ii];
                if (scaled) getMultipoleScaleFactors(ii, jj, scale);
                energy += calculatePmeDirectElectrostaticPairIxn(pd[ii], pd[jj], scale, forces, torques);
                if (scaled) std::fill(scale.begin(), scale.end(), 1.0);
            }
        return energy;
    }
    const vector<Vec3>& inducedDipoles() const { return _inducedDipole; }
};

// Flat parameters exactly as ReferenceCalcMPIDForceKernel::initialize keeps them (MPIDReferenceKernels.cpp:84-177)
struct CellHandle {
    int n = 0;
    vector<double> charges, dipoles, quadrupoles, octopoles, tholes, dampingFactors;
    vector<vector<double> > polarity;
    vector<int> axisTypes, atomZ, atomX, atomY;
    vector<vector<vector<int> > > covalent;
    int polarization = 0, maxIter = 60;
    double cutoff = 1, alpha = 0, defaultThole = 5, scale14 = 1, epsilon = 1e-5;
    vector<int> grid;
    vector<double> coefs;
    Vec3 box[3];
    vector<double> lastInduced;
    double seconds[5] = {0, 0, 0, 0, 0};
    long long stats[4] = {0, 0, 0, 0};
};
std::string cellLastError;

}  // namespace

extern "C" {

const char* mpidcell_last_error() { return cellLastError.c_str(); }

int mpidcell_create(int n,
                    const double* charges, const double* dipoles, const double* quadrupoles, const double* octopoles,
                    const int* axisTypes, const int* atomZ, const int* atomX, const int* atomY,
                    const double* tholes, const double* alphas,
                    const int* cov_offsets, const int* cov_indices,
                    int polarization, double cutoff, double ewaldAlpha, int nx, int ny, int nz,
                    double defaultThole, double scale14, int maxIter, double epsilon,
                    int ncoef, const double* coefs, const double* box9, void** out) {
    try {
        if (ewaldAlpha == 0.0 || nx == 0) throw OpenMMException("mpidcell_create: explicit PME parameters are required");
        CellHandle* h = new CellHandle();
        h->n = n;
        h->charges.assign(charges, charges + n); h->dipoles.assign(dipoles, dipoles + 3*(size_t) n);
        h->quadrupoles.assign(quadrupoles, quadrupoles + 6*(size_t) n); h->octopoles.assign(octopoles, octopoles + 10*(size_t) n);
        h->tholes.assign(tholes, tholes + n);
        h->axisTypes.assign(axisTypes, axisTypes + n); h->atomZ.assign(atomZ, atomZ + n); h->atomX.assign(atomX, atomX + n); h->atomY.assign(atomY, atomY + n);
        h->dampingFactors.resize(n); h->polarity.resize(n); h->covalent.resize(n);
        for (int i = 0; i < n; i++) {
            h->polarity[i].assign(alphas + 3*(size_t) i, alphas + 3*(size_t) i + 3);
            h->dampingFactors[i] = pow((alphas[3*(size_t) i] + alphas[3*(size_t) i+1] + alphas[3*(size_t) i+2])/3.0, 1.0/6.0);     // :123
            h->covalent[i].resize(8);
            for (int t = 0; t < 8; t++) {
                const int b = cov_offsets[(size_t) t*(n+1) + i], e = cov_offsets[(size_t) t*(n+1) + i + 1];
                h->covalent[i][t].assign(cov_indices + b, cov_indices + e);
            }
        }
        h->polarization = polarization; h->cutoff = cutoff; h->alpha = ewaldAlpha;
        h->grid = {nx, ny, nz};
        h->defaultThole = defaultThole; h->scale14 = scale14; h->maxIter = maxIter; h->epsilon = epsilon;
        h->coefs.assign(coefs, coefs + ncoef);
        for (int k = 0; k < 3; k++) h->box[k] = Vec3(box9[3*k], box9[3*k+1], box9[3*k+2]);
        *out = h;
        return 0;
    } catch (const std::exception& e) { cellLastError = e.what(); return 1; }
}

int mpidcell_set_box(void* handle, const double* box9) {
    CellHandle* h = static_cast<CellHandle*>(handle);
    for (int k = 0; k < 3; k++) h->box[k] = Vec3(box9[3*k], box9[3*k+1], box9[3*k+2]);
    return 0;
}

// One evaluation, set up as ReferenceCalcMPIDForceKernel::setupMPIDReferenceForce + execute do
// (MPIDReferenceKernels.cpp:179-239): a fresh force object per call, forces accumulated into `forces`.
int mpidcell_execute(void* handle, const double* pos, int threads, double* energy, double* forces) {
    CellHandle* h = static_cast<CellHandle*>(handle);
    try {
        const double t0 = nowSeconds();
        CellListPmeForce ref;
        ref.threads = std::max(1, threads);
        ref.setAlphaEwald(h->alpha);
        ref.setCutoffDistance(h->cutoff);
        ref.setPmeGridDimensions(h->grid);
        const double minAllowed = 1.999999*h->cutoff;
        if (h->box[0][0] < minAllowed || h->box[1][1] < minAllowed || h->box[2][2] < minAllowed)
            throw OpenMMException("The periodic box size has decreased to less than twice the nonbonded cutoff.");
        ref.setPeriodicBoxSize(h->box);
        ref.setDefaultTholeWidth(h->defaultThole);
        if (h->polarization == MPIDForce::Mutual) {
            ref.setPolarizationType(MPIDReferenceForce::Mutual);
            ref.setMutualInducedDipoleTargetEpsilon(h->epsilon);
            ref.setMaximumMutualInducedDipoleIterations(h->maxIter);
        } else if (h->polarization == MPIDForce::Direct) {
            ref.setPolarizationType(MPIDReferenceForce::Direct);
        } else {
            ref.setPolarizationType(MPIDReferenceForce::Extrapolated);
            ref.setExtrapolationCoefficients(h->coefs);
        }
        ref.set14ScaleFactor(h->scale14);
        vector<Vec3> p(h->n), f(h->n, Vec3(0, 0, 0));
        for (int i = 0; i < h->n; i++) p[i] = Vec3(pos[3*(size_t) i], pos[3*(size_t) i+1], pos[3*(size_t) i+2]);
        const double e = ref.calculateForceAndEnergy(p, h->charges, h->dipoles, h->quadrupoles, h->octopoles, h->tholes, h->dampingFactors,
                                                     h->polarity, h->axisTypes, h->atomZ, h->atomX, h->atomY, h->covalent, f);
        if (energy) *energy = e;
        if (forces) for (int i = 0; i < h->n; i++) for (int k = 0; k < 3; k++) forces[3*(size_t) i+k] += f[i][k];
        h->lastInduced.resize(3*(size_t) h->n);
        for (int i = 0; i < h->n; i++) for (int k = 0; k < 3; k++) h->lastInduced[3*(size_t) i+k] = ref.inducedDipoles()[i][k];
        h->seconds[0] = ref.secondsBuild; h->seconds[1] = ref.secondsFixed; h->seconds[2] = ref.secondsInduced;
        h->seconds[3] = ref.secondsElectrostatics; h->seconds[4] = nowSeconds() - t0;
        h->stats[0] = (long long) ref.partner.size(); h->stats[1] = ref.inducedFieldCalls;
        h->stats[2] = ref.getMutualInducedDipoleIterations(); h->stats[3] = ref.threads;
        return 0;
    } catch (const std::exception& e) { cellLastError = e.what(); return 1; }
}

// induced dipoles of the last mpidcell_execute
int mpidcell_get_induced(void* handle, double* out) {
    CellHandle* h = static_cast<CellHandle*>(handle);
    if (h->lastInduced.empty()) { cellLastError = "mpidcell_get_induced: no evaluation yet"; return 1; }
    memcpy(out, h->lastInduced.data(), h->lastInduced.size()*sizeof(double));
    return 0;
}

// seconds: candidate list, fixed field, induced fields (all passes), electrostatics stage, whole call;
// stats: candidate pairs, induced-field evaluations, DIIS iterations, threads
void mpidcell_get_profile(void* handle, double* seconds5, long long* stats4) {
    CellHandle* h = static_cast<CellHandle*>(handle);
    for (int k = 0; k < 5; k++) seconds5[k] = h->seconds[k];
    for (int k = 0; k < 4; k++) stats4[k] = h->stats[k];
}

void mpidcell_destroy(void* handle) { delete static_cast<CellHandle*>(handle); }

}  // extern "C"
