// TEST INFRASTRUCTURE ONLY.
//
// Host-side driver for the per-atom / per-pair arithmetic in mpidopenmmplugin_b200/csrc/mpid_math.h
// (the very functions the CUDA kernels inline).  It strings them together in plain O(N^2) loops so
// that the arithmetic can be checked against the oracle in this GPU-less container before the same
// code runs on a B200.  It is never linked into the product library and is not a fallback path.
#include "../../mpidopenmmplugin_b200/csrc/mpid_math.h"
#include "../../compat/fftpack.h"
#include <cstdio>
#include <cstring>
#include <vector>

using namespace mpid;

namespace {

struct Sys {
    int n;
    std::vector<double> pos;
    std::vector<LabAtom> lab;
    std::vector<double> damp, thole;
    std::vector<int> axis, az, ax, ay;
    std::vector<std::vector<int> > special;   // per atom: (partner, class) pairs, class 1 = excluded, 2 = 1-4
    int method, polarization;
    double cutoff, alpha, defaultThole, scale14, eps;
    int grid[3], maxIter;
    std::vector<double> coefs;
    Box box;
    PmeGeom geom;
};

template <typename T> struct Emul {
    const Sys& S;
    int n;
    std::vector<T> cart, pk;                 // 20/atom, 16/atom
    std::vector<double> mu;                  // induced dipoles (3/atom), kept in double like the engine
    std::vector<double> efix;                // alpha.E_fixed
    std::vector<T> phi, phidp;               // 35/atom
    std::vector<double> moduli[3];
    Emul(const Sys& s) : S(s), n(s.n) {}

    int pairClass(int i, int j) const {
        for (size_t k = 0; k + 1 < S.special[i].size(); k += 2)
            if (S.special[i][k] == j) return S.special[i][k+1];
        return 0;
    }
    double classScale(int c) const { return c == 0 ? 1.0 : (c == 1 ? 0.0 : S.scale14); }

    bool delta(int i, int j, double& dx, double& dy, double& dz, double& r2) const {
        dx = S.pos[3*j] - S.pos[3*i]; dy = S.pos[3*j+1] - S.pos[3*i+1]; dz = S.pos[3*j+2] - S.pos[3*i+2];
        if (S.method == PME) periodicDelta(S.box, dx, dy, dz);
        r2 = dist2Exact(dx, dy, dz);
        return S.method != PME || !(r2 > S.cutoff*S.cutoff);
    }

    void setupAtoms() {
        cart.resize(20*n); pk.resize(16*n);
        for (int i = 0; i < n; i++) {
            const LabAtom& a = S.lab[i];
            T* c = &cart[20*i];
            c[0] = T(a.charge);
            for (int k = 0; k < 3; k++) c[1+k] = T(a.dip[k]);
            for (int k = 0; k < 6; k++) c[4+k] = T(a.quad[k]);
            for (int k = 0; k < 10; k++) c[10+k] = T(a.oct[k]);
            double p[16];
            packPairMoments(a, p);
            for (int k = 0; k < 16; k++) pk[16*i+k] = T(p[k]);
        }
    }

    // ---- PME plumbing -------------------------------------------------------------------------------
    void initModuli() {
        // same construction as initializeBSplineModuli (:2720-2810)
        T th[6][5];
        bsplineWeights<T>(T(0), th);
        for (int d = 0; d < 3; d++) {
            int size = S.grid[d];
            std::vector<double> bs(size + 8, 0.0);
            for (int i = 0; i < 6; i++) bs[i+1] = th[i][0];    // bsarray[i+2] = array[i], j-1 shift folded
            moduli[d].assign(size, 0.0);
            for (int i = 0; i < size; i++) {
                double s1 = 0, s2 = 0;
                for (int j = 0; j < size; j++) {
                    double arg = 2.0*MPID_PI*i*j/size;
                    double b = (j >= 1 && j <= 6) ? bs[j] : 0.0;
                    s1 += b*cos(arg); s2 += b*sin(arg);
                }
                moduli[d][i] = s1*s1 + s2*s2;
            }
            double eps = 1e-7;
            if (moduli[d][0] < eps) moduli[d][0] = 0.5*moduli[d][1];
            for (int i = 1; i < size-1; i++)
                if (moduli[d][i] < eps) moduli[d][i] = 0.5*(moduli[d][i-1] + moduli[d][i+1]);
            if (moduli[d][size-1] < eps) moduli[d][size-1] = 0.5*moduli[d][size-2];
            for (int i = 1; i <= size; i++) {
                int k = i - 1;
                if (i > size/2) k -= size;
                double zeta = 1.0;
                if (k != 0) {
                    double s1 = 1, s2 = 1, f = MPID_PI*k/size;
                    for (int j = 1; j <= 50; j++) { double a = f/(f + MPID_PI*j); s1 += pow(a, 6); s2 += pow(a, 12); }
                    for (int j = 1; j <= 50; j++) { double a = f/(f - MPID_PI*j); s1 += pow(a, 6); s2 += pow(a, 12); }
                    zeta = s2/s1;
                }
                moduli[d][i-1] *= zeta*zeta;
            }
        }
    }

    void reciprocal(std::vector<double>& grid) {
        int nx = S.grid[0], ny = S.grid[1], nz = S.grid[2];
        std::vector<t_complex> g((size_t) nx*ny*nz);
        for (size_t i = 0; i < g.size(); i++) g[i] = t_complex(grid[i], 0);
        fftpack_t plan; fftpack_init_3d(&plan, nx, ny, nz);
        fftpack_exec_3d(plan, FFTPACK_FORWARD, g.data(), g.data());
        double expFactor = MPID_PI*MPID_PI/(S.alpha*S.alpha);
        double scaleFactor = 1.0/(MPID_PI*S.box.a[0]*S.box.b[1]*S.box.c[2]);
        for (int kx = 0; kx < nx; kx++) for (int ky = 0; ky < ny; ky++) for (int kz = 0; kz < nz; kz++) {
            size_t idx = ((size_t) kx*ny + ky)*nz + kz;
            if (kx == 0 && ky == 0 && kz == 0) { g[idx] = t_complex(0, 0); continue; }
            int mx = kx < (nx+1)/2 ? kx : kx - nx, my = ky < (ny+1)/2 ? ky : ky - ny, mz = kz < (nz+1)/2 ? kz : kz - nz;
            double hx = mx*S.box.ra[0], hy = mx*S.box.rb[0] + my*S.box.rb[1], hz = mx*S.box.rc[0] + my*S.box.rc[1] + mz*S.box.rc[2];
            double m2 = hx*hx + hy*hy + hz*hz;
            double e = scaleFactor*exp(-expFactor*m2)/(m2*moduli[0][kx]*moduli[1][ky]*moduli[2][kz]);
            g[idx].re *= e; g[idx].im *= e;
        }
        fftpack_exec_3d(plan, FFTPACK_BACKWARD, g.data(), g.data());
        fftpack_destroy(plan);
        for (size_t i = 0; i < g.size(); i++) grid[i] = T(g[i].re);   // grid lives in T on the device
    }

    template <bool FIXED> void spread(const T* frac, int stride, std::vector<double>& grid) {
        int nx = S.grid[0], ny = S.grid[1], nz = S.grid[2];
        grid.assign((size_t) nx*ny*nz, 0.0);
        for (int i = 0; i < n; i++) {
            int ig[3]; double w[3];
            pmeAtomCell(S.box, S.geom, S.pos[3*i], S.pos[3*i+1], S.pos[3*i+2], ig, w);
            T tx[6][5], ty[6][5], tz[6][5];
            bsplineWeights<T>(T(w[0]), tx); bsplineWeights<T>(T(w[1]), ty); bsplineWeights<T>(T(w[2]), tz);
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) for (int c = 0; c < 6; c++) {
                int x = (ig[0]+a) % nx, y = (ig[1]+b) % ny, z = (ig[2]+c) % nz;
                T v = spreadTerm<T, FIXED>(frac + (size_t) stride*i, tx[a], ty[b], tz[c]);
                size_t idx = ((size_t) x*ny + y)*nz + z;
                grid[idx] = T(T(grid[idx]) + v);
            }
        }
    }

    void gather(const std::vector<double>& grid, std::vector<T>& out) {
        int nx = S.grid[0], ny = S.grid[1], nz = S.grid[2];
        out.assign((size_t) 35*n, T(0));
        for (int i = 0; i < n; i++) {
            int ig[3]; double w[3];
            pmeAtomCell(S.box, S.geom, S.pos[3*i], S.pos[3*i+1], S.pos[3*i+2], ig, w);
            T tx[6][5], ty[6][5], tz[6][5];
            bsplineWeights<T>(T(w[0]), tx); bsplineWeights<T>(T(w[1]), ty); bsplineWeights<T>(T(w[2]), tz);
            T* p = &out[(size_t) 35*i];
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) for (int c = 0; c < 6; c++) {
                int x = (ig[0]+a) % nx, y = (ig[1]+b) % ny, z = (ig[2]+c) % nz;
                T q = T(grid[((size_t) x*ny + y)*nz + z]);
                for (int t = 0; t <= 4; t++) for (int u = 0; t + u <= 4; u++) for (int v = 0; t + u + v <= 4; v++)
                    p[phiIndex(t, u, v)] += q*tx[a][t]*ty[b][u]*tz[c][v];
            }
        }
    }

    // field[3n] (double accumulators) += -A^T.phi(1..3)
    void addReciprocalField(const std::vector<T>& p, std::vector<double>& field) {
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++)
                field[3*i+k] -= double(p[35*i+1])*S.geom.A[0][k] + double(p[35*i+2])*S.geom.A[1][k] + double(p[35*i+3])*S.geom.A[2][k];
    }

    // ---- fields ---------------------------------------------------------------------------------------
    void fixedField(std::vector<double>& field) {
        field.assign(3*n, 0.0);
        if (S.method == PME) {
            std::vector<T> frac(20*n);
            for (int i = 0; i < n; i++) multipolesToFractional<T>(S.geom.A, &cart[20*i], &frac[20*i]);
            std::vector<double> grid;
            spread<true>(frac.data(), 20, grid);
            reciprocal(grid);
            gather(grid, phi);
            addReciprocalField(phi, field);
            double term = (4.0/3.0)*S.alpha*S.alpha*S.alpha/MPID_SQRT_PI;
            for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) field[3*i+k] += term*S.lab[i].dip[k];
        }
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
            if (i == j) continue;
            double dx, dy, dz, r2;
            if (!delta(i, j, dx, dy, dz, r2)) continue;
            int cls = pairClass(i, j);
            T scale = T(classScale(cls));
            T r = t_sqrt(T(r2));
            T e[4], c[4];
            if (cls == 0) {      // ordinary pair: the hot-kernel formulation
                T inv = (S.damp[i] != 0 && S.damp[j] != 0) ? T(1.0/S.damp[i])*T(1.0/S.damp[j]) : T(0);
                if (S.method == PME) fieldCoefficientsOrdinary<T, true, 4>(T(r2), T(S.alpha), T(S.defaultThole), inv, c);
                else fieldCoefficientsOrdinary<T, false, 4>(T(r2), T(0), T(S.defaultThole), inv, c);
            } else {
            tholeComplements<T>(T(S.damp[i]), T(S.damp[j]), T(S.thole[i] + S.thole[j]), T(S.defaultThole), scale == T(0), r, e);
            if (S.method == PME) fieldCoefficients<T, true>(r, T(S.alpha), scale, e, 4, c);
            else fieldCoefficients<T, false>(r, T(0), scale, e, 4, c);
            }
            T ex = 0, ey = 0, ez = 0;
            fixedFieldDirected<T>(&cart[20*j], T(dx), T(dy), T(dz), c, ex, ey, ez);
            field[3*i] += ex; field[3*i+1] += ey; field[3*i+2] += ez;
        }
    }

    // field (and optionally gradient) of the current mu; leaves phidp for the reciprocal part
    void inducedField(const std::vector<double>& dip, std::vector<double>& field, std::vector<double>* grad) {
        field.assign(3*n, 0.0);
        bool wantGrad = grad != 0;
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
            if (i == j) continue;
            double dx, dy, dz, r2;
            if (!delta(i, j, dx, dy, dz, r2)) continue;
            int cls = pairClass(i, j);
            T r = t_sqrt(T(r2));
            T e[4], c[4];
            if (cls == 0) {
                T inv = (S.damp[i] != 0 && S.damp[j] != 0) ? T(1.0/S.damp[i])*T(1.0/S.damp[j]) : T(0);
                if (S.method == PME) fieldCoefficientsOrdinary<T, true, 3>(T(r2), T(S.alpha), T(S.defaultThole), inv, c);
                else fieldCoefficientsOrdinary<T, false, 3>(T(r2), T(0), T(S.defaultThole), inv, c);
            } else {
            tholeComplements<T>(T(S.damp[i]), T(S.damp[j]), T(S.thole[i] + S.thole[j]), T(S.defaultThole), classScale(cls) == 0.0, r, e);
            if (S.method == PME) fieldCoefficients<T, true>(r, T(S.alpha), T(1), e, 3, c);
            else fieldCoefficients<T, false>(r, T(0), T(1), e, 3, c);
            }
            T ex = 0, ey = 0, ez = 0;
            inducedFieldDirected<T>(T(dip[3*j]), T(dip[3*j+1]), T(dip[3*j+2]), T(dx), T(dy), T(dz), c, ex, ey, ez);
            field[3*i] += ex; field[3*i+1] += ey; field[3*i+2] += ez;
            if (wantGrad) {
                T g[6] = {0, 0, 0, 0, 0, 0};
                inducedFieldGradientDirected<T>(T(dip[3*j]), T(dip[3*j+1]), T(dip[3*j+2]), T(dx), T(dy), T(dz), c, g);
                for (int k = 0; k < 6; k++) (*grad)[6*i+k] += g[k];
            }
        }
        if (S.method == PME) {
            std::vector<T> frac(4*n);
            for (int i = 0; i < n; i++) {
                frac[4*i] = 0;
                for (int k = 0; k < 3; k++)
                    frac[4*i+1+k] = T(S.geom.A[k][0]*dip[3*i] + S.geom.A[k][1]*dip[3*i+1] + S.geom.A[k][2]*dip[3*i+2]);
            }
            std::vector<double> grid;
            spread<false>(frac.data(), 4, grid);
            reciprocal(grid);
            gather(grid, phidp);
            addReciprocalField(phidp, field);
            if (wantGrad) {
                // (:4094-4129) reciprocal field gradient, fractional -> Cartesian
                for (int i = 0; i < n; i++) {
                    const T* p = &phidp[35*i];
                    double E[3][3] = {{double(p[4]), double(p[7]), double(p[8])}, {double(p[7]), double(p[5]), double(p[9])}, {double(p[8]), double(p[9]), double(p[6])}};
                    const int gi[6] = {0, 1, 2, 0, 0, 1}, gj[6] = {0, 1, 2, 1, 2, 2};
                    for (int c = 0; c < 6; c++) {
                        double s = 0;
                        for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) s += S.geom.A[k][gi[c]]*E[k][l]*S.geom.A[l][gj[c]];
                        (*grad)[6*i+c] -= s;
                    }
                }
            }
            double term = (4.0/3.0)*S.alpha*S.alpha*S.alpha/MPID_SQRT_PI;
            for (int i = 0; i < 3*n; i++) field[i] += term*dip[i];
        }
    }

    void applyAlpha(int i, const double* f, double* out) const {
        const double* a = S.lab[i].alpha;
        out[0] = a[0]*f[0] + a[1]*f[1] + a[2]*f[2];
        out[1] = a[1]*f[0] + a[3]*f[1] + a[4]*f[2];
        out[2] = a[2]*f[0] + a[4]*f[1] + a[5]*f[2];
    }

    // small dense solve for the DIIS coefficients (:1254-1291 uses an SVD; any stable solver agrees)
    static void solveDiis(int m, const std::vector<double>& Bm, std::vector<double>& coef) {
        // Same algorithm as the device solve (mpid_kernels.cuh: diisSolveScaled): the error-overlap block is scaled to a unit
        // diagonal (its entries span 20+ orders of magnitude near convergence), Gauss-Jordan with partial pivoting, and pivots
        // below 1e-12 are dropped (coefficient 0) -- the rank truncation the reference gets from its SVD (:1276-1289).
        int rank = m + 1;
        std::vector<double> a(rank*(rank+1), 0.0), d(m, 1.0);
        for (int i = 0; i < m; i++) d[i] = Bm[i*m+i] > 0 ? 1.0/sqrt(Bm[i*m+i]) : 1.0;
        for (int i = 0; i < rank; i++) for (int j = 0; j < rank; j++) {
            double v;
            if (i == 0 && j == 0) v = 0; else if (i == 0) v = -d[j-1]; else if (j == 0) v = -d[i-1]; else v = Bm[(i-1)*m + (j-1)]*d[i-1]*d[j-1];
            a[i*(rank+1)+j] = v;
        }
        a[0*(rank+1)+rank] = -1;
        std::vector<char> dropped(rank, 0);
        for (int c = 0; c < rank; c++) {
            int piv = c;
            for (int r = c+1; r < rank; r++) if (fabs(a[r*(rank+1)+c]) > fabs(a[piv*(rank+1)+c])) piv = r;
            if (piv != c) for (int k = 0; k <= rank; k++) std::swap(a[c*(rank+1)+k], a[piv*(rank+1)+k]);
            double dd = a[c*(rank+1)+c];
            if (fabs(dd) < 1e-12) { dropped[c] = 1; continue; }
            for (int r = 0; r < rank; r++) {
                if (r == c) continue;
                double f = a[r*(rank+1)+c]/dd;
                for (int k = c; k <= rank; k++) a[r*(rank+1)+k] -= f*a[c*(rank+1)+k];
            }
        }
        coef.resize(m);
        for (int i = 0; i < m; i++) coef[i] = dropped[i+1] ? 0.0 : d[i]*a[(i+1)*(rank+1)+rank]/a[(i+1)*(rank+1)+(i+1)];
    }

    // OPT storage
    std::vector<std::vector<double> > ptDip, ptField, ptGrad;
    int iterations; double finalEps;

    void solve() {
        std::vector<double> field;
        fixedField(field);
        efix.resize(3*n); mu.resize(3*n);
        for (int i = 0; i < n; i++) applyAlpha(i, &field[3*i], &efix[3*i]);
        mu = efix;
        iterations = 0; finalEps = 0;
        std::vector<double> ifield;
        if (S.polarization == Direct) {
            if (S.method == PME) inducedField(mu, ifield, 0);   // only for phidp (:4039-4044)
            return;
        }
        if (S.polarization == Extrapolated) {
            int K = (int) S.coefs.size();
            ptDip.assign(1, mu); ptField.clear(); ptGrad.clear();
            for (int order = 1; order < K; order++) {
                std::vector<double> grad(6*n, 0.0);
                inducedField(mu, ifield, &grad);
                for (int i = 0; i < n; i++) applyAlpha(i, &ifield[3*i], &mu[3*i]);
                ptDip.push_back(mu); ptField.push_back(ifield); ptGrad.push_back(grad);
            }
            std::vector<double> part(K, 0.0);
            for (int i = 0; i < K; i++) for (int j = i; j < K; j++) part[i] += S.coefs[j];
            std::fill(mu.begin(), mu.end(), 0.0);
            for (int o = 0; o < K; o++) for (int i = 0; i < 3*n; i++) mu[i] += ptDip[o][i]*part[o];
            std::vector<double> grad(6*n, 0.0);
            inducedField(mu, ifield, &grad);
            return;
        }
        std::vector<std::vector<double> > prevDip, prevErr;
        for (int it = 0; ; it++) {
            inducedField(mu, ifield, 0);
            std::vector<double> nd(3*n), err(3*n);
            double e2 = 0;
            for (int i = 0; i < n; i++) {
                double ad[3];
                applyAlpha(i, &ifield[3*i], ad);
                for (int k = 0; k < 3; k++) {
                    nd[3*i+k] = efix[3*i+k] + ad[k];
                    err[3*i+k] = nd[3*i+k] - mu[3*i+k];
                    e2 += err[3*i+k]*err[3*i+k];
                }
            }
            prevDip.push_back(nd); prevErr.push_back(err);
            double eps = MPID_DEBYE*sqrt(e2/n);
            iterations = it; finalEps = eps;
            if (eps < S.eps || it == S.maxIter) return;
            if ((int) prevErr.size() > 20) { prevErr.erase(prevErr.begin()); prevDip.erase(prevDip.begin()); }
            int m = (int) prevErr.size();
            std::vector<double> coef(m, 1.0);
            if (m > 1) {
                std::vector<double> Bm(m*m);
                for (int a = 0; a < m; a++) for (int b = 0; b < m; b++) {
                    double s = 0;
                    for (int k = 0; k < 3*n; k++) s += prevErr[a][k]*prevErr[b][k];
                    Bm[a*m+b] = s;
                }
                solveDiis(m, Bm, coef);
            }
            std::fill(mu.begin(), mu.end(), 0.0);
            for (int a = 0; a < m; a++) for (int k = 0; k < 3*n; k++) mu[k] += prevDip[a][k]*coef[a];
        }
    }

    double energyAndForces(std::vector<double>& forces) {
        std::vector<double> torque(3*n, 0.0);
        forces.assign(3*n, 0.0);
        double energy = 0;
        bool mutual = S.polarization == Mutual;
        // "simple" sites: charge only, never polarized (the specialised instantiations of the pair kernel)
        std::vector<char> simple(n, 0);
        for (int i = 0; i < n; i++) {
            bool s0 = true;
            for (int k = 1; k < 16; k++) if (pk[16*i+k] != T(0)) s0 = false;
            for (int k = 0; k < 6; k++) if (S.lab[i].alpha[k] != 0.0) s0 = false;
            simple[i] = s0;
        }
        for (int i = 0; i < n; i++) for (int j = i+1; j < n; j++) {
            double dx, dy, dz, r2;
            if (!delta(i, j, dx, dy, dz, r2)) continue;
            int cls = pairClass(i, j);
            T sc = T(classScale(cls));
            T uI[3] = {T(mu[3*i]), T(mu[3*i+1]), T(mu[3*i+2])}, uJ[3] = {T(mu[3*j]), T(mu[3*j+1]), T(mu[3*j+2])};
            T f[3], ti[3], tj[3], e;
#define CALL4(EW, MU, A, B) e = pairElectrostatics<T, EW, MU, A, B>(&pk[16*i], &pk[16*j], uI, uJ, T(S.damp[i]), T(S.damp[j]), T(S.thole[i]), T(S.thole[j]), \
                S.lab[i].aniso != 0, S.lab[j].aniso != 0, T(dx), T(dy), T(dz), T(r2), T(S.alpha), T(S.defaultThole), sc, sc, f, ti, tj)
#define CALL(EW, MU) { if (simple[i] && simple[j]) CALL4(EW, MU, true, true); else if (simple[i]) CALL4(EW, MU, true, false); \
                       else if (simple[j]) CALL4(EW, MU, false, true); else CALL4(EW, MU, false, false); }
            if (cls == 0 && (simple[i] != simple[j])) {
                // ordinary charge-site x full-site pair: the Cartesian specialisation the GPU uses (chargeSitePair)
                const int b = simple[i] ? i : j, a = simple[i] ? j : i;        // B = bare charge, A = full site
                const T sgn = simple[i] ? T(1) : T(-1);                        // d = r_A - r_B
                const T dmp = T(S.damp[a]*S.damp[b]);
                T fB[3], tA[3], mA[20];
                unpackPairMoments<T>(&pk[16*a], mA);
                const T* ua = simple[i] ? uJ : uI;
                T phi;
                if (S.method == PME) phi = chargeSitePair<T, true>(mA, ua[0], ua[1], ua[2], dmp != T(0) ? T(1)/dmp : T(0), S.lab[a].aniso != 0,
                                                                  sgn*T(dx), sgn*T(dy), sgn*T(dz), T(r2), T(S.alpha), T(S.defaultThole), fB, tA);
                else phi = chargeSitePair<T, false>(mA, ua[0], ua[1], ua[2], dmp != T(0) ? T(1)/dmp : T(0), S.lab[a].aniso != 0,
                                                    sgn*T(dx), sgn*T(dy), sgn*T(dz), T(r2), T(0), T(S.defaultThole), fB, tA);
                const T kq = T(MPID_ELECTRIC)*pk[16*b];
                e = kq*phi;
                for (int k = 0; k < 3; k++) {
                    f[k] = (b == j ? T(1) : T(-1))*kq*fB[k];      // f is the force on j
                    ti[k] = a == i ? kq*tA[k] : T(0);
                    tj[k] = a == j ? kq*tA[k] : T(0);
                }
            }
            else if (S.method == PME) { if (mutual) CALL(true, true) else CALL(true, false) }
            else { if (mutual) CALL(false, true) else CALL(false, false) }
#undef CALL
#undef CALL4
            energy += e;
            for (int k = 0; k < 3; k++) {
                forces[3*i+k] -= f[k]; forces[3*j+k] += f[k];
                torque[3*i+k] += ti[k]; torque[3*j+k] += tj[k];
            }
        }
        const double ke = MPID_ELECTRIC;
        if (S.method == PME) {
            double a = S.alpha;
            // self torque (:4320-4333)
            double term = (2.0/3.0)*ke*a*a*a/MPID_SQRT_PI;
            for (int i = 0; i < n; i++) {
                if (S.lab[i].aniso) continue;
                const double* d = S.lab[i].dip;
                double ux = 2*mu[3*i], uy = 2*mu[3*i+1], uz = 2*mu[3*i+2];
                torque[3*i]   += term*(d[1]*uz - d[2]*uy);
                torque[3*i+1] += term*(d[2]*ux - d[0]*uz);
                torque[3*i+2] += term*(d[0]*uy - d[1]*ux);
            }
            // reciprocal induced (:3871-4024) and permanent (:3739-3866) terms
            double eInd = 0, ePerm = 0;
            for (int i = 0; i < n; i++) {
                T frac[20], find[4], m[20], cp[20], tq[3];
                multipolesToFractional<T>(S.geom.A, &cart[20*i], frac);
                find[0] = 0;
                for (int k = 0; k < 3; k++) find[1+k] = T(S.geom.A[k][0]*mu[3*i] + S.geom.A[k][1]*mu[3*i+1] + S.geom.A[k][2]*mu[3*i+2]);
                const T* p = &phi[35*i]; const T* pd = &phidp[35*i];
                // induced part
                bool addU = mutual && S.lab[i].aniso;
                torqueMultipoles<T>(&cart[20*i], addU ? T(mu[3*i]) : T(0), addU ? T(mu[3*i+1]) : T(0), addU ? T(mu[3*i+2]) : T(0), m);
                potentialToCartesian<T>(S.geom.A, pd, cp);
                reciprocalTorque<T>(m, cp, tq);
                for (int k = 0; k < 3; k++) torque[3*i+k] += ke*tq[k];
                eInd += 2.0*(find[1]*p[1] + find[2]*p[2] + find[3]*p[3]);
                double f[3];
                for (int dd = 0; dd < 3; dd++) {
                    int dt = dd == 0, du = dd == 1, dv = dd == 2;
                    double s = 2.0*contractFractional<T>(find, 4, p, dt, du, dv);
                    if (mutual) s += 2.0*contractFractional<T>(find, 4, pd, dt, du, dv);
                    s += 2.0*contractFractional<T>(frac, 20, pd, dt, du, dv);
                    f[dd] = 0.5*ke*s;
                }
                for (int k = 0; k < 3; k++) forces[3*i+k] -= f[0]*S.geom.A[0][k] + f[1]*S.geom.A[1][k] + f[2]*S.geom.A[2][k];
                // permanent part
                bool addU2 = S.lab[i].aniso != 0;
                torqueMultipoles<T>(&cart[20*i], addU2 ? T(mu[3*i]) : T(0), addU2 ? T(mu[3*i+1]) : T(0), addU2 ? T(mu[3*i+2]) : T(0), m);
                potentialToCartesian<T>(S.geom.A, p, cp);
                reciprocalTorque<T>(m, cp, tq);
                for (int k = 0; k < 3; k++) torque[3*i+k] += ke*tq[k];
                ePerm += contractFractional<T>(frac, 20, p, 0, 0, 0);
                for (int dd = 0; dd < 3; dd++) f[dd] = ke*contractFractional<T>(frac, 20, p, dd == 0, dd == 1, dd == 2);
                for (int k = 0; k < 3; k++) forces[3*i+k] -= f[0]*S.geom.A[0][k] + f[1]*S.geom.A[1][k] + f[2]*S.geom.A[2][k];
            }
            energy += 0.25*ke*eInd + 0.5*ke*ePerm;
            // self energy (:4283-4318)
            double cii = 0, dii = 0, qii = 0, oii = 0;
            for (int i = 0; i < n; i++) {
                const double* s = S.lab[i].sph;
                cii += s[0]*s[0];
                dii += s[2]*(s[2] + mu[3*i]) + s[3]*(s[3] + mu[3*i+1]) + s[1]*(s[1] + mu[3*i+2]);
                for (int k = 4; k < 9; k++) qii += s[k]*s[k];
                for (int k = 9; k < 16; k++) oii += s[k]*s[k];
            }
            double a2 = a*a;
            energy += -a*ke/MPID_SQRT_PI*(cii + (2.0/3.0)*a2*dii + (4.0/15.0)*a2*a2*qii + (8.0/105.0)*a2*a2*a2*oii);
        }
        if (S.polarization == Extrapolated) {   // (:4956-4984, :2160-2188)
            int K = (int) S.coefs.size();
            std::vector<double> part(K, 0.0);
            for (int i = 0; i < K; i++) for (int j = i; j < K; j++) part[i] += S.coefs[j];
            for (int i = 0; i < n; i++)
                for (int l = 0; l < K-1; l++) for (int m = 0; m < K-1-l; m++) {
                    double p = part[l+m+1];
                    if (fabs(p) < 1e-6) continue;
                    const double* u = &ptDip[l][3*i]; const double* g = &ptGrad[m][6*i]; const double* fl = &ptField[m][3*i];
                    forces[3*i]   += p*ke*(u[0]*g[0] + u[1]*g[3] + u[2]*g[4]);
                    forces[3*i+1] += p*ke*(u[0]*g[3] + u[1]*g[1] + u[2]*g[5]);
                    forces[3*i+2] += p*ke*(u[0]*g[4] + u[1]*g[5] + u[2]*g[2]);
                    if (S.lab[i].aniso) {
                        torque[3*i]   += p*ke*(u[1]*fl[2] - u[2]*fl[1]);
                        torque[3*i+1] += p*ke*(u[2]*fl[0] - u[0]*fl[2]);
                        torque[3*i+2] += p*ke*(u[0]*fl[1] - u[1]*fl[0]);
                    }
                }
        }
        for (int i = 0; i < n; i++) {
            if (S.axis[i] == NoAxisType) continue;
            double fI[3], fZ[3], fX[3], fY[3];
            const double* pz = &S.pos[3*S.az[i]];
            const double* px = S.ax[i] >= 0 ? &S.pos[3*S.ax[i]] : pz;
            const double* py = S.ay[i] >= 0 ? &S.pos[3*S.ay[i]] : pz;
            torqueToForce(S.axis[i], &S.pos[3*i], pz, px, py, S.ax[i] >= 0, S.ay[i] >= 0, &torque[3*i], fI, fZ, fX, fY);
            for (int k = 0; k < 3; k++) {
                forces[3*i+k] += fI[k];
                forces[3*S.az[i]+k] += fZ[k];
                if (S.ax[i] >= 0) forces[3*S.ax[i]+k] += fX[k];
                if (S.ay[i] >= 0) forces[3*S.ay[i]+k] += fY[k];
            }
        }
        return energy;
    }
};

template <typename T> int run(const Sys& S, double* energy, double* forces, double* dipoles, int* iterations) {
    Emul<T> E(S);
    E.setupAtoms();
    if (S.method == PME) E.initModuli();
    E.solve();
    std::vector<double> f;
    *energy = E.energyAndForces(f);
    memcpy(forces, f.data(), sizeof(double)*3*S.n);
    memcpy(dipoles, E.mu.data(), sizeof(double)*3*S.n);
    *iterations = E.iterations;
    return 0;
}

} // namespace

// Same flat argument convention as oracle/ref_driver.cpp (mpidref_create), plus a precision switch.
extern "C" int emul_evaluate(int n, const double* pos,
                             const double* charges, const double* dipoles, const double* quadrupoles, const double* octopoles,
                             const int* axisTypes, const int* atomZ, const int* atomX, const int* atomY,
                             const double* tholes, const double* alphas,
                             const int* cov_offsets, const int* cov_indices,
                             int method, int polarization, double cutoff, double ewaldAlpha, int nx, int ny, int nz,
                             double defaultThole, double scale14, int maxIter, double epsilon,
                             int ncoef, const double* coefs, const double* box9, int useFloat,
                             double* energy, double* forces, double* inducedDipoles, int* iterations) {
    Sys S;
    S.n = n; S.pos.assign(pos, pos + 3*n);
    S.method = method; S.polarization = polarization; S.cutoff = cutoff; S.alpha = ewaldAlpha;
    S.grid[0] = nx; S.grid[1] = ny; S.grid[2] = nz;
    S.defaultThole = defaultThole; S.scale14 = scale14; S.maxIter = maxIter; S.eps = epsilon;
    S.coefs.assign(coefs, coefs + ncoef);
    makeBox(S.box, box9, box9 + 3, box9 + 6);
    makePmeGeom(S.geom, S.box, nx, ny, nz);
    S.lab.resize(n); S.damp.resize(n); S.thole.assign(tholes, tholes + n);
    S.axis.assign(axisTypes, axisTypes + n); S.az.assign(atomZ, atomZ + n); S.ax.assign(atomX, atomX + n); S.ay.assign(atomY, atomY + n);
    for (int i = 0; i < n; i++) {
        const double* pz = atomZ[i] >= 0 ? pos + 3*atomZ[i] : pos;
        const double* px = atomX[i] >= 0 ? pos + 3*atomX[i] : pos;
        const double* py = atomY[i] >= 0 ? pos + 3*atomY[i] : pos;
        labFrameAtom(pos + 3*i, pz, px, py, axisTypes[i], atomZ[i], atomX[i], atomY[i], charges[i],
                     dipoles + 3*i, quadrupoles + 6*i, octopoles + 10*i, alphas + 3*i, S.lab[i]);
        S.damp[i] = pow((alphas[3*i] + alphas[3*i+1] + alphas[3*i+2])/3.0, 1.0/6.0);
    }
    // covalent classes exactly as setupScaleMaps (:190-225): lists 0..3 = 1-2,1-3,1-4,1-5 -> 0,0,scale14,1;
    // a later list overrides an earlier one for the same partner; only partners with a higher index count.
    S.special.resize(n);
    for (int i = 0; i < n; i++) {
        for (int t = 0; t < 4; t++) {
            int b = cov_offsets[t*(n+1)+i], e = cov_offsets[t*(n+1)+i+1];
            for (int k = b; k < e; k++) {
                int j = cov_indices[k];
                if (j < i) continue;
                int cls = t < 2 ? 1 : (t == 2 ? 2 : 0);
                bool found = false;
                for (size_t q = 0; q + 1 < S.special[i].size(); q += 2)
                    if (S.special[i][q] == j) { S.special[i][q+1] = cls; found = true; }
                if (!found) { S.special[i].push_back(j); S.special[i].push_back(cls); }
            }
        }
    }
    // make the relation symmetric for the directed loops (the scale belongs to the pair (min,max))
    for (int i = 0; i < n; i++)
        for (size_t q = 0; q + 1 < S.special[i].size(); q += 2) {
            int j = S.special[i][q];
            if (j > i) {
                bool found = false;
                for (size_t p = 0; p + 1 < S.special[j].size(); p += 2)
                    if (S.special[j][p] == i) found = true;
                if (!found) { S.special[j].push_back(i); S.special[j].push_back(S.special[i][q+1]); }
            }
        }
    if (useFloat) return run<float>(S, energy, forces, inducedDipoles, iterations);
    return run<double>(S, energy, forces, inducedDipoles, iterations);
}
