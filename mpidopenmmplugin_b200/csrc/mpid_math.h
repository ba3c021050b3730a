// MPIDB200 -- per-atom and per-pair arithmetic of the MPIDForce hot path.
//
// Every function here is a pure, templated `__host__ __device__` inline so that the SAME code is
// (a) inlined into the sm_100a kernels in mpid_kernels.cu and (b) compiled by g++ into the
// host-side unit-test harness under tests/ (where it is checked against the oracle without a GPU).
// Nothing in this file is a CPU fallback for the product: the shipped library only ever calls it
// from device code.
//
// Physics restated from the reference's Reference platform (file:line cited per function, relative
// to /root/reference/platforms/reference/src/SimTKReference/MPIDReferenceForce.cpp).  Organisation,
// data layout and the way the pair interaction is evaluated are our own:
//   * QI-frame moments are obtained by contracting traceless Cartesian tensors with the pair frame
//     axes instead of building the 5x5 / 7x7 spherical rotation matrices per pair;
//   * the real-space Ewald factors are carried as B_k = mScale + bVec[k] so that ordinary pairs
//     (mScale = 1) start from erfc() and never subtract nearly equal numbers in FP32;
//   * field kernels are written "directed" (field at me due to the other atom) so they can be used
//     from a gather over a full neighbour list with no atomics.
#ifndef MPIDB200_MATH_H_
#define MPIDB200_MATH_H_

#include <cmath>

#if defined(__CUDACC__)
#define MPID_HD __host__ __device__ __forceinline__
#else
#define MPID_HD inline
#endif

namespace mpid {

// ---- enums mirrored from openmmapi/include/openmm/MPIDForce.h:58-101 ---------------------------
enum AxisType { ZThenX = 0, Bisector = 1, ZBisect = 2, ThreeFold = 3, ZOnly = 4, NoAxisType = 5 };
enum Polarization { Mutual = 0, Direct = 1, Extrapolated = 2 };
enum Method { NoCutoff = 0, PME = 1 };

// MPIDReferenceForce.cpp:39 (the Reference class' own Coulomb constant; the API uses ...8456)
#define MPID_ELECTRIC 138.935455846
// MPIDReferenceForce.cpp:2541 -- the reference truncates sqrt(pi); keep its value for parity.
#define MPID_SQRT_PI 1.77245385091
#define MPID_PI 3.14159265358979323846
#define MPID_DEBYE 48.033324

// ---- tiny vector helpers -------------------------------------------------------------------------
template <typename T> struct V3 { T x, y, z; };
template <typename T> MPID_HD V3<T> mk(T x, T y, T z) { V3<T> v; v.x = x; v.y = y; v.z = z; return v; }
template <typename T> MPID_HD T dot(const V3<T>& a, const V3<T>& b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
template <typename T> MPID_HD V3<T> cross(const V3<T>& a, const V3<T>& b) {
    return mk<T>(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x);
}
template <typename T> MPID_HD V3<T> operator+(const V3<T>& a, const V3<T>& b) { return mk<T>(a.x+b.x, a.y+b.y, a.z+b.z); }
template <typename T> MPID_HD V3<T> operator-(const V3<T>& a, const V3<T>& b) { return mk<T>(a.x-b.x, a.y-b.y, a.z-b.z); }
template <typename T> MPID_HD V3<T> operator*(const V3<T>& a, T s) { return mk<T>(a.x*s, a.y*s, a.z*s); }
template <typename T> MPID_HD T comp(const V3<T>& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// normalise in place, return the old norm (MPIDReferenceForce.cpp:251-258: zero vectors stay zero)
template <typename T> MPID_HD T normalize(V3<T>& a) {
    T n = sqrt(dot(a, a));
    if (n > T(0)) { T inv = T(1)/n; a.x *= inv; a.y *= inv; a.z *= inv; }
    return n;
}

MPID_HD float  t_exp(float x)  { return expf(x); }
MPID_HD double t_exp(double x) { return exp(x); }
MPID_HD float  t_erfc(float x)  { return erfcf(x); }
MPID_HD double t_erfc(double x) { return erfc(x); }
MPID_HD float  t_sqrt(float x)  { return sqrtf(x); }
MPID_HD double t_sqrt(double x) { return sqrt(x); }
MPID_HD float  t_abs(float x)  { return fabsf(x); }
MPID_HD double t_abs(double x) { return fabs(x); }

// ---- fast single-precision primitives for the field kernels (the double overloads stay exact) ---------
// On the device these map to MUFU.RSQ / MUFU.EX2 / MUFU.RCP; the host build uses libm so the unit tests
// exercise the same formulas.
MPID_HD float t_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f/sqrtf(x);
#endif
}
MPID_HD double t_rsqrt(double x) { return 1.0/sqrt(x); }
MPID_HD float t_expneg(float x) {     // exp(x) for x <= 0; relative error ~ |x| * 1e-7
#if defined(__CUDA_ARCH__)
    return __expf(x);
#else
    return expf(x);
#endif
}
MPID_HD double t_expneg(double x) { return exp(x); }
MPID_HD float t_rcp(float x) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f/x;
#endif
}
MPID_HD double t_rcp(double x) { return 1.0/x; }
// erfc(x) for x >= 0 given ex = exp(-x*x): erfc(x) = ex * erfcx(x), erfcx as a degree-10 polynomial in
// t = 1/(1 + 0.75 x) fitted on [0,4] (max relative error 2.7e-7 including FP32 Horner rounding).
MPID_HD float t_erfc_ex(float x, float ex) {
    if (x < 4.0f) {
        const float t = t_rcp(1.0f + 0.75f*x);
        float p = -9.0269587385e-02f;
        p = p*t + 5.4411171075e-01f;
        p = p*t - 1.3480634979e+00f;
        p = p*t + 1.6405107015e+00f;
        p = p*t - 7.4142197215e-01f;
        p = p*t - 3.4753290535e-01f;
        p = p*t + 2.5065307994e-01f;
        p = p*t + 2.3229455381e-01f;
        p = p*t + 4.3818957598e-01f;
        p = p*t + 4.2144743440e-01f;
        p = p*t + 8.0885749310e-05f;
        return p*ex;
    }
    return erfcf(x);
}
MPID_HD double t_erfc_ex(double x, double) { return erfc(x); }

// ---- periodic box ---------------------------------------------------------------------------------
// Box vectors a=(ax,0,0), b=(bx,by,0), c=(cx,cy,cz) and the reciprocal vectors of
// MPIDReferencePmeForce::setPeriodicBoxSize (:2618-2636).
struct Box {
    double a[3], b[3], c[3];
    double ra[3], rb[3], rc[3];   // _recipBoxVectors[0..2]
};
inline void makeBox(Box& B, const double* a, const double* b, const double* c) {
    for (int i = 0; i < 3; i++) { B.a[i] = a[i]; B.b[i] = b[i]; B.c[i] = c[i]; }
    double det = a[0]*b[1]*c[2];
    double s = 1.0/det;
    B.ra[0] = b[1]*c[2]*s; B.ra[1] = 0; B.ra[2] = 0;
    B.rb[0] = -b[0]*c[2]*s; B.rb[1] = a[0]*c[2]*s; B.rb[2] = 0;
    B.rc[0] = (b[0]*c[1]-b[1]*c[0])*s; B.rc[1] = -a[0]*c[1]*s; B.rc[2] = a[0]*b[1]*s;
}

// Minimum-image displacement exactly as MPIDReferencePmeForce::getPeriodicDelta (:2671-2676).
// Written with explicit non-fused multiplies/adds so that the device result is bit-identical to the
// oracle's (the pair set is decided by r2 <= rc2 on this value).
#if defined(__CUDA_ARCH__)
#define MPID_DMUL(a, b) __dmul_rn((a), (b))
#define MPID_DADD(a, b) __dadd_rn((a), (b))
#else
// host: keep products and sums as separate roundings (volatile defeats FMA contraction)
inline double mpid_dmul_(double a, double b) { volatile double r = a*b; return r; }
inline double mpid_dadd_(double a, double b) { volatile double r = a+b; return r; }
#define MPID_DMUL(a, b) mpid::mpid_dmul_((a), (b))
#define MPID_DADD(a, b) mpid::mpid_dadd_((a), (b))
#endif
MPID_HD void periodicDelta(const Box& B, double& dx, double& dy, double& dz) {
    double s = floor(MPID_DADD(MPID_DMUL(dz, B.rc[2]), 0.5));
    dx = MPID_DADD(dx, -MPID_DMUL(B.c[0], s)); dy = MPID_DADD(dy, -MPID_DMUL(B.c[1], s)); dz = MPID_DADD(dz, -MPID_DMUL(B.c[2], s));
    s = floor(MPID_DADD(MPID_DMUL(dy, B.rb[1]), 0.5));
    dx = MPID_DADD(dx, -MPID_DMUL(B.b[0], s)); dy = MPID_DADD(dy, -MPID_DMUL(B.b[1], s)); dz = MPID_DADD(dz, -MPID_DMUL(B.b[2], s));
    s = floor(MPID_DADD(MPID_DMUL(dx, B.ra[0]), 0.5));
    dx = MPID_DADD(dx, -MPID_DMUL(B.a[0], s)); dy = MPID_DADD(dy, -MPID_DMUL(B.a[1], s)); dz = MPID_DADD(dz, -MPID_DMUL(B.a[2], s));
}
// r^2 with the same association as Vec3::dot (x*x + y*y + z*z, left to right, no FMA)
MPID_HD double dist2Exact(double dx, double dy, double dz) {
    return MPID_DADD(MPID_DADD(MPID_DMUL(dx, dx), MPID_DMUL(dy, dy)), MPID_DMUL(dz, dz));
}

// =====================================================================================================
// Per-atom: molecular-frame parameters -> lab-frame moments
// =====================================================================================================
// Internal component orders (MPIDReferenceForce.h:809-810):
//   quadrupole  QXX QXY QXZ QYY QYZ QZZ ; octopole QXXX QXXY QXXZ QXYY QXYZ QXZZ QYYY QYYZ QYZZ QZZZ
// API orders (MPIDForce.h, MPIDReferenceForce.cpp:305-321):
//   quadrupole  XX XY YY XZ YZ ZZ       ; octopole XXX XXY XYY YYY XXZ XYZ YYZ XZZ YZZ ZZZ
struct LabAtom {
    double charge;
    double dip[3];       // lab Cartesian dipole (x,y,z)
    double quad[6];      // lab Cartesian quadrupole, internal order
    double oct[10];      // lab Cartesian octopole, internal order
    double sph[16];      // q, Q10 Q11c Q11s, Q20 Q21c Q21s Q22c Q22s, Q30 Q31c Q31s Q32c Q32s Q33c Q33s (lab)
    double alpha[6];     // lab polarizability tensor, internal quadrupole order (zero for frameless atoms)
    int    aniso;
};

MPID_HD int symIdx2(int i, int j) {   // internal quadrupole order
    if (i > j) { int t = i; i = j; j = t; }
    return i == 0 ? j : (i == 1 ? 2 + j : 5);
}
MPID_HD int symIdx3(int i, int j, int k) {   // internal octopole order, any permutation
    int a = i, b = j, c = k, t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
    // (a,b,c) sorted: 000 001 002 011 012 022 111 112 122 222
    if (a == 0) return b == 0 ? c : (b == 1 ? 2 + c : 5);
    if (a == 1) return b == 1 ? 5 + c : 8;
    return 9;
}

// Real-spherical l=2 and l=3 rotation matrices generated from the l=1 matrix D1 (ordering z,x,y).
// Same mathematics as buildSphericalQuadrupoleRotationMatrix / ...Octopole... (:686-766): the
// l-th representation is obtained by pushing products of D1 rows through the Cartesian->spherical
// maps below, so that rotating a traceless Cartesian tensor and converting it is the same thing as
// converting and rotating the spherical vector.  We never build them per pair; they are only used
// per atom (molecular -> lab frame), where we generate the rotated spherical moments by rotating an
// equivalent traceless Cartesian tensor (see sphToTraceless / tracelessToSph).

// spherical (5) <-> traceless Cartesian quadrupole (internal order, 6 with zz dependent)
MPID_HD void sphToTraceless2(const double* s, double* q) {
    const double c = 0.28867513459481287;  // 1/(2 sqrt 3)
    double zz = s[0]/3.0;
    double d  = s[3]*0.5773502691896258;   // (xx - yy) = Q22c / sqrt(3)
    q[5] = zz; q[2] = s[1]*c; q[4] = s[2]*c; q[1] = s[4]*c;
    q[0] = 0.5*(d - zz); q[3] = 0.5*(-d - zz);
}
template <typename T> MPID_HD void tracelessToSph2(T zz, T xz, T yz, T xxmyy, T xy, T* s) {
    s[0] = T(3)*zz;
    s[1] = T(3.4641016151377544)*xz;   // 3 * 2/sqrt(3)
    s[2] = T(3.4641016151377544)*yz;
    s[3] = T(1.7320508075688772)*xxmyy; // 3 / sqrt(3)
    s[4] = T(3.4641016151377544)*xy;
}
// spherical (7) <-> traceless Cartesian octopole (internal order, 10 with 3 dependent)
MPID_HD void sphToTraceless3(const double* s, double* o) {
    const double c1 = 1.0/(15.0*1.224744871391589);     // 1/(15 sqrt(3/2))
    const double c2 = 1.0/(15.0*0.7745966692414834);    // 1/(15 sqrt(3/5))
    const double c3 = 1.0/(15.0*0.31622776601683794);   // 1/(15 sqrt(1/10))
    double zzz = s[0]/15.0, xzz = s[1]*c1, yzz = s[2]*c1;
    double A = s[3]*c2, xyz = 0.5*s[4]*c2, Bc = s[5]*c3, Cs = s[6]*c3;
    double xxz = 0.5*(A - zzz), yyz = 0.5*(-A - zzz);
    double xyy = 0.25*(-xzz - Bc), xxx = -xzz - xyy;
    double xxy = 0.25*(Cs - yzz), yyy = -yzz - xxy;
    o[0] = xxx; o[1] = xxy; o[2] = xxz; o[3] = xyy; o[4] = xyz; o[5] = xzz; o[6] = yyy; o[7] = yyz; o[8] = yzz; o[9] = zzz;
}
template <typename T> MPID_HD void tracelessToSph3(T zzz, T xzz, T yzz, T xxz, T yyz, T xyz, T xxx, T xyy, T xxy, T yyy, T* s) {
    s[0] = T(15)*zzz;
    s[1] = T(18.371173070873837)*xzz;               // 15 sqrt(3/2)
    s[2] = T(18.371173070873837)*yzz;
    s[3] = T(11.618950038622252)*(xxz - yyz);       // 15 sqrt(3/5)
    s[4] = T(23.237900077244504)*xyz;               // 30 sqrt(3/5)
    s[5] = T(4.743416490252569)*(xxx - T(3)*xyy);   // 15 sqrt(1/10)
    s[6] = T(4.743416490252569)*(T(3)*xxy - yyy);
}

// Rotate a symmetric rank-2 tensor (internal order) : out_ij = sum_kl R[k][i] R[l][j] in_kl, where
// R rows are the frame axes expressed in the lab (R[0]=x axis, R[1]=y axis, R[2]=z axis).
MPID_HD void rotateSym2(const double R[3][3], const double* in, double* out) {
    double m[3][3] = {{in[0], in[1], in[2]}, {in[1], in[3], in[4]}, {in[2], in[4], in[5]}};
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) s += R[k][i]*R[l][j]*m[k][l];
            out[symIdx2(i, j)] = s;
        }
}
MPID_HD void rotateSym3(const double R[3][3], const double* in, double* out) {
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++)
            for (int k = j; k < 3; k++) {
                double s = 0.0;
                for (int l = 0; l < 3; l++)
                    for (int m = 0; m < 3; m++)
                        for (int n = 0; n < 3; n++) s += R[l][i]*R[m][j]*R[n][k]*in[symIdx3(l, m, n)];
                out[symIdx3(i, j, k)] = s;
            }
}

// Build the lab-frame description of one atom.
//   reference: loadParticleData (:282-350), checkChiralCenterAtParticle (:357-382),
//              applyRotationMatrixToParticle (:399-648), applyRotationMatrix (:788-800).
// pos* are the positions of the atom and of its Z/X/Y anchors (ignored when the index is < 0).
// dipole/quadrupole/octopole/polarity are in the API orders.
MPID_HD void labFrameAtom(const double* pi, const double* pz, const double* px, const double* py,
                          int axisType, int atomZ, int atomX, int atomY,
                          double charge, const double* dipole, const double* quadrupole, const double* octopole,
                          const double* polarity, LabAtom& out) {
    double d[3] = {dipole[0], dipole[1], dipole[2]};
    double q[6] = {quadrupole[0], quadrupole[1], quadrupole[3], quadrupole[2], quadrupole[4], quadrupole[5]};
    double o[10] = {octopole[0], octopole[1], octopole[4], octopole[2], octopole[5], octopole[7], octopole[3], octopole[6], octopole[8], octopole[9]};
    // spherical moments in the molecular frame (:323-341)
    double s[16];
    s[0] = charge;
    s[1] = d[2]; s[2] = d[0]; s[3] = d[1];
    tracelessToSph2<double>(q[5], q[2], q[4], q[0] - q[3], q[1], s + 4);
    tracelessToSph3<double>(o[9], o[5], o[8], o[2], o[7], o[4], o[0], o[3], o[1], o[6], s + 9);
    out.charge = charge;
    out.aniso = (polarity[0] != polarity[1] || polarity[0] != polarity[2]) ? 1 : 0;
    for (int i = 0; i < 6; i++) out.alpha[i] = 0.0;   // frameless atoms keep a zero tensor

    // chirality (:357-382): only ZThenX with a Y anchor; flips y-odd dipole/quadrupole parts, not octopoles
    if (atomY > -1 && axisType == ZThenX) {
        V3<double> ad = mk(pi[0]-py[0], pi[1]-py[1], pi[2]-py[2]);
        V3<double> bd = mk(pz[0]-py[0], pz[1]-py[1], pz[2]-py[2]);
        V3<double> cd = mk(px[0]-py[0], px[1]-py[1], px[2]-py[2]);
        if (dot(cross(bd, cd), ad) < 0.0) {
            d[1] = -d[1]; q[1] = -q[1]; q[4] = -q[4];
            s[3] = -s[3]; s[6] = -s[6]; s[8] = -s[8];
        }
    }
    if (atomZ < 0) {   // no frame: moments are taken as already being in the lab frame (:795)
        for (int i = 0; i < 3; i++) out.dip[i] = d[i];
        for (int i = 0; i < 6; i++) out.quad[i] = q[i];
        for (int i = 0; i < 10; i++) out.oct[i] = o[i];
        for (int i = 0; i < 16; i++) out.sph[i] = s[i];
        return;
    }
    // frame axes (:412-470)
    V3<double> vz = mk(pz[0]-pi[0], pz[1]-pi[1], pz[2]-pi[2]);
    normalize(vz);
    V3<double> vx, vy;
    if (axisType == ZOnly) {
        vx = (fabs(vz.x) < 0.866) ? mk(1.0, 0.0, 0.0) : mk(0.0, 1.0, 0.0);
    } else {
        vx = mk(px[0]-pi[0], px[1]-pi[1], px[2]-pi[2]);
        if (axisType == Bisector) {
            normalize(vx);
            vz = vz + vx;
            normalize(vz);
        } else if (axisType == ZBisect) {
            normalize(vx);
            vy = mk(py[0]-pi[0], py[1]-pi[1], py[2]-pi[2]);
            normalize(vy);
            vx = vx + vy;
            normalize(vx);
        } else if (axisType == ThreeFold) {
            normalize(vx);
            vy = mk(py[0]-pi[0], py[1]-pi[1], py[2]-pi[2]);
            normalize(vy);
            vz = vz + vx + vy;
            normalize(vz);
        }
    }
    double dt = dot(vz, vx);
    vx = vx - vz*dt;
    normalize(vx);
    vy = cross(vz, vx);
    double R[3][3] = {{vx.x, vx.y, vx.z}, {vy.x, vy.y, vy.z}, {vz.x, vz.y, vz.z}};
    // Cartesian moments -> lab
    for (int i = 0; i < 3; i++) out.dip[i] = d[0]*R[0][i] + d[1]*R[1][i] + d[2]*R[2][i];
    rotateSym2(R, q, out.quad);
    rotateSym3(R, o, out.oct);
    double ba[6] = {polarity[0], 0.0, 0.0, polarity[1], 0.0, polarity[2]};
    rotateSym2(R, ba, out.alpha);
    // spherical moments -> lab.  The reference multiplies by the l=1,2,3 rotation matrices; an l-vector
    // and the traceless tensor it stands for transform identically, so rotate the equivalent tensor.
    out.sph[0] = charge;
    {
        double sd[3] = {s[2], s[3], s[1]};   // x,y,z
        double ld[3];
        for (int i = 0; i < 3; i++) ld[i] = sd[0]*R[0][i] + sd[1]*R[1][i] + sd[2]*R[2][i];
        out.sph[1] = ld[2]; out.sph[2] = ld[0]; out.sph[3] = ld[1];
        double tq[6], lq[6], to[10], lo[10];
        sphToTraceless2(s + 4, tq);
        rotateSym2(R, tq, lq);
        tracelessToSph2<double>(lq[5], lq[2], lq[4], lq[0] - lq[3], lq[1], out.sph + 4);
        sphToTraceless3(s + 9, to);
        rotateSym3(R, to, lo);
        tracelessToSph3<double>(lo[9], lo[5], lo[8], lo[2], lo[7], lo[4], lo[0], lo[3], lo[1], lo[6], out.sph + 9);
    }
}

// Pack the 16 numbers the pair-energy kernel needs per atom: charge, lab dipole (x,y,z) and the
// traceless Cartesian tensors equivalent to the lab spherical quadrupole / octopole.
//   pk[0]=q  pk[1..3]=d(x,y,z)  pk[4..8]= Txx Txy Txz Tyy Tyz  pk[9..15]= Oxxx Oxxy Oxxz Oxyy Oxyz Oyyy Oyyz
// (Tzz, Oxzz, Oyzz, Ozzz follow from tracelessness.)
MPID_HD void packPairMoments(const LabAtom& a, double* pk) {
    double tq[6], to[10];
    sphToTraceless2(a.sph + 4, tq);
    sphToTraceless3(a.sph + 9, to);
    pk[0] = a.charge;
    pk[1] = a.sph[2]; pk[2] = a.sph[3]; pk[3] = a.sph[1];
    pk[4] = tq[0]; pk[5] = tq[1]; pk[6] = tq[2]; pk[7] = tq[3]; pk[8] = tq[4];
    pk[9] = to[0]; pk[10] = to[1]; pk[11] = to[2]; pk[12] = to[3]; pk[13] = to[4]; pk[14] = to[6]; pk[15] = to[7];
}

// =====================================================================================================
// Radial factors shared by the field kernels
// =====================================================================================================
// Thole polynomials of getAndScaleInverseRs (:802-835) / getDampedInverseDistances (:2678-2718).
// Returns e_k = 1 - lambda_k = exp(-au)*poly_k(au) (k = 3,5,7,9), i.e. the *complement* of the
// reference's scale factors, which is what survives without cancellation for ordinary pairs.
template <typename T> MPID_HD void tholeComplements(T dampI, T dampJ, T tholeSum, T defaultThole, bool useSum, T r, T* e) {
    e[0] = e[1] = e[2] = e[3] = T(0);
    T damp = dampI*dampJ;
    if (damp != T(0)) {
        T au = (useSum ? tholeSum : defaultThole)*(r/damp);
        if (au < T(50)) {
            T ex = t_exp(-au);
            T au2 = au*au, au3 = au2*au, au4 = au3*au, au5 = au4*au;
            T p3 = T(1) + au + T(0.5)*au2;
            T p5 = p3 + au3*T(1.0/6.0);
            e[0] = ex*p3;
            e[1] = ex*p5;
            e[2] = ex*(p5 + au4*T(1.0/30.0));
            e[3] = ex*(p5 + au4*T(4.0/105.0) + au5*T(1.0/210.0));
        }
    }
}

// Coefficients c_k of the permanent/induced field kernels,
//   PME      : c_k = bn_k - (1 - s*lambda_k) * (2k-1)!!/r^(2k+1)       (:2839-2870, :4186-4231)
//   NoCutoff : c_k = s*lambda_k * (2k-1)!!/r^(2k+1)                      (:802-835)
// for k = 1..4, with s the d/p-scale of the pair and lambda_k = 1 - e_k.
template <typename T, bool EWALD> MPID_HD void fieldCoefficients(T r, T alphaEwald, T scale, const T* e, int nk, T* c) {
    T rinv = T(1)/r, rinv2 = rinv*rinv;
    T bare = rinv;                // becomes (2k-1)!!/r^(2k+1)
    T bn = T(0), ex = T(0), a2n = T(0), alsq2 = T(0);
    if (EWALD) {
        T ra = alphaEwald*r;
        bn = t_erfc(ra)*rinv;
        ex = t_exp(-(ra*ra));
        alsq2 = T(2)*alphaEwald*alphaEwald;
        a2n = T(1)/(T(MPID_SQRT_PI)*alphaEwald);
    }
    T fac = T(1);
    for (int k = 0; k < nk; k++) {
        bare = bare*fac*rinv2;
        T oneMinus = (T(1) - scale) + scale*e[k];    // 1 - s*lambda_k
        if (EWALD) {
            a2n *= alsq2;
            bn = (fac*bn + a2n*ex)*rinv2;
            c[k] = bn - oneMinus*bare;
        } else {
            c[k] = (T(1) - oneMinus)*bare;
        }
        fac += T(2);
    }
}

// The same coefficients for an ORDINARY pair (d/p/u-scale = 1, default Thole width), from r^2, arranged
// for the hot field kernels: one rsqrt, one exp shared between erfc and the Gaussian terms, no divisions,
// and a branch-free Thole part (invDamp = 1/(damp_i damp_j), or 0 when either site is undamped).
//   c_k = bn_k - e_k (2k-1)!!/r^(2k+1)  (PME)      c_k = (1 - e_k) (2k-1)!!/r^(2k+1)  (no cutoff)
template <typename T, bool EWALD, int NK>
MPID_HD void fieldCoefficientsOrdinary(T r2, T alphaEwald, T defaultThole, T invDamp, T* c) {
    const T rinv = t_rsqrt(r2);
    const T r = r2*rinv, rinv2 = rinv*rinv;
    // Thole complements e_k = exp(-au) poly_k(au); au >= 50 (or undamped) -> 0  (:2693-2701)
    T e[4] = {T(0), T(0), T(0), T(0)};
    {
        const T au = defaultThole*r*invDamp;
        const bool damped = (invDamp != T(0)) && (au < T(50));
        const T ex = damped ? t_expneg(-au) : T(0);
        const T au2 = au*au, au3 = au2*au;
        const T p3 = T(1) + au + T(0.5)*au2;
        const T p5 = p3 + au3*T(1.0/6.0);
        e[0] = ex*p3;
        e[1] = ex*p5;
        if (NK > 2) e[2] = ex*(p5 + au2*au2*T(1.0/30.0));
        if (NK > 3) e[3] = ex*(p5 + au2*au2*(T(4.0/105.0) + au*T(1.0/210.0)));
    }
    T bare = rinv;
    T bn = T(0), ex2 = T(0), a2n = T(0), alsq2 = T(0);
    if (EWALD) {
        const T x = alphaEwald*r;
        ex2 = t_expneg(-(x*x));
        bn = t_erfc_ex(x, ex2)*rinv;
        alsq2 = T(2)*alphaEwald*alphaEwald;
        a2n = T(1.0/MPID_SQRT_PI)/alphaEwald;
    }
    T fac = T(1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < NK; k++) {
        bare = bare*fac*rinv2;
        if (EWALD) {
            a2n *= alsq2;
            bn = (fac*bn + a2n*ex2)*rinv2;
            c[k] = bn - e[k]*bare;
        } else {
            c[k] = (T(1) - e[k])*bare;
        }
        fac += T(2);
    }
}

// =====================================================================================================
// Directed field kernels: field (and field gradient) at "me" due to the moments of "other";
// d = r_other - r_me (minimum image).
// =====================================================================================================

// Permanent-multipole field.  m = {q, dx,dy,dz, Qxx,Qxy,Qxz,Qyy,Qyz,Qzz, Oxxx..Ozzz (internal order)}
// of the source atom, c[0..3] from fieldCoefficients.
//   reference: calculateFixedMultipoleFieldPairIxn, PME (:2812-2920) and no-cutoff (:837-909).
template <typename T> MPID_HD void fixedFieldDirected(const T* m, T dx, T dy, T dz, const T* c, T& ex, T& ey, T& ez) {
    T dd = m[1]*dx + m[2]*dy + m[3]*dz;
    T qx = m[4]*dx + m[5]*dy + m[6]*dz;
    T qy = m[5]*dx + m[7]*dy + m[8]*dz;
    T qz = m[6]*dx + m[8]*dy + m[9]*dz;
    T qdd = qx*dx + qy*dy + qz*dz;
    // O.d (symmetric matrix), then (O.d).d
    T oxx = m[10]*dx + m[11]*dy + m[12]*dz;
    T oxy = m[11]*dx + m[13]*dy + m[14]*dz;
    T oxz = m[12]*dx + m[14]*dy + m[15]*dz;
    T oyy = m[13]*dx + m[16]*dy + m[17]*dz;
    T oyz = m[14]*dx + m[17]*dy + m[18]*dz;
    T ozz = m[15]*dx + m[18]*dy + m[19]*dz;
    T ox = oxx*dx + oxy*dy + oxz*dz;
    T oy = oxy*dx + oyy*dy + oyz*dz;
    T oz = oxz*dx + oyz*dy + ozz*dz;
    T oddd = ox*dx + oy*dy + oz*dz;
    T radial = -(c[0]*m[0] - c[1]*dd + c[2]*qdd) + c[3]*oddd;
    ex += T(2)*c[1]*qx - c[0]*m[1] - T(3)*c[2]*ox + radial*dx;
    ey += T(2)*c[1]*qy - c[0]*m[2] - T(3)*c[2]*oy + radial*dy;
    ez += T(2)*c[1]*qz - c[0]*m[3] - T(3)*c[2]*oz + radial*dz;
}

// Field of an induced dipole mu of "other" at "me"; c from fieldCoefficients with scale = 1 (uscale).
//   reference: calculateDirectInducedDipolePairIxn (:4143-4159), calculateInducedDipolePairIxn (:948-960).
template <typename T> MPID_HD void inducedFieldDirected(T mx, T my, T mz, T dx, T dy, T dz, const T* c, T& ex, T& ey, T& ez) {
    T mud = (mx*dx + my*dy + mz*dz)*c[1];
    ex += mud*dx - c[0]*mx;
    ey += mud*dy - c[0]*my;
    ez += mud*dz - c[0]*mz;
}
// Field gradient of the same dipole at "me" (Extrapolated polarization only), order xx,yy,zz,xy,xz,yz.
//   reference: :4234-4279 (PME) / :984-1030 (no cutoff); g_me += E(mu_other, d).
template <typename T> MPID_HD void inducedFieldGradientDirected(T mx, T my, T mz, T dx, T dy, T dz, const T* c, T* g) {
    T mud = mx*dx + my*dy + mz*dz;
    T a = mud*c[2], b = c[1];
    g[0] += a*dx*dx - (T(2)*mx*dx + mud)*b;
    g[1] += a*dy*dy - (T(2)*my*dy + mud)*b;
    g[2] += a*dz*dz - (T(2)*mz*dz + mud)*b;
    g[3] += a*dx*dy - (mx*dy + my*dx)*b;
    g[4] += a*dx*dz - (mx*dz + mz*dx)*b;
    g[5] += a*dy*dz - (my*dz + mz*dy)*b;
}

// Inverse of packPairMoments: the 20 Cartesian components {c, d, Q (6), O (10)} (internal orders) of the traceless
// tensors the pair-energy code works with.  They equal the lab Cartesian moments with the trace removed -- the
// reference's energy routine sees the moments through their spherical components, i.e. without any trace the input had.
template <typename T> MPID_HD void unpackPairMoments(const T* pk, T* m) {
    m[0] = pk[0]; m[1] = pk[1]; m[2] = pk[2]; m[3] = pk[3];
    m[4] = pk[4]; m[5] = pk[5]; m[6] = pk[6]; m[7] = pk[7]; m[8] = pk[8]; m[9] = -(pk[4] + pk[7]);
    m[10] = pk[9]; m[11] = pk[10]; m[12] = pk[11]; m[13] = pk[12]; m[14] = pk[13]; m[15] = -(pk[9] + pk[12]);
    m[16] = pk[14]; m[17] = pk[15]; m[18] = -(pk[10] + pk[14]); m[19] = -(pk[11] + pk[15]);
}

// Ordinary pair between a bare-charge site B ("me": charge only, never polarized) and a full site A ("other"),
// d = r_A - r_B (minimum image).  For this pair class the interaction of calculatePmeDirectElectrostaticPairIxn
// (:4335-4920; no cutoff :1331-1893) collapses to the potential and the field of A's moments at B, in Cartesian form
// and without a pair frame:
//     U     = k qB [ phi_perm(B) + 1/2 phi_mu(B) ]                 (induced dipoles carry the usual 1/2 in the energy)
//     F_B   = k qB [ E_perm(B) + E_mu(B) ] = -F_A
//     tau_A = d x k qB E_perm(B)   (+ d x k qB E_mu(B) on anisotropic sites, :4888-4897)
// with phi = bn0 c - bn1 (p.d) + bn2 (d.Q.d) - bn3 (O:ddd) and E from fixedFieldDirected / inducedFieldDirected.
// The torque needs no multipole-by-multipole formula: a point charge feels none, so tau_A = -R x F_B with R = -d.
// Permanent moments are undamped; the induced dipole is Thole-damped with the default width when both sites have a
// damping factor (invDamp = 1/(damp_A damp_B), 0 otherwise) -- the same factors as the field kernels (:2693-2701),
// whose radial derivative is what the reference's dthole_c encodes.
// mA = {c, dx,dy,dz, Qxx..Qzz, Oxxx..Ozzz} from unpackPairMoments (traceless, as the energy routine sees them),
// u = induced dipole of A.
// Outputs: fB[3] force on B, tqA[3] torque on A, without the Coulomb constant and without qB (caller scales);
// returns phi_perm + phi_mu/2.
template <typename T, bool EWALD>
MPID_HD T chargeSitePair(const T* mA, T ux, T uy, T uz, T invDamp, bool anisoA, T dx, T dy, T dz, T r2,
                         T alphaEwald, T defaultThole, T* fB, T* tqA) {
    const T rinv = t_rsqrt(r2);
    const T r = r2*rinv, rinv2 = rinv*rinv;
    T c[4], bn0;
    T bare3 = rinv*rinv2, bare5 = T(3)*bare3*rinv2;
    if (EWALD) {
        // full-accuracy erfc / exp here (not the fast field-kernel primitives): these terms enter the ENERGY, a sum
        // of large cancelling pair terms in which a systematic 3e-7 relative error of erfc shows up at the 1e-6 level
        const T x = alphaEwald*r;
        const T ex2 = t_exp(-(x*x));
        bn0 = t_erfc(x)*rinv;
        const T alsq2 = T(2)*alphaEwald*alphaEwald;
        T a2n = T(1.0/MPID_SQRT_PI)/alphaEwald;
        T bn = bn0, fac = T(1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; k++) {
            a2n *= alsq2;
            bn = (fac*bn + a2n*ex2)*rinv2;
            c[k] = bn;
            fac += T(2);
        }
    } else {
        bn0 = rinv;
        c[0] = bare3; c[1] = bare5; c[2] = T(5)*bare5*rinv2; c[3] = T(7)*c[2]*rinv2;
    }
    T cu[2] = {c[0], c[1]};
    if (invDamp != T(0)) {
        const T au = defaultThole*r*invDamp;
        if (au < T(50)) {
            const T ex = t_expneg(-au);
            const T p3 = T(1) + au + T(0.5)*au*au;
            cu[0] -= ex*p3*bare3;
            cu[1] -= ex*(p3 + au*au*au*T(1.0/6.0))*bare5;
        }
    }
    T ex_ = T(0), ey_ = T(0), ez_ = T(0), ix = T(0), iy = T(0), iz = T(0);
    fixedFieldDirected<T>(mA, dx, dy, dz, c, ex_, ey_, ez_);
    inducedFieldDirected<T>(ux, uy, uz, dx, dy, dz, cu, ix, iy, iz);
    // potential: the contractions are recomputed here (cheap) rather than threaded out of fixedFieldDirected
    const T dd = mA[1]*dx + mA[2]*dy + mA[3]*dz;
    const T qx = mA[4]*dx + mA[5]*dy + mA[6]*dz, qy = mA[5]*dx + mA[7]*dy + mA[8]*dz, qz = mA[6]*dx + mA[8]*dy + mA[9]*dz;
    const T qdd = qx*dx + qy*dy + qz*dz;
    const T oxx = mA[10]*dx + mA[11]*dy + mA[12]*dz, oxy = mA[11]*dx + mA[13]*dy + mA[14]*dz, oxz = mA[12]*dx + mA[14]*dy + mA[15]*dz;
    const T oyy = mA[13]*dx + mA[16]*dy + mA[17]*dz, oyz = mA[14]*dx + mA[17]*dy + mA[18]*dz, ozz = mA[15]*dx + mA[18]*dy + mA[19]*dz;
    const T oddd = (oxx*dx + oxy*dy + oxz*dz)*dx + (oxy*dx + oyy*dy + oyz*dz)*dy + (oxz*dx + oyz*dy + ozz*dz)*dz;
    const T phi = bn0*mA[0] - c[0]*dd + c[1]*qdd - c[2]*oddd - T(0.5)*cu[0]*(ux*dx + uy*dy + uz*dz);
    fB[0] = ex_ + ix; fB[1] = ey_ + iy; fB[2] = ez_ + iz;
    const T tx = anisoA ? fB[0] : ex_, ty = anisoA ? fB[1] : ey_, tz = anisoA ? fB[2] : ez_;
    tqA[0] = dy*tz - dz*ty; tqA[1] = dz*tx - dx*tz; tqA[2] = dx*ty - dy*tx;
    return phi;
}

// =====================================================================================================
// Pair energy / force / torque in the quasi-internal (QI) frame
// =====================================================================================================
// Rotate the packed moments of one atom (packPairMoments) into the pair frame with axes X,Y,Z (Z along
// the inter-atomic vector) and return the 16 real-spherical components in the reference's order.
template <typename T> MPID_HD void momentsToQI(const T* pk, const V3<T>& X, const V3<T>& Y, const V3<T>& Z, T* Q) {
    Q[0] = pk[0];
    V3<T> d = mk<T>(pk[1], pk[2], pk[3]);
    Q[1] = dot(Z, d); Q[2] = dot(X, d); Q[3] = dot(Y, d);
    {
        T xx = pk[4], xy = pk[5], xz = pk[6], yy = pk[7], yz = pk[8], zz = -(xx + yy);
        V3<T> vz = mk<T>(xx*Z.x + xy*Z.y + xz*Z.z, xy*Z.x + yy*Z.y + yz*Z.z, xz*Z.x + yz*Z.y + zz*Z.z);
        V3<T> vx = mk<T>(xx*X.x + xy*X.y + xz*X.z, xy*X.x + yy*X.y + yz*X.z, xz*X.x + yz*X.y + zz*X.z);
        T qzz = dot(Z, vz), qxz = dot(X, vz), qyz = dot(Y, vz), qxx = dot(X, vx), qxy = dot(Y, vx);
        T qyy = -(qxx + qzz);
        tracelessToSph2<T>(qzz, qxz, qyz, qxx - qyy, qxy, Q + 4);
    }
    {
        T xxx = pk[9], xxy = pk[10], xxz = pk[11], xyy = pk[12], xyz = pk[13], yyy = pk[14], yyz = pk[15];
        T xzz = -(xxx + xyy), yzz = -(xxy + yyy), zzz = -(xxz + yyz);
        // M^Z = O . Z
        T zxx = xxx*Z.x + xxy*Z.y + xxz*Z.z, zxy = xxy*Z.x + xyy*Z.y + xyz*Z.z, zxz = xxz*Z.x + xyz*Z.y + xzz*Z.z;
        T zyy = xyy*Z.x + yyy*Z.y + yyz*Z.z, zyz = xyz*Z.x + yyz*Z.y + yzz*Z.z, zzz_ = xzz*Z.x + yzz*Z.y + zzz*Z.z;
        V3<T> w = mk<T>(zxx*Z.x + zxy*Z.y + zxz*Z.z, zxy*Z.x + zyy*Z.y + zyz*Z.z, zxz*Z.x + zyz*Z.y + zzz_*Z.z);
        V3<T> u = mk<T>(zxx*X.x + zxy*X.y + zxz*X.z, zxy*X.x + zyy*X.y + zyz*X.z, zxz*X.x + zyz*X.y + zzz_*X.z);
        T ozzz = dot(Z, w), oxzz = dot(X, w), oyzz = dot(Y, w);
        T oxxz = dot(X, u), oxyz = dot(Y, u);
        T oyyz = -(oxxz + ozzz);
        // M^X = O . X, then (M^X . X)
        T axx = xxx*X.x + xxy*X.y + xxz*X.z, axy = xxy*X.x + xyy*X.y + xyz*X.z, axz = xxz*X.x + xyz*X.y + xzz*X.z;
        T ayy = xyy*X.x + yyy*X.y + yyz*X.z, ayz = xyz*X.x + yyz*X.y + yzz*X.z, azz = xzz*X.x + yzz*X.y + zzz*X.z;
        V3<T> t = mk<T>(axx*X.x + axy*X.y + axz*X.z, axy*X.x + ayy*X.y + ayz*X.z, axz*X.x + ayz*X.y + azz*X.z);
        T oxxx = dot(X, t), oxxy = dot(Y, t);
        T oxyy = -(oxxx + oxzz), oyyy = -(oxxy + oyzz);
        tracelessToSph3<T>(ozzz, oxzz, oyzz, oxxz, oyyz, oxyz, oxxx, oxyy, oxxy, oyyy, Q + 9);
    }
}

// Torque intermediates of the QI frame: rotation generators about the frame's x, y, z axes acting on the
// 16-vector Q, contracted with the potential-derivative vector V (:4442-4457, :4816-4838).  VM flags the
// components of V that can be non-zero; the others are skipped at compile time.
#define MPID_GEN_TERM(k, expr) if (VM & (1u << (k))) acc += (expr)*V[k];
template <typename T, unsigned VM> MPID_HD T qiGenX(const T* Q, const T* V) {
    const T s3 = T(1.7320508075688772), s6 = T(2.4494897427831779), s52 = T(1.5811388300841898), s32 = T(1.2247448713915890);
    T acc = T(0);
    MPID_GEN_TERM(1, Q[3]) MPID_GEN_TERM(3, -Q[1])
    MPID_GEN_TERM(4, s3*Q[6]) MPID_GEN_TERM(5, Q[8]) MPID_GEN_TERM(6, -(s3*Q[4] + Q[7])) MPID_GEN_TERM(7, Q[6]) MPID_GEN_TERM(8, -Q[5])
    MPID_GEN_TERM(9, s6*Q[11]) MPID_GEN_TERM(10, s52*Q[13]) MPID_GEN_TERM(11, -(s6*Q[9] + s52*Q[12]))
    MPID_GEN_TERM(12, s52*Q[11] + s32*Q[15]) MPID_GEN_TERM(13, -(s52*Q[10] + s32*Q[14])) MPID_GEN_TERM(14, s32*Q[13]) MPID_GEN_TERM(15, -s32*Q[12])
    return acc;
}
template <typename T, unsigned VM> MPID_HD T qiGenY(const T* Q, const T* V) {
    const T s3 = T(1.7320508075688772), s6 = T(2.4494897427831779), s52 = T(1.5811388300841898), s32 = T(1.2247448713915890);
    T acc = T(0);
    MPID_GEN_TERM(1, -Q[2]) MPID_GEN_TERM(2, Q[1])
    MPID_GEN_TERM(4, -s3*Q[5]) MPID_GEN_TERM(5, s3*Q[4] - Q[7]) MPID_GEN_TERM(6, -Q[8]) MPID_GEN_TERM(7, Q[5]) MPID_GEN_TERM(8, Q[6])
    MPID_GEN_TERM(9, -s6*Q[10]) MPID_GEN_TERM(10, s6*Q[9] - s52*Q[12]) MPID_GEN_TERM(11, -s52*Q[13])
    MPID_GEN_TERM(12, s52*Q[10] - s32*Q[14]) MPID_GEN_TERM(13, s52*Q[11] - s32*Q[15]) MPID_GEN_TERM(14, s32*Q[12]) MPID_GEN_TERM(15, s32*Q[13])
    return acc;
}
template <typename T, unsigned VM> MPID_HD T qiGenZ(const T* Q, const T* V) {
    T acc = T(0);
    MPID_GEN_TERM(2, -Q[3]) MPID_GEN_TERM(3, Q[2])
    MPID_GEN_TERM(5, -Q[6]) MPID_GEN_TERM(6, Q[5]) MPID_GEN_TERM(7, T(-2)*Q[8]) MPID_GEN_TERM(8, T(2)*Q[7])
    MPID_GEN_TERM(10, -Q[11]) MPID_GEN_TERM(11, Q[10]) MPID_GEN_TERM(12, T(-2)*Q[13]) MPID_GEN_TERM(13, T(2)*Q[12])
    MPID_GEN_TERM(14, T(-3)*Q[15]) MPID_GEN_TERM(15, T(3)*Q[14])
    return acc;
}
#undef MPID_GEN_TERM

struct PairParams {
    double alphaEwald;
    double defaultThole;
    double mScale, pScale;     // 1,1 for ordinary pairs; 0,0 for 1-2/1-3; scale14 for 1-4
};

// One pair (I,J), d = r_J - r_I.  Returns the energy; force is the force on J (I gets the opposite),
// tqI / tqJ are lab-frame torques.  U are the lab-frame induced dipoles.
//   reference: calculatePmeDirectElectrostaticPairIxn (:4335-4920) and calculateElectrostaticPairIxn
//   (:1331-1893); the no-cutoff routine is the alpha -> 0 limit (every bVec and X term vanishes).
// SI / SJ ("simple" site): the atom carries a charge only -- no permanent dipole/quadrupole/octopole and no
// induced dipole -- so every term that multiplies one of those vanishes identically and is compiled out.
template <typename T, bool EWALD, bool MUTUAL, bool SI = false, bool SJ = false>
MPID_HD T pairElectrostatics(const T* pkI, const T* pkJ, const T* uI, const T* uJ,
                             T dampI, T dampJ, T tholeI, T tholeJ, bool anisoI, bool anisoJ,
                             T dx, T dy, T dz, T r2, T alphaEwald, T defaultThole, T mScale, T pScale,
                             T* force, T* tqI, T* tqJ) {
    const T r = t_sqrt(r2);
    const T rInv = T(1)/r;
    // pair frame (formQIRotationMatrix, :650-681): x axis = lab x made orthogonal to Z, or lab y when the
    // pair lies exactly along x.  The choice matters: the reference's induced-induced z-torque term on
    // anisotropic sites (:4877-4878) is not invariant under rotations about Z, so parity needs the same
    // axis.  1 - Zx^2 is formed as Zy^2 + Zz^2 so nearly x-aligned pairs stay accurate in FP32.
    V3<T> Z = mk<T>(dx*rInv, dy*rInv, dz*rInv);
    V3<T> X;
    if (dy != T(0) || dz != T(0)) {
        T s2 = Z.y*Z.y + Z.z*Z.z;
        T sInv = T(1)/t_sqrt(s2);
        X = mk<T>(s2*sInv, -Z.x*Z.y*sInv, -Z.x*Z.z*sInv);
    } else {
        T s2 = Z.x*Z.x + Z.z*Z.z;
        T sInv = T(1)/t_sqrt(s2);
        X = mk<T>(-Z.x*Z.y*sInv, s2*sInv, -Z.y*Z.z*sInv);
    }
    V3<T> Y = cross(Z, X);

    T QI[16], QJ[16];
    if (SI) { QI[0] = pkI[0]; for (int k = 1; k < 16; k++) QI[k] = T(0); } else momentsToQI<T>(pkI, X, Y, Z, QI);
    if (SJ) { QJ[0] = pkJ[0]; for (int k = 1; k < 16; k++) QJ[k] = T(0); } else momentsToQI<T>(pkJ, X, Y, Z, QJ);
    // induced dipoles, with the factor 1/2 of the reference folded in (:4381-4401)
    T UI[3] = {T(0), T(0), T(0)}, UJ[3] = {T(0), T(0), T(0)};
    if (!SI) {
        V3<T> a = mk<T>(uI[0], uI[1], uI[2]);
        UI[0] = T(0.5)*dot(Z, a); UI[1] = T(0.5)*dot(X, a); UI[2] = T(0.5)*dot(Y, a);
    }
    if (!SJ) {
        V3<T> b = mk<T>(uJ[0], uJ[1], uJ[2]);
        UJ[0] = T(0.5)*dot(Z, b); UJ[1] = T(0.5)*dot(X, b); UJ[2] = T(0.5)*dot(Y, b);
    }

    // ---- radial functions ---------------------------------------------------------------------------
    T ri[9];
    ri[1] = T(MPID_ELECTRIC)*rInv;
    for (int i = 2; i < 9; i++) ri[i] = ri[i-1]*rInv;
    // B[k] = mScale + bVec[k]  (:4488-4496); xKX = (alpha r)^K * X with X = 2 exp(-(alpha r)^2)/sqrt(pi)
    T B1, B2, B3, B4, B5, x2 = T(0), x3X = T(0), x5X = T(0), x7X = T(0), x9X = T(0);
    if (EWALD) {
        T x = alphaEwald*r;
        x2 = x*x;
        T X0 = T(2.0/MPID_SQRT_PI)*t_exp(-x2);
        T xX = x*X0;
        x3X = xX*x2; x5X = x3X*x2; x7X = x5X*x2; x9X = x7X*x2;
        B1 = (mScale - T(1)) + t_erfc(x);
        B2 = B1 + xX;
        B3 = B2 + T(2.0/3.0)*x3X;
        B4 = B3 + T(4.0/15.0)*x5X;
        B5 = B4 + T(8.0/105.0)*x7X;
    } else {
        B1 = B2 = B3 = B4 = B5 = mScale;
    }
    // Thole complements tc = 1 - thole_* = exp(-au) * poly (:4499-4524)
    T tc_c = T(0), tc_d0 = T(0), tc_d1 = T(0), tc_q0 = T(0), tc_q1 = T(0), tc_o0 = T(0), tc_o1 = T(0);
    T dc_c = T(0), dc_d0 = T(0), dc_d1 = T(0), dc_q0 = T(0), dc_q1 = T(0), dc_o0 = T(0), dc_o1 = T(0);
    {
        T dmp = dampI*dampJ;
        T a = (pScale == T(0)) ? tholeI + tholeJ : defaultThole;
        T u = t_abs(dmp) > T(1.0e-5) ? r/dmp : T(1e10);
        T au = a*u;
        if (au < T(50)) {
            T ex = t_exp(-au);
            T au2 = au*au, au3 = au2*au, au4 = au3*au, au5 = au4*au, au6 = au5*au;
            T p2 = T(1) + au + T(0.5)*au2;
            T p3 = p2 + au3*T(1.0/6.0);
            T p4 = p3 + au4*T(1.0/24.0);
            tc_c  = ex*p2;
            tc_d0 = ex*(p2 + au3*T(0.25));
            tc_d1 = ex*p2;
            tc_q0 = ex*(p3 + au4*T(1.0/18.0));
            tc_q1 = ex*p3;
            tc_o0 = ex*(p4 + au5*T(1.0/120.0));
            tc_o1 = ex*(p3 + au4*T(1.0/30.0));
            dc_c  = ex*(p2 + au3*T(0.25));
            dc_d0 = ex*(p3 + au4*T(1.0/12.0));
            dc_d1 = ex*p3;
            dc_q0 = ex*(p4 + au5*T(1.0/72.0));
            dc_q1 = ex*p4;
            dc_o0 = ex*(p4 + au5*T(1.0/120.0) + au6*T(1.0/600.0));
            dc_o1 = ex*(p3 + au4*T(0.04) + au5*T(1.0/150.0));
        }
    }
    // (pScale*thole + bVec[k]) = (pScale - mScale) + B_k - pScale*tc ; uScale == 1 for the U-U block
    const T dps = pScale - mScale;
#define MPID_UB(Bk, tc) (dps + (Bk) - pScale*(tc))
#define MPID_UUB(Bk, tc) ((T(1) - mScale) + (Bk) - (tc))

    T Vij[16], Vji[16], VijR[16], VjiR[16], Vijd[3], Vjid[3];
    for (int i = 0; i < 16; i++) { Vij[i] = Vji[i] = VijR[i] = VjiR[i] = T(0); }
    Vijd[0] = Vijd[1] = Vijd[2] = Vjid[0] = Vjid[1] = Vjid[2] = T(0);

    // A term "coef * QJ[k]" exists only if site J has that moment (k == 0, or J not simple); likewise for I and
    // for the induced dipoles.  The indices are literals, so these tests fold at compile time.
#define MPID_HASJ(k) (!SJ || (k) == 0)
#define MPID_HASI(k) (!SI || (k) == 0)
    // same-rank block, component a
#define MPID_SAME(a, E, D) { if (MPID_HASJ(a)) { Vij[a] += (E)*QJ[a]; VijR[a] += (D)*QJ[a]; } \
                             if (MPID_HASI(a)) { Vji[a] += (E)*QI[a]; VjiR[a] += (D)*QI[a]; } }
    // induced dipoles riding on a dipole-dipole block (slot k of the induced dipole)
#define MPID_SAME_U(a, k, EU, DU) { if (!SJ) { Vij[a] += (EU)*UJ[k]; VijR[a] += (DU)*UJ[k]; Vijd[k] += (EU)*QJ[a]; } \
                                    if (!SI) { Vji[a] += (EU)*UI[k]; VjiR[a] += (DU)*UI[k]; Vjid[k] += (EU)*QI[a]; } }
    // mixed-rank block: a belongs to the lower rank, b to the higher one; S1/S2 are the two parities
#define MPID_CROSS(a, b, S1, S2, E, D) { if (MPID_HASJ(b)) { Vij[a] += (S1)*(E)*QJ[b]; VijR[a] += (S1)*(D)*QJ[b]; } \
                                         if (MPID_HASI(a)) { Vji[b] += (S1)*(E)*QI[a]; VjiR[b] += (S1)*(D)*QI[a]; } \
                                         if (MPID_HASJ(a)) { Vij[b] += (S2)*(E)*QJ[a]; VijR[b] += (S2)*(D)*QJ[a]; } \
                                         if (MPID_HASI(b)) { Vji[a] += (S2)*(E)*QI[b]; VjiR[a] += (S2)*(D)*QI[b]; } }
    // induced dipole in the lower-rank slot (dipole-quadrupole, dipole-octopole blocks)
#define MPID_CROSS_U_LO(b, k, S1, S2, EU, DU) { if (!SI && MPID_HASJ(b)) Vijd[k] += (S1)*(EU)*QJ[b]; \
                                                if (!SI) { Vji[b] += (S1)*(EU)*UI[k]; VjiR[b] += (S1)*(DU)*UI[k]; } \
                                                if (!SJ) { Vij[b] += (S2)*(EU)*UJ[k]; VijR[b] += (S2)*(DU)*UJ[k]; } \
                                                if (!SJ && MPID_HASI(b)) Vjid[k] += (S2)*(EU)*QI[b]; }
    // induced dipole in the higher-rank slot (charge-dipole block)
#define MPID_CROSS_U_HI(a, k, S1, S2, EU, DU) { if (!SJ) { Vij[a] += (S1)*(EU)*UJ[k]; VijR[a] += (S1)*(DU)*UJ[k]; Vjid[k] += (S1)*(EU)*QI[a]; } \
                                                if (!SI) { Vijd[k] += (S2)*(EU)*QJ[a]; Vji[a] += (S2)*(EU)*UI[k]; VjiR[a] += (S2)*(DU)*UI[k]; } }
    const T P = T(1), M = T(-1);
    T e, d_, eU, dU;
    // charge-charge
    e = ri[1]*B1; d_ = T(-0.5)*B2*ri[2];
    MPID_SAME(0, e, d_)
    // charge-dipole (m=0)
    e = ri[2]*B2; d_ = -ri[3]*(B2 + x3X);
    eU = T(2)*ri[2]*MPID_UB(B2, tc_c); dU = T(-4)*ri[3]*(MPID_UB(B2, dc_c) + x3X);
    MPID_CROSS(0, 1, M, P, e, d_)
    MPID_CROSS_U_HI(0, 0, M, P, eU, dU)
    // dipole-dipole (m=0)
    e = T(-2.0/3.0)*ri[3]*(T(3)*B3 + x3X); d_ = ri[4]*(T(3)*B3 + T(2)*x5X);
    eU = T(-4.0/3.0)*ri[3]*(T(3)*MPID_UB(B3, tc_d0) + x3X); dU = T(2)*ri[4]*(T(6)*MPID_UB(B3, dc_d0) + T(4)*x5X);
    MPID_SAME(1, e, d_)
    MPID_SAME_U(1, 0, eU, dU)
    // dipole-dipole (m=1)
    e = ri[3]*(B3 - T(2.0/3.0)*x3X); d_ = T(-1.5)*ri[4]*B3;
    eU = T(2)*ri[3]*(MPID_UB(B3, tc_d1) - T(2.0/3.0)*x3X); dU = T(-6)*ri[4]*MPID_UB(B3, dc_d1);
    MPID_SAME(2, e, d_) MPID_SAME(3, e, d_)
    MPID_SAME_U(2, 1, eU, dU) MPID_SAME_U(3, 2, eU, dU)
    // charge-quadrupole (m=0)
    e = B3*ri[3]; d_ = T(-1.0/3.0)*ri[4]*(T(4.5)*B3 + T(2)*x5X);
    MPID_CROSS(0, 4, P, P, e, d_)
    // dipole-quadrupole (m=0)
    e = ri[4]*(T(3)*B3 + T(4.0/3.0)*x5X); d_ = T(-4.0/3.0)*ri[5]*(T(4.5)*B3 + (T(1) + x2)*x5X);
    eU = T(2)*ri[4]*(T(3)*MPID_UB(B3, tc_q0) + T(4.0/3.0)*x5X); dU = T(-8.0/3.0)*ri[5]*(T(9)*MPID_UB(B3, dc_q0) + T(2)*(T(1) + x2)*x5X);
    MPID_CROSS(1, 4, P, M, e, d_)
    MPID_CROSS_U_LO(4, 0, P, M, eU, dU)
    // dipole-quadrupole (m=1)
    e = T(-1.7320508075688772)*ri[4]*B3; d_ = T(2.3094010767585030)*ri[5]*(T(1.5)*B3 + T(0.5)*x5X);
    eU = T(-3.4641016151377544)*ri[4]*MPID_UB(B3, tc_q1); dU = T(4.6188021535170060)*ri[5]*(T(3)*MPID_UB(B3, dc_q1) + x5X);
    MPID_CROSS(2, 5, P, M, e, d_) MPID_CROSS(3, 6, P, M, e, d_)
    MPID_CROSS_U_LO(5, 1, P, M, eU, dU) MPID_CROSS_U_LO(6, 2, P, M, eU, dU)
    // quadrupole-quadrupole (m=0,1,2)
    e = ri[5]*(T(6)*B4 + T(4.0/45.0)*(T(-3) + T(10)*x2)*x5X); d_ = T(-1.0/9.0)*ri[6]*(T(135)*B4 + T(4)*(T(1) + T(2)*x2)*x7X);
    MPID_SAME(4, e, d_)
    e = T(-4.0/15.0)*ri[5]*(T(15)*B4 + x5X); d_ = ri[6]*(T(10)*B4 + T(4.0/3.0)*x7X);
    MPID_SAME(5, e, d_) MPID_SAME(6, e, d_)
    e = ri[5]*(B4 - T(4.0/15.0)*x5X); d_ = T(-2.5)*B4*ri[6];
    MPID_SAME(7, e, d_) MPID_SAME(8, e, d_)
    // charge-octopole (m=0)
    e = ri[4]*(-B3 - T(4.0/15.0)*x5X); d_ = T(2.0/15.0)*ri[5]*(T(15)*B3 + T(2)*(T(2)*x5X + x7X));
    MPID_CROSS(0, 9, P, M, e, d_)
    // dipole-octopole (m=0)
    e = T(-4)*ri[5]*(B4 + T(2.0/15.0)*x7X); d_ = T(2.0/15.0)*ri[6]*(T(75)*B4 + T(4)*(T(1) + x2)*x7X);
    eU = T(-8)*ri[5]*(MPID_UB(B4, tc_o0) + T(2.0/15.0)*x7X); dU = T(8.0/15.0)*ri[6]*(T(75)*MPID_UB(B4, dc_o0) + T(4)*(T(1) + x2)*x7X);
    MPID_CROSS(1, 9, P, P, e, d_)
    MPID_CROSS_U_LO(9, 0, P, P, eU, dU)
    // dipole-octopole (m=1)
    e = T(2.4494897427831779)*B4*ri[5]; d_ = T(-0.081649658092772609)*ri[6]*(T(75)*B4 + T(8)*x7X);
    eU = T(4.8989794855663558)*MPID_UB(B4, tc_o1)*ri[5]; dU = T(-0.32659863237109044)*ri[6]*(T(75)*MPID_UB(B4, dc_o1) + T(8)*x7X);
    MPID_CROSS(2, 10, P, P, e, d_) MPID_CROSS(3, 11, P, P, e, d_)
    MPID_CROSS_U_LO(10, 1, P, P, eU, dU) MPID_CROSS_U_LO(11, 2, P, P, eU, dU)
    // quadrupole-octopole (m=0,1,2)
    e = ri[6]*(T(-10)*B4 - T(8.0/45.0)*(T(3) + T(2)*x2)*x7X); d_ = T(2.0/45.0)*ri[7]*(T(675)*B4 + T(2)*(T(27) + T(4)*x2*x2)*x7X);
    MPID_CROSS(4, 9, P, M, e, d_)
    e = T(7.0710678118654752)*ri[6]*(B4 + T(8.0/75.0)*x7X); d_ = T(-0.094280904158206336)*ri[7]*(T(225)*B4 + T(8)*(T(2) + x2)*x7X);
    MPID_CROSS(5, 10, P, M, e, d_) MPID_CROSS(6, 11, P, M, e, d_)
    e = T(-2.2360679774997897)*B4*ri[6]; d_ = T(0.14907119849998598)*ri[7]*(T(45)*B4 + T(4)*x7X);
    MPID_CROSS(7, 12, P, M, e, d_) MPID_CROSS(8, 13, P, M, e, d_)
    // octopole-octopole (m=0..3)
    e = ri[7]*(T(-20)*B5 - T(8.0/1575.0)*(T(15) + T(28)*x2 + T(28)*x2*x2)*x7X);
    d_ = T(2.0/225.0)*ri[8]*(T(7875)*B5 + T(4)*(T(41) - T(4)*x2 + T(4)*x2*x2)*x9X);
    MPID_SAME(9, e, d_)
    e = ri[7]*(T(15)*B5 + T(8.0/525.0)*(T(-5) + T(28)*x2)*x7X); d_ = T(-1.0/150.0)*ri[8]*(T(7875)*B5 + T(32)*(T(3) + T(2)*x2)*x9X);
    MPID_SAME(10, e, d_) MPID_SAME(11, e, d_)
    e = ri[7]*(T(-6)*B5 - T(8.0/105.0)*x7X); d_ = T(0.5)*ri[8]*(T(42)*B5 + T(16.0/15.0)*x9X);
    MPID_SAME(12, e, d_) MPID_SAME(13, e, d_)
    e = ri[7]*(B5 - T(8.0/105.0)*x7X); d_ = T(-3.5)*B5*ri[8];
    MPID_SAME(14, e, d_) MPID_SAME(15, e, d_)

    // ---- energy, radial force and torque intermediates (:4799-4838) ---------------------------------
    T energy = T(0), fIZ = T(0), fJZ = T(0);
    constexpr unsigned VMI = SJ ? 0x213u : 0xFFFFu, VMJ = SI ? 0x213u : 0xFFFFu;   // components of Vij / Vji that can be non-zero
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if ((!SI || i == 0) && ((VMI >> i) & 1u)) { energy += QI[i]*Vij[i]; fIZ += QI[i]*VijR[i]; }
        if ((!SJ || i == 0) && ((VMJ >> i) & 1u)) { energy += QJ[i]*Vji[i]; fJZ += QJ[i]*VjiR[i]; }
    }
    energy *= T(0.5);
    // rotation generators about the frame's x, y, z axes acting on the 16-vector, contracted with V (qiGenX/Y/Z);
    // when the partner is a bare charge only the m = 0 components of V exist (mask 0x213 = {0,1,4,9}), which also
    // makes the |m| >= 2 components of Q dead code
    T EIX = T(0), EIY = T(0), EIZ = T(0), EJX = T(0), EJY = T(0), EJZ = T(0);
    if (!SI) { EIX = qiGenX<T, VMI>(QI, Vij); EIY = qiGenY<T, VMI>(QI, Vij); EIZ = qiGenZ<T, VMI>(QI, Vij); }
    if (!SJ) { EJX = qiGenX<T, VMJ>(QJ, Vji); EJY = qiGenY<T, VMJ>(QJ, Vji); EJZ = qiGenZ<T, VMJ>(QJ, Vji); }
    // the same for the induced dipoles against the field of the permanent moments only
    T iEIX = UI[2]*Vijd[0] - UI[0]*Vijd[2], iEJX = UJ[2]*Vjid[0] - UJ[0]*Vjid[2];
    T iEIY = UI[0]*Vijd[1] - UI[1]*Vijd[0], iEJY = UJ[0]*Vjid[1] - UJ[1]*Vjid[0];
    T iEIZ = UI[1]*Vijd[2] - UI[2]*Vijd[1], iEJZ = UJ[1]*Vjid[2] - UJ[2]*Vjid[1];
    if (MUTUAL && !SI && !SJ) {   // induced-induced coupling (:4860-4881)
        T eC = T(-8.0/3.0)*ri[3]*(T(3)*MPID_UUB(B3, tc_d0) + x3X);
        T dC = T(2)*ri[4]*(T(6)*MPID_UUB(B3, dc_d0) + T(4)*x5X);
        iEIX += eC*UI[2]*UJ[0]; iEJX += eC*UJ[2]*UI[0];
        iEIY -= eC*UI[1]*UJ[0]; iEJY -= eC*UJ[1]*UI[0];
        fIZ += dC*UI[0]*UJ[0];  fJZ += dC*UJ[0]*UI[0];
        eC = T(4)*ri[3]*(MPID_UUB(B3, tc_d1) - T(2.0/3.0)*x3X);
        dC = T(-6)*ri[4]*MPID_UUB(B3, dc_d1);
        iEIX -= eC*UI[0]*UJ[2]; iEJX -= eC*UJ[0]*UI[2];
        iEIY += eC*UI[0]*UJ[1]; iEJY += eC*UJ[0]*UI[1];
        iEIZ += eC*UI[1]*UJ[2]; iEJZ += eC*UJ[1]*UI[2];
        T uu = UI[1]*UJ[1] + UI[2]*UJ[2];
        fIZ += dC*uu; fJZ += dC*uu;
    }
    // frame-local force and torques, then back to the lab axes (:4883-4917)
    T fx = rInv*(EIY + EJY + iEIY + iEJY), fy = -rInv*(EIX + EJX + iEIX + iEJX), fz = -(fJZ + fIZ);
    T tIx = -EIX, tIy = -EIY, tIz = -EIZ, tJx = -EJX, tJy = -EJY, tJz = -EJZ;
    if (anisoI) { tIx -= iEIX; tIy -= iEIY; tIz -= iEIZ; }
    if (anisoJ) { tJx -= iEJX; tJy -= iEJY; tJz -= iEJZ; }
    force[0] = X.x*fx + Y.x*fy + Z.x*fz; force[1] = X.y*fx + Y.y*fy + Z.y*fz; force[2] = X.z*fx + Y.z*fy + Z.z*fz;
    tqI[0] = X.x*tIx + Y.x*tIy + Z.x*tIz; tqI[1] = X.y*tIx + Y.y*tIy + Z.y*tIz; tqI[2] = X.z*tIx + Y.z*tIy + Z.z*tIz;
    tqJ[0] = X.x*tJx + Y.x*tJy + Z.x*tJz; tqJ[1] = X.y*tJx + Y.y*tJy + Z.y*tJz; tqJ[2] = X.z*tJx + Y.z*tJy + Z.z*tJz;
    return energy;
#undef MPID_UB
#undef MPID_UUB
#undef MPID_HASI
#undef MPID_HASJ
#undef MPID_SAME
#undef MPID_SAME_U
#undef MPID_CROSS
#undef MPID_CROSS_U_LO
#undef MPID_CROSS_U_HI
}

// =====================================================================================================
// PME: order-6 B-splines, fractional-coordinate transforms, reciprocal-space per-atom terms
// =====================================================================================================
#define MPID_PME_ORDER 6

// Index of d^(t+u+v) phi / dx^t dy^u dz^v in the 35-vector of computeFixedPotentialFromGrid (:3494-3528).
MPID_HD int phiIndex(int t, int u, int v) {
    // key = 25 t + 5 u + v
    switch (25*t + 5*u + v) {
        case 0: return 0;
        case 25: return 1;  case 5: return 2;   case 1: return 3;
        case 50: return 4;  case 10: return 5;  case 2: return 6;  case 30: return 7;  case 26: return 8;  case 6: return 9;
        case 75: return 10; case 15: return 11; case 3: return 12; case 55: return 13; case 51: return 14; case 35: return 15;
        case 11: return 16; case 27: return 17; case 7: return 18; case 31: return 19;
        case 100: return 20; case 20: return 21; case 4: return 22; case 80: return 23; case 76: return 24; case 40: return 25;
        case 16: return 26; case 28: return 27; case 8: return 28; case 60: return 29; case 52: return 30; case 12: return 31;
        case 56: return 32; case 36: return 33; case 32: return 34;
    }
    return -1;
}

// theta[i][k] = k-th derivative (k = 0..4) of the order-6 cardinal B-spline weight of grid point i (0..5)
// for fractional offset w in [0,1).   reference: computeBSplinePoint (:2956-3044).
template <typename T> MPID_HD void bsplineWeights(T w, T theta[MPID_PME_ORDER][5]) {
    // M[n][i], i = 0..n-1 : order-n spline; built by the standard two-term recursion
    T M[MPID_PME_ORDER + 1][MPID_PME_ORDER + 2];
    for (int n = 0; n <= MPID_PME_ORDER; n++)
        for (int i = 0; i < MPID_PME_ORDER + 2; i++) M[n][i] = T(0);
    // stored with a one-slot offset so that index -1 reads as zero
    M[2][1] = T(1) - w; M[2][2] = w;
    for (int n = 3; n <= MPID_PME_ORDER; n++) {
        T inv = T(1)/T(n - 1);
        for (int i = 0; i < n; i++) {
            // M_n(i) = ((w + n-1-i) M_{n-1}(i-1) + (i + 1 - w) M_{n-1}(i)) / (n-1)
            M[n][i+1] = inv*((w + T(n - 1 - i))*M[n-1][i] + (T(i + 1) - w)*M[n-1][i+1]);
        }
    }
#define MPID_M(n, i) (((i) < 0) ? T(0) : M[n][(i)+1])
    for (int i = 0; i < MPID_PME_ORDER; i++) {
        theta[i][0] = MPID_M(6, i);
        theta[i][1] = MPID_M(5, i-1) - MPID_M(5, i);
        theta[i][2] = MPID_M(4, i-2) - T(2)*MPID_M(4, i-1) + MPID_M(4, i);
        theta[i][3] = MPID_M(3, i-3) - T(3)*MPID_M(3, i-2) + T(3)*MPID_M(3, i-1) - MPID_M(3, i);
        theta[i][4] = MPID_M(2, i-4) - T(4)*MPID_M(2, i-3) + T(6)*MPID_M(2, i-2) - T(4)*MPID_M(2, i-1) + MPID_M(2, i);
    }
#undef MPID_M
}

// Grid cell and fractional offsets of one atom, in double exactly as computeMPIDBsplines (:3049-3075).
struct PmeGeom {
    int n[3];
    double A[3][3];     // A[j][k] = n_j * recip_k[j] : Cartesian -> scaled-fractional ("cartToFrac")
};
inline void makePmeGeom(PmeGeom& G, const Box& B, int nx, int ny, int nz) {
    G.n[0] = nx; G.n[1] = ny; G.n[2] = nz;
    const double* rv[3] = {B.ra, B.rb, B.rc};
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) G.A[j][k] = G.n[j]*rv[k][j];
}
MPID_HD void pmeAtomCell(const Box& B, const PmeGeom& G, double x, double y, double z, int* igrid, double* w) {
    periodicDelta(B, x, y, z);
    const double* rv[3] = {B.ra, B.rb, B.rc};
    for (int j = 0; j < 3; j++) {
        double f = x*rv[0][j] + y*rv[1][j] + z*rv[2][j];
        double fr = G.n[j]*(f - (int)(f + 0.5) + 0.5);
        int ifr = (int) floor(fr);
        w[j] = fr - ifr;
        int g = ifr - MPID_PME_ORDER + 1;
        igrid[j] = g + (g < 0 ? G.n[j] : 0);
    }
}

// Lab Cartesian multipoles (20, internal orders) -> scaled-fractional multipoles with the symmetry
// multiplicities folded in, order {q, d0 d1 d2, Qxx Qxy Qxz Qyy Qyz Qzz, O (internal order)}.
//   reference: transformMultipolesToFractionalCoordinates (:3077-3169)
template <typename T> MPID_HD void multipolesToFractional(const double A[3][3], const T* m, T* f) {
    f[0] = m[0];
    for (int j = 0; j < 3; j++) f[1+j] = T(A[j][0])*m[1] + T(A[j][1])*m[2] + T(A[j][2])*m[3];
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
            T s = T(0);
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) s += T(A[i][k]*A[j][l])*m[4 + symIdx2(k, l)];
            f[4 + symIdx2(i, j)] = (i == j) ? s : T(2)*s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++)
            for (int k = j; k < 3; k++) {
                T s = T(0);
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++)
                        for (int c = 0; c < 3; c++) s += T(A[i][a]*A[j][b]*A[k][c])*m[10 + symIdx3(a, b, c)];
                T mult = (i == j && j == k) ? T(1) : ((i == j || j == k) ? T(3) : T(6));
                f[10 + symIdx3(i, j, k)] = mult*s;
            }
}

// exponents (t,u,v) of the k-th entry of the fractional multipole vector above
MPID_HD void multipoleExponents(int k, int& t, int& u, int& v) {
    const int tt[20] = {0, 1,0,0, 2,1,1,0,0,0, 3,2,2,1,1,1,0,0,0,0};
    const int uu[20] = {0, 0,1,0, 0,1,0,2,1,0, 0,1,0,2,1,0,3,2,1,0};
    const int vv[20] = {0, 0,0,1, 0,0,1,0,1,2, 0,0,1,0,1,2,0,1,2,3};
    t = tt[k]; u = uu[k]; v = vv[k];
}

// Value one atom adds to one grid point: fm = fractional multipoles (nm entries used: 20 for permanent
// moments, 4 with f[0]=0 for induced dipoles), tx/ty/tz the three weight rows of that point.
//   reference: spreadFixedMultipolesOntoGrid (:3269-3327), spreadInducedDipolesOnGrid (:3532-3573)
template <typename T, bool FIXED> MPID_HD T spreadTerm(const T* f, const T* tx, const T* ty, const T* tz) {
    if (FIXED) {
        T term0 = f[0]*ty[0]*tz[0] + f[2]*ty[1]*tz[0] + f[3]*ty[0]*tz[1]
                + f[7]*ty[2]*tz[0] + f[9]*ty[0]*tz[2] + f[8]*ty[1]*tz[1]
                + f[16]*ty[3]*tz[0] + f[17]*ty[2]*tz[1] + f[18]*ty[1]*tz[2] + f[19]*ty[0]*tz[3];
        T term1 = f[1]*ty[0]*tz[0] + f[5]*ty[1]*tz[0] + f[6]*ty[0]*tz[1]
                + f[13]*ty[2]*tz[0] + f[14]*ty[1]*tz[1] + f[15]*ty[0]*tz[2];
        T term2 = f[4]*ty[0]*tz[0] + f[11]*ty[1]*tz[0] + f[12]*ty[0]*tz[1];
        T term3 = f[10]*ty[0]*tz[0];
        return term0*tx[0] + term1*tx[1] + term2*tx[2] + term3*tx[3];
    } else {
        return (f[2]*ty[1]*tz[0] + f[3]*ty[0]*tz[1])*tx[0] + f[1]*ty[0]*tz[0]*tx[1];
    }
}

// Scaled-fractional potential derivatives (up to total order 3) -> Cartesian, 20 entries
// {phi, x y z, xx yy zz xy xz yz, xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz}.
//   reference: transformPotentialToCartesianCoordinates (:3171-3267)
template <typename T> MPID_HD void potentialToCartesian(const double A[3][3], const T* fphi, T* cphi) {
    cphi[0] = fphi[0];
    for (int i = 0; i < 3; i++) cphi[1+i] = T(A[0][i])*fphi[1] + T(A[1][i])*fphi[2] + T(A[2][i])*fphi[3];
    const int second[6] = {4, 7, 8, 5, 9, 6};      // cphi slot of internal sym index xx xy xz yy yz zz
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
            T s = T(0);
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) {
                    int e[3] = {0, 0, 0}; e[k]++; e[l]++;
                    s += T(A[k][i]*A[l][j])*fphi[phiIndex(e[0], e[1], e[2])];
                }
            cphi[second[symIdx2(i, j)]] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++)
            for (int k = j; k < 3; k++) {
                T s = T(0);
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++)
                        for (int c = 0; c < 3; c++) {
                            int e[3] = {0, 0, 0}; e[a]++; e[b]++; e[c]++;
                            s += T(A[a][i]*A[b][j]*A[c][k])*fphi[phiIndex(e[0], e[1], e[2])];
                        }
                cphi[10 + symIdx3(i, j, k)] = s;
            }
}

// Torque of a Cartesian multipole set m[20] = {q, d, Qxx Qyy Qzz 2Qxy 2Qxz 2Qyz, Oxxx 3Oxxy 3Oxxz 3Oxyy
// 6Oxyz 3Oxzz Oyyy 3Oyyz 3Oyzz Ozzz} in the Cartesian potential derivatives p[20] (no Coulomb constant).
// The octopole part is the reference's trace-projected closed form.
//   reference: computeReciprocalSpaceFixedMultipoleForceAndEnergy (:3788-3836)
template <typename T> MPID_HD void reciprocalTorque(const T* m, const T* p, T* tq) {
    const T fifth = T(0.2);
    tq[0] = m[3]*p[2] - m[2]*p[3]
          + T(2)*(m[6] - m[5])*p[9] + m[8]*p[7] + m[9]*p[5] - m[7]*p[8] - m[9]*p[6]
          + fifth*(p[11]*(T(4)*m[12] - m[17] - T(3)*m[19]) + p[16]*(-m[12] + T(4)*m[17] - T(3)*m[19])
                 + p[12]*(T(-4)*m[11] + T(3)*m[16] + m[18]) + p[17]*(m[11] - T(12)*m[16] + T(11)*m[18])
                 + p[18]*(-m[12] - T(11)*m[17] + T(12)*m[19]) + p[19]*(m[11] + T(3)*m[16] - T(4)*m[18]))
          + (p[13] - p[15])*m[14] + T(2)*p[14]*(m[15] - m[13]);
    tq[1] = m[1]*p[3] - m[3]*p[1]
          + T(2)*(m[4] - m[6])*p[8] + m[7]*p[9] + m[8]*p[6] - m[8]*p[4] - m[9]*p[7]
          + fifth*(p[10]*(T(-4)*m[12] + m[17] + T(3)*m[19]) + p[13]*(m[12] - T(4)*m[17] + T(3)*m[19])
                 + p[12]*(T(12)*m[10] - m[13] - T(11)*m[15]) + p[17]*(T(-3)*m[10] + T(4)*m[13] - m[15])
                 + p[15]*(T(11)*m[12] + m[17] - T(12)*m[19]) + p[19]*(T(-3)*m[10] - m[13] + T(4)*m[15]))
          + (p[18] - p[11])*m[14] + T(2)*p[14]*(m[11] - m[18]);
    tq[2] = m[2]*p[1] - m[1]*p[2]
          + T(2)*(m[5] - m[4])*p[7] + m[7]*p[4] + m[9]*p[8] - m[7]*p[5] - m[8]*p[9]
          + fifth*(p[10]*(T(4)*m[11] - T(3)*m[16] - m[18]) + p[11]*(T(-12)*m[10] + T(11)*m[13] + m[15])
                 + p[13]*(T(-11)*m[11] + T(12)*m[16] - m[18]) + p[16]*(T(3)*m[10] - T(4)*m[13] + m[15])
                 + p[15]*(-m[11] - T(3)*m[16] + T(4)*m[18]) + p[18]*(T(3)*m[10] + m[13] - T(4)*m[15]))
          + (p[12] - p[17])*m[14] + T(2)*p[14]*(m[17] - m[12]);
}

// Build the torque multipole vector from lab Cartesian moments cart[20] (internal orders), optionally
// adding an induced dipole to the dipole slots (:3752-3786).
template <typename T> MPID_HD void torqueMultipoles(const T* cart, T ux, T uy, T uz, T* m) {
    m[0] = cart[0];
    m[1] = cart[1] + ux; m[2] = cart[2] + uy; m[3] = cart[3] + uz;
    m[4] = cart[4]; m[5] = cart[7]; m[6] = cart[9];
    m[7] = T(2)*cart[5]; m[8] = T(2)*cart[6]; m[9] = T(2)*cart[8];
    m[10] = cart[10]; m[11] = T(3)*cart[11]; m[12] = T(3)*cart[12]; m[13] = T(3)*cart[13]; m[14] = T(6)*cart[14];
    m[15] = T(3)*cart[15]; m[16] = cart[16]; m[17] = T(3)*cart[17]; m[18] = T(3)*cart[18]; m[19] = cart[19];
}

// sum_k f[k] * phi[d^(t+dt, u+du, v+dv)] for the first nk fractional multipoles (energy: dt=du=dv=0;
// force components: one of them = 1).   reference: deriv0..deriv3 tables (:3743-3746)
template <typename T> MPID_HD T contractFractional(const T* f, int nk, const T* phi, int dt, int du, int dv) {
    T s = T(0);
    for (int k = 0; k < nk; k++) {
        int t, u, v;
        multipoleExponents(k, t, u, v);
        s += f[k]*phi[phiIndex(t + dt, u + du, v + dv)];
    }
    return s;
}

// =====================================================================================================
// Torque -> forces on the frame-defining atoms
// =====================================================================================================
// fI/fZ/fX/fY receive the force increments for the atom and its z/x/y anchors.
//   reference: mapTorqueToForceForParticle (:1895-2110).  The reference takes the second direction from
//   particleData[atomX] for every axis type (:2124-2127); for a ZOnly site without an x anchor that is an
//   out-of-range read, and there -- like the plugin's CUDA platform -- we use the lab axis least aligned
//   with u.  (The choice only matters when the torque has a component along u, e.g. anisotropic mutual
//   polarization on a ZOnly site.)
MPID_HD void torqueToForce(int axisType, const double* pi, const double* pz, const double* px, const double* py, bool hasX, bool hasY,
                           const double* torque, double* fI, double* fZ, double* fX, double* fY) {
    for (int i = 0; i < 3; i++) fI[i] = fZ[i] = fX[i] = fY[i] = 0.0;
    if (axisType == NoAxisType) return;
    V3<double> tq = mk(torque[0], torque[1], torque[2]);
    V3<double> U = mk(pz[0]-pi[0], pz[1]-pi[1], pz[2]-pi[2]);
    double nU = normalize(U);
    V3<double> V;
    if (axisType == ZOnly && !hasX)
        V = (fabs(U.x) < 0.866) ? mk(1.0, 0.0, 0.0) : mk(0.0, 1.0, 0.0);
    else
        V = mk(px[0]-pi[0], px[1]-pi[1], px[2]-pi[2]);
    double nV = normalize(V);
    V3<double> W;
    if (hasY && (axisType == ZBisect || axisType == ThreeFold)) W = mk(py[0]-pi[0], py[1]-pi[1], py[2]-pi[2]);
    else W = cross(U, V);
    double nW = normalize(W);
    V3<double> UV = cross(V, U), UW = cross(W, U), VW = cross(W, V);
    normalize(UV); normalize(UW); normalize(VW);
    double cUV = dot(U, V), sUV = sqrt(1.0 - cUV*cUV);
    double cUW = dot(U, W), sUW = sqrt(1.0 - cUW*cUW);
    double cVW = dot(V, W), sVW = sqrt(1.0 - cVW*cVW);
    double dU = -dot(U, tq), dV = -dot(V, tq), dW = -dot(W, tq);
    V3<double> FZ = mk(0.0, 0.0, 0.0), FX = FZ, FY = FZ;
    if (axisType == ZThenX || axisType == Bisector) {
        double f1 = dV/(nU*sUV), f2 = dW/nU, f3 = -dU/(nV*sUV), f4 = 0.0;
        if (axisType == Bisector) { f2 *= 0.5; f4 = 0.5*dW/nV; }
        FZ = UV*f1 + UW*f2;
        FX = UV*f3 + VW*f4;
    } else if (axisType == ZBisect) {
        V3<double> R = V + W;
        V3<double> S = cross(U, R);
        normalize(R); normalize(S);
        V3<double> UR = cross(R, U), US = cross(S, U);
        normalize(UR); normalize(US);
        double cUR = dot(U, R), sUR = sqrt(1.0 - cUR*cUR);
        double cVS = dot(V, S), sVS = sqrt(1.0 - cVS*cVS);
        double cWS = dot(W, S), sWS = sqrt(1.0 - cWS*cWS);
        V3<double> t1 = V - S*cVS, t2 = W - S*cWS;
        normalize(t1); normalize(t2);
        double c1 = dot(U, t1), s1 = sqrt(1.0 - c1*c1);
        double c2 = dot(U, t2), s2 = sqrt(1.0 - c2*c2);
        double dR = -dot(R, tq), dS = -dot(S, tq);
        double f1 = dR/(nU*sUR), f2 = dS/nU, f3 = dU/(nV*(s1 + s2)), f4 = dU/(nW*(s1 + s2));
        FZ = UR*f1 + US*f2;
        FX = (S*sVS - t1*cVS)*f3;
        FY = (S*sWS - t2*cWS)*f4;
    } else if (axisType == ThreeFold) {
        FZ = (UW*(dW/(nU*sUW)) + UV*(dV/(nU*sUV)) - UW*(dU/(nU*sUW)) - UV*(dU/(nU*sUV)))*(1.0/3.0);
        FX = (VW*(dW/(nV*sVW)) - UV*(dU/(nV*sUV)) - VW*(dV/(nV*sVW)) + UV*(dV/(nV*sUV)))*(1.0/3.0);
        FY = (UW*(-dU/(nW*sUW)) - VW*(dV/(nW*sVW)) + UW*(dW/(nW*sUW)) + VW*(dW/(nW*sVW)))*(1.0/3.0);
    } else if (axisType == ZOnly) {
        FZ = UV*(dV/(nU*sUV)) + UW*(dW/nU);
    }
    V3<double> FI = FZ + FX + FY;
    fI[0] = FI.x; fI[1] = FI.y; fI[2] = FI.z;
    fZ[0] = -FZ.x; fZ[1] = -FZ.y; fZ[2] = -FZ.z;
    fX[0] = -FX.x; fX[1] = -FX.y; fX[2] = -FX.z;
    fY[0] = -FY.x; fY[1] = -FY.y; fY[2] = -FY.z;
}

} // namespace mpid
#endif
