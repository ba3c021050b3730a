// Counts the floating-point operations of the per-pair functions of mpid_math.h by instantiating them with a counting
// scalar type (every +, -, *, / = 1 flop; FMAs are never formed; exp / erfc / rsqrt / rcp listed separately) -- the way
// SURVEY.md 8(d) counted the oracle's generic routines (2,240 / 430 / 150 flop per pair), applied to the pair classes
// the oracle has no specialised routine for.  bench.py's rooflines use the printed figures.
//   g++ -O1 -std=c++17 -I mpidopenmmplugin_b200/csrc tools/count_flops.cpp -o /tmp/count_flops && /tmp/count_flops
#include <cmath>
#include <cstdio>
namespace mpid {
struct C {
    double v;
    C() : v(0) {}
    C(double x) : v(x) {}
    explicit operator double() const { return v; }
};
static long long nAdd = 0, nMul = 0, nDiv = 0, nExp = 0, nErfc = 0, nSqrt = 0;
inline C operator+(C a, C b) { nAdd++; return C(a.v + b.v); }
inline C operator-(C a, C b) { nAdd++; return C(a.v - b.v); }
inline C operator*(C a, C b) { nMul++; return C(a.v*b.v); }
inline C operator/(C a, C b) { nDiv++; return C(a.v/b.v); }
inline C operator-(C a) { return C(-a.v); }
inline C& operator+=(C& a, C b) { nAdd++; a.v += b.v; return a; }
inline C& operator-=(C& a, C b) { nAdd++; a.v -= b.v; return a; }
inline C& operator*=(C& a, C b) { nMul++; a.v *= b.v; return a; }
inline bool operator<(C a, C b) { return a.v < b.v; }
inline bool operator>(C a, C b) { return a.v > b.v; }
inline bool operator!=(C a, C b) { return a.v != b.v; }
inline bool operator==(C a, C b) { return a.v == b.v; }
inline C t_exp(C x) { nExp++; return C(std::exp(x.v)); }
inline C t_expneg(C x) { nExp++; return C(std::exp(x.v)); }
inline C t_erfc(C x) { nErfc++; return C(std::erfc(x.v)); }
inline C t_erfc_ex(C x, C) { nErfc++; return C(std::erfc(x.v)); }
inline C t_sqrt(C x) { nSqrt++; return C(std::sqrt(x.v)); }
inline C t_rsqrt(C x) { nSqrt++; return C(1.0/std::sqrt(x.v)); }
inline C t_rcp(C x) { nDiv++; return C(1.0/x.v); }
inline C t_abs(C x) { return C(std::fabs(x.v)); }
}
#include "mpid_math.h"
using namespace mpid;

static void reset() { nAdd = nMul = nDiv = nExp = nErfc = nSqrt = 0; }
static void report(const char* what) {
    printf("%-58s %4lld flop (add %lld, mul %lld, div %lld) + %lld exp, %lld erfc, %lld sqrt/rsqrt\n", what, nAdd + nMul + nDiv, nAdd, nMul, nDiv, nExp, nErfc, nSqrt);
}

int main() {
    C m[20];
    for (int k = 0; k < 20; k++) m[k] = C(0.01*(k + 1));
    const C dx(0.31), dy(-0.22), dz(0.45);
    const int minImage = 24;        // the oracle's getPeriodicDelta + delta + r2, counted per pair
    // full site x bare charge: energy, force on both, torque on the full site (k_charge_site_pairs)
    {
        reset();
        C fB[3], tq[3];
        const C r2 = dx*dx + dy*dy + dz*dz;
        C phi = chargeSitePair<C, true>(m, C(0.001), C(0.002), C(-0.001), C(5.0), true, dx, dy, dz, r2, C(3.2853), C(8.0), fB, tq);
        // caller's part: kq = k qB, energy, force on both sites, torque on A
        C kq = C(138.9)*C(0.5), en = kq*phi, f0 = kq*fB[0], f1 = kq*fB[1], f2 = kq*fB[2], t0 = kq*tq[0], t1 = kq*tq[1], t2 = kq*tq[2];
        C acc(0); acc += en; acc += f0; acc += f1; acc += f2; acc -= f0; acc -= f1; acc -= f2; acc += t0; acc += t1; acc += t2;
        nAdd += minImage - 5;       // r2 is already counted above
        report("full x bare-charge pair (chargeSitePair + accumulation)");
    }
    // charge x charge (k_simple_pairs evaluates it from both sides; counted once, both forces)
    {
        reset();
        const C r2 = dx*dx + dy*dy + dz*dz;
        const C rinv = t_rsqrt(r2);
        const C qq = C(138.9)*C(0.5)*C(0.5);
        const C x = C(3.2853)*r2*rinv;
        const C ex = t_expneg(-(x*x));
        const C B1 = t_erfc_ex(x, ex);
        const C B2 = B1 + C(2.0/1.77245385091)*x*ex;
        C en(0); en += qq*B1*rinv;
        const C fr = -(qq*B2*rinv*rinv*rinv);
        C fx(0), fy(0), fz(0);
        fx += fr*dx; fy += fr*dy; fz += fr*dz;
        fx -= fr*dx; fy -= fr*dy; fz -= fr*dz;      // the partner's force (Newton's third law: 3 more adds, no multiplies)
        nMul -= 3;
        nAdd += minImage - 5;
        report("charge x charge pair");
    }
    // directed permanent field at a polarizable site (k_fixed_field: one direction of the 430-flop pair routine)
    {
        reset();
        const C r2 = dx*dx + dy*dy + dz*dz;
        C c[4], ex(0), ey(0), ez(0), sx(0), sy(0), sz(0);
        fieldCoefficientsOrdinary<C, true, 4>(r2, C(3.2853), C(8.0), C(5.0), c);
        fixedFieldDirected<C>(m, dx, dy, dz, c, ex, ey, ez);
        sx += ex; sy += ey; sz += ez;
        nAdd += minImage - 5;
        report("directed permanent field (coefficients + field)");
    }
    // directed induced field (k_induced_field: both directions of a pair = 2 x this)
    {
        reset();
        const C r2 = dx*dx + dy*dy + dz*dz;
        C c[4], ex(0), ey(0), ez(0);
        fieldCoefficientsOrdinary<C, true, 2>(r2, C(3.2853), C(8.0), C(5.0), c);
        inducedFieldDirected<C>(C(0.001), C(0.002), C(-0.001), dx, dy, dz, c, ex, ey, ez);
        nAdd += minImage - 5;
        report("directed induced-dipole field (coefficients + field)");
    }
    return 0;
}
