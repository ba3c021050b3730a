#ifndef OPENMM_COMPAT_VEC3_H_
#define OPENMM_COMPAT_VEC3_H_
// Minimal stand-in for OpenMM's Vec3 (public API subset used by the MPID plugin).
// Part of the MPIDB200 "openmm compat" layer: just enough of the OpenMM C++ API to
// build and run plugin code (the reference's and ours) in a container without OpenMM.
#include <cassert>
#include <cmath>
#include <iosfwd>
#include <ostream>

namespace OpenMM {

class Vec3 {
public:
    Vec3() : v{0.0, 0.0, 0.0} {}
    Vec3(double x, double y, double z) : v{x, y, z} {}
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    bool operator==(const Vec3& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
    bool operator!=(const Vec3& o) const { return !(*this == o); }
    Vec3 operator+() const { return *this; }
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0]+o.v[0], v[1]+o.v[1], v[2]+o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0]-o.v[0], v[1]-o.v[1], v[2]-o.v[2]); }
    Vec3& operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
    Vec3 operator*(double s) const { return Vec3(v[0]*s, v[1]*s, v[2]*s); }
    Vec3& operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    Vec3 operator/(double s) const { double r = 1.0/s; return Vec3(v[0]*r, v[1]*r, v[2]*r); }
    Vec3& operator/=(double s) { double r = 1.0/s; v[0] *= r; v[1] *= r; v[2] *= r; return *this; }
    double dot(const Vec3& o) const { return v[0]*o.v[0] + v[1]*o.v[1] + v[2]*o.v[2]; }
    Vec3 cross(const Vec3& o) const {
        return Vec3(v[1]*o.v[2]-v[2]*o.v[1], v[2]*o.v[0]-v[0]*o.v[2], v[0]*o.v[1]-v[1]*o.v[0]);
    }
private:
    double v[3];
};

static inline Vec3 operator*(double s, const Vec3& a) { return a*s; }

template <class CHAR, class TRAITS>
std::basic_ostream<CHAR, TRAITS>& operator<<(std::basic_ostream<CHAR, TRAITS>& o, const Vec3& v) {
    o << '[' << v[0] << ", " << v[1] << ", " << v[2] << ']';
    return o;
}

} // namespace OpenMM
#endif
