#ifndef OPENMM_COMPAT_UNITS_H_
#define OPENMM_COMPAT_UNITS_H_
namespace OpenMM {
static const double NmPerAngstrom = 0.1;
static const double AngstromsPerNm = 10.0;
static const double KJPerKcal = 4.184;
static const double KcalPerKJ = 1.0/4.184;
static const double RadiansPerDegree = 3.1415926535897932385/180.0;
static const double DegreesPerRadian = 180.0/3.1415926535897932385;
} // namespace OpenMM
#endif
