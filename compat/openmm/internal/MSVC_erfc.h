#ifndef OPENMM_COMPAT_MSVC_ERFC_H_
#define OPENMM_COMPAT_MSVC_ERFC_H_
#include <cmath>
#endif
