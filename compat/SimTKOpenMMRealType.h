#ifndef COMPAT_SIMTKREAL_H_
#define COMPAT_SIMTKREAL_H_
#include <cmath>
#endif
