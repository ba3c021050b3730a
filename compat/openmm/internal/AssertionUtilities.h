#ifndef OPENMM_COMPAT_ASSERTIONUTILITIES_H_
#define OPENMM_COMPAT_ASSERTIONUTILITIES_H_
#include "openmm/OpenMMException.h"
#include <cmath>
#include <sstream>
#include <string>
namespace OpenMM {
static inline void throwException(const char* file, int line, const std::string& details) {
    std::stringstream m;
    m << "Assertion failure at " << file << ":" << line;
    if (details.size() > 0) m << ".  " << details;
    throw OpenMMException(m.str());
}
} // namespace OpenMM
#define ASSERT(cond) {if (!(cond)) OpenMM::throwException(__FILE__, __LINE__, "");};
#define ASSERT_EQUAL(expected, found) {if (!((expected) == (found))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_TOL(expected, found, tol) {double _scale_ = std::abs(expected) > 1.0 ? std::abs(expected) : 1.0; if (!(std::abs((expected)-(found))/_scale_ <= (tol))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_VEC(expected, found, tol) {double _norm_ = std::sqrt((expected).dot(expected)); double _scale_ = _norm_ > 1.0 ? _norm_ : 1.0; if ((std::abs(((expected)[0])-((found)[0]))/_scale_ > (tol)) || (std::abs(((expected)[1])-((found)[1]))/_scale_ > (tol)) || (std::abs(((expected)[2])-((found)[2]))/_scale_ > (tol))) {std::stringstream details; details << " Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_USUALLY_TRUE(cond) ASSERT(cond)
#define ASSERT_USUALLY_EQUAL_TOL(expected, found, tol) ASSERT_EQUAL_TOL(expected, found, tol)
#define ASSERT_EQUAL_CONTAINERS(expected, found) {if ((expected).size() != (found).size()) OpenMM::throwException(__FILE__, __LINE__, "container size mismatch"); auto _a_ = (expected).begin(); auto _b_ = (found).begin(); for (; _a_ != (expected).end(); ++_a_, ++_b_) if (!(*_a_ == *_b_)) OpenMM::throwException(__FILE__, __LINE__, "container element mismatch");};
#endif
