#ifndef OPENMM_COMPAT_FORCE_H_
#define OPENMM_COMPAT_FORCE_H_
// Minimal stand-in for OpenMM::Force (compat layer, see Vec3.h).
#include "openmm/internal/windowsExport.h"
namespace OpenMM {
class Context;
class ContextImpl;
class ForceImpl;
class OPENMM_EXPORT Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int group) { forceGroup = group; }
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
protected:
    friend class ContextImpl;
    virtual ForceImpl* createImpl() const = 0;
    ForceImpl& getImplInContext(Context& context);
    const ForceImpl& getImplInContext(const Context& context) const;
    ContextImpl& getContextImpl(Context& context);
private:
    int forceGroup;
};
} // namespace OpenMM
#endif
