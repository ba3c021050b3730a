#ifndef OPENMM_COMPAT_PLATFORM_H_
#define OPENMM_COMPAT_PLATFORM_H_
#include "openmm/Kernel.h"
#include "openmm/KernelFactory.h"
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class Context;
class ContextImpl;
class OPENMM_EXPORT Platform {
public:
    virtual ~Platform();
    virtual const std::string& getName() const = 0;
    virtual double getSpeed() const { return 1.0; }
    virtual bool supportsDoublePrecision() const { return true; }
    const std::vector<std::string>& getPropertyNames() const { return propertyNames; }
    virtual const std::string& getPropertyValue(const Context& context, const std::string& property) const;
    const std::string& getPropertyDefaultValue(const std::string& property) const;
    void setPropertyDefaultValue(const std::string& property, const std::string& value);
    // Called by ContextImpl; platforms allocate / free their per-Context data here.
    virtual void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {}
    virtual void contextDestroyed(ContextImpl& context) const {}
    void registerKernelFactory(const std::string& name, KernelFactory* factory);
    bool supportsKernels(const std::vector<std::string>& kernelNames) const;
    Kernel createKernel(const std::string& name, ContextImpl& context) const;
    static void registerPlatform(Platform* platform);
    static int getNumPlatforms();
    static Platform& getPlatform(int index);
    static Platform& getPlatformByName(const std::string& name);
protected:
    Platform() {}
    void platformProperties(const std::vector<std::string>& names) { propertyNames = names; }
    std::vector<std::string> propertyNames;
    std::map<std::string, std::string> defaultProperties;
private:
    std::map<std::string, KernelFactory*> kernelFactories;
};
} // namespace OpenMM
#endif
