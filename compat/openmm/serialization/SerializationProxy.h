// openmm-compat: proxy registry of OpenMM's serialization layer (subset; see SerializationNode.h).
#ifndef OPENMM_COMPAT_SERIALIZATIONPROXY_H_
#define OPENMM_COMPAT_SERIALIZATIONPROXY_H_

#include "openmm/serialization/SerializationNode.h"
#include <string>
#include <typeinfo>

namespace OpenMM {

class SerializationProxy {
public:
    explicit SerializationProxy(const std::string& typeName) : typeName(typeName) {}
    virtual ~SerializationProxy() {}
    const std::string& getTypeName() const { return typeName; }
    virtual void serialize(const void* object, SerializationNode& node) const = 0;
    virtual void* deserialize(const SerializationNode& node) const = 0;
    static void registerProxy(const std::type_info& type, const SerializationProxy* proxy);
    static const SerializationProxy& getProxy(const std::string& typeName);
    static const SerializationProxy& getProxy(const std::type_info& type);
private:
    std::string typeName;
};

} // namespace OpenMM
#endif
