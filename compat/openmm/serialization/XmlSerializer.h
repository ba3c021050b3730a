// openmm-compat: XML front end of OpenMM's serialization layer (subset; see SerializationNode.h).  The root element
// carries the proxy's type name in a "type" attribute, properties become attributes, child nodes child elements.
#ifndef OPENMM_COMPAT_XMLSERIALIZER_H_
#define OPENMM_COMPAT_XMLSERIALIZER_H_

#include "openmm/serialization/SerializationNode.h"
#include "openmm/serialization/SerializationProxy.h"
#include <iosfwd>
#include <string>
#include <typeinfo>

namespace OpenMM {

class XmlSerializer {
public:
    template <class T>
    static void serialize(const T* object, const std::string& rootName, std::ostream& stream) {
        const SerializationProxy& proxy = SerializationProxy::getProxy(typeid(*object));
        SerializationNode node;
        node.setName(rootName);
        proxy.serialize(object, node);
        if (node.hasProperty("type"))
            throw_reserved();
        node.setStringProperty("type", proxy.getTypeName());
        write(node, stream);
    }
    template <class T>
    static T* deserialize(std::istream& stream) {
        return reinterpret_cast<T*>(deserializeStream(stream));
    }
    // node <-> text, exposed for tools that want the tree itself
    static void write(const SerializationNode& node, std::ostream& stream);
    static void read(std::istream& stream, SerializationNode& node);
private:
    static void* deserializeStream(std::istream& stream);
    static void throw_reserved();
};

} // namespace OpenMM
#endif
