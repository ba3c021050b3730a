// openmm-compat: the subset of OpenMM's serialization tree that the MPID serialization proxy and its test use
// (OpenMM itself is not installable here).  A node has a name, string-valued properties and child nodes;
// numbers are stored as text that round-trips a double exactly.  Written for this repository; not OpenMM source.
#ifndef OPENMM_COMPAT_SERIALIZATIONNODE_H_
#define OPENMM_COMPAT_SERIALIZATIONNODE_H_

#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class SerializationNode {
public:
    const std::string& getName() const { return name; }
    void setName(const std::string& n) { name = n; }
    const std::vector<SerializationNode>& getChildren() const { return children; }
    std::vector<SerializationNode>& getChildren() { return children; }
    const SerializationNode& getChildNode(const std::string& childName) const;
    SerializationNode& getChildNode(const std::string& childName);
    SerializationNode& createChildNode(const std::string& childName);
    const std::map<std::string, std::string>& getProperties() const { return properties; }
    bool hasProperty(const std::string& key) const { return properties.find(key) != properties.end(); }

    const std::string& getStringProperty(const std::string& key) const;
    const std::string& getStringProperty(const std::string& key, const std::string& defaultValue) const;
    SerializationNode& setStringProperty(const std::string& key, const std::string& value);
    int getIntProperty(const std::string& key) const;
    int getIntProperty(const std::string& key, int defaultValue) const;
    SerializationNode& setIntProperty(const std::string& key, int value);
    long long getLongProperty(const std::string& key) const;
    long long getLongProperty(const std::string& key, long long defaultValue) const;
    SerializationNode& setLongProperty(const std::string& key, long long value);
    bool getBoolProperty(const std::string& key) const;
    bool getBoolProperty(const std::string& key, bool defaultValue) const;
    SerializationNode& setBoolProperty(const std::string& key, bool value);
    double getDoubleProperty(const std::string& key) const;
    double getDoubleProperty(const std::string& key, double defaultValue) const;
    SerializationNode& setDoubleProperty(const std::string& key, double value);

private:
    std::string name;
    std::vector<SerializationNode> children;
    std::map<std::string, std::string> properties;
};

} // namespace OpenMM
#endif
