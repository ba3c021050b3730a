// openmm-compat: implementation of the serialization subset (SerializationNode, SerializationProxy, XmlSerializer).
// Written for this repository so that the reference's MPIDForceProxy and its own test compile and run unmodified.
#include "openmm/OpenMMException.h"
#include "openmm/serialization/SerializationNode.h"
#include "openmm/serialization/SerializationProxy.h"
#include "openmm/serialization/XmlSerializer.h"
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <istream>
#include <iterator>
#include <ostream>
#include <sstream>

namespace OpenMM {

// ---- SerializationNode ---------------------------------------------------------------------------------------------
const SerializationNode& SerializationNode::getChildNode(const std::string& childName) const {
    for (const SerializationNode& c : children)
        if (c.name == childName) return c;
    throw OpenMMException("Unknown child '" + childName + "' in node '" + name + "'");
}
SerializationNode& SerializationNode::getChildNode(const std::string& childName) {
    for (SerializationNode& c : children)
        if (c.name == childName) return c;
    throw OpenMMException("Unknown child '" + childName + "' in node '" + name + "'");
}
SerializationNode& SerializationNode::createChildNode(const std::string& childName) {
    children.push_back(SerializationNode());
    children.back().setName(childName);
    return children.back();
}
const std::string& SerializationNode::getStringProperty(const std::string& key) const {
    std::map<std::string, std::string>::const_iterator it = properties.find(key);
    if (it == properties.end()) throw OpenMMException("Unknown property '" + key + "' in node '" + name + "'");
    return it->second;
}
const std::string& SerializationNode::getStringProperty(const std::string& key, const std::string& defaultValue) const {
    std::map<std::string, std::string>::const_iterator it = properties.find(key);
    return it == properties.end() ? defaultValue : it->second;
}
SerializationNode& SerializationNode::setStringProperty(const std::string& key, const std::string& value) {
    properties[key] = value;
    return *this;
}
int SerializationNode::getIntProperty(const std::string& key) const { return (int) std::strtol(getStringProperty(key).c_str(), nullptr, 10); }
int SerializationNode::getIntProperty(const std::string& key, int d) const { return hasProperty(key) ? getIntProperty(key) : d; }
SerializationNode& SerializationNode::setIntProperty(const std::string& key, int value) { return setStringProperty(key, std::to_string(value)); }
long long SerializationNode::getLongProperty(const std::string& key) const { return std::strtoll(getStringProperty(key).c_str(), nullptr, 10); }
long long SerializationNode::getLongProperty(const std::string& key, long long d) const { return hasProperty(key) ? getLongProperty(key) : d; }
SerializationNode& SerializationNode::setLongProperty(const std::string& key, long long value) { return setStringProperty(key, std::to_string(value)); }
bool SerializationNode::getBoolProperty(const std::string& key) const { return getIntProperty(key) != 0; }
bool SerializationNode::getBoolProperty(const std::string& key, bool d) const { return hasProperty(key) ? getBoolProperty(key) : d; }
SerializationNode& SerializationNode::setBoolProperty(const std::string& key, bool value) { return setStringProperty(key, value ? "1" : "0"); }
double SerializationNode::getDoubleProperty(const std::string& key) const { return std::strtod(getStringProperty(key).c_str(), nullptr); }
double SerializationNode::getDoubleProperty(const std::string& key, double d) const { return hasProperty(key) ? getDoubleProperty(key) : d; }
SerializationNode& SerializationNode::setDoubleProperty(const std::string& key, double value) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", value);      // 17 significant digits round-trip every double
    return setStringProperty(key, buf);
}

// ---- SerializationProxy registry -------------------------------------------------------------------------------------
namespace {
struct Registry {
    std::map<std::string, const SerializationProxy*> byTypeName, byTypeId;
};
Registry& registry() { static Registry r; return r; }
}
void SerializationProxy::registerProxy(const std::type_info& type, const SerializationProxy* proxy) {
    registry().byTypeId[type.name()] = proxy;
    registry().byTypeName[proxy->getTypeName()] = proxy;
}
const SerializationProxy& SerializationProxy::getProxy(const std::string& typeName) {
    std::map<std::string, const SerializationProxy*>::const_iterator it = registry().byTypeName.find(typeName);
    if (it == registry().byTypeName.end()) throw OpenMMException("There is no serialization proxy registered for type " + typeName);
    return *it->second;
}
const SerializationProxy& SerializationProxy::getProxy(const std::type_info& type) {
    std::map<std::string, const SerializationProxy*>::const_iterator it = registry().byTypeId.find(type.name());
    if (it == registry().byTypeId.end()) throw OpenMMException(std::string("There is no serialization proxy registered for type ") + type.name());
    return *it->second;
}

// ---- XmlSerializer ---------------------------------------------------------------------------------------------------
namespace {
std::string escape(const std::string& s) {
    std::string out;
    for (char c : s) {
        switch (c) {
            case '&': out += "&amp;"; break;
            case '<': out += "&lt;"; break;
            case '>': out += "&gt;"; break;
            case '"': out += "&quot;"; break;
            default: out += c;
        }
    }
    return out;
}
std::string unescape(const std::string& s) {
    std::string out;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] != '&') { out += s[i]; continue; }
        if (s.compare(i, 5, "&amp;") == 0) { out += '&'; i += 4; }
        else if (s.compare(i, 4, "&lt;") == 0) { out += '<'; i += 3; }
        else if (s.compare(i, 4, "&gt;") == 0) { out += '>'; i += 3; }
        else if (s.compare(i, 6, "&quot;") == 0) { out += '"'; i += 5; }
        else out += s[i];
    }
    return out;
}
void writeNode(const SerializationNode& node, std::ostream& os, int depth) {
    os << std::string(depth, '\t') << '<' << node.getName();
    for (const auto& kv : node.getProperties()) os << ' ' << kv.first << "=\"" << escape(kv.second) << '"';
    if (node.getChildren().empty()) { os << "/>\n"; return; }
    os << ">\n";
    for (const SerializationNode& c : node.getChildren()) writeNode(c, os, depth + 1);
    os << std::string(depth, '\t') << "</" << node.getName() << ">\n";
}
struct Parser {
    const std::string& t;
    size_t p = 0;
    explicit Parser(const std::string& text) : t(text) {}
    void fail(const std::string& why) const { throw OpenMMException("XmlSerializer: " + why + " at offset " + std::to_string(p)); }
    void skipSpace() { while (p < t.size() && std::isspace((unsigned char) t[p])) p++; }
    void skipProlog() {
        for (;;) {
            skipSpace();
            if (t.compare(p, 2, "<?") == 0) { size_t e = t.find("?>", p); if (e == std::string::npos) fail("unterminated declaration"); p = e + 2; }
            else if (t.compare(p, 4, "<!--") == 0) { size_t e = t.find("-->", p); if (e == std::string::npos) fail("unterminated comment"); p = e + 3; }
            else return;
        }
    }
    std::string name() {
        size_t b = p;
        while (p < t.size() && (std::isalnum((unsigned char) t[p]) || t[p] == '_' || t[p] == ':' || t[p] == '-' || t[p] == '.')) p++;
        if (p == b) fail("expected a name");
        return t.substr(b, p - b);
    }
    void element(SerializationNode& node) {
        skipProlog();
        if (p >= t.size() || t[p] != '<') fail("expected '<'");
        p++;
        node.setName(name());
        for (;;) {
            skipSpace();
            if (p >= t.size()) fail("unterminated element");
            if (t[p] == '/') { if (t.compare(p, 2, "/>") != 0) fail("expected '/>'"); p += 2; return; }
            if (t[p] == '>') { p++; break; }
            std::string key = name();
            skipSpace();
            if (p >= t.size() || t[p] != '=') fail("expected '='");
            p++;
            skipSpace();
            if (p >= t.size() || (t[p] != '"' && t[p] != '\'')) fail("expected a quoted value");
            const char q = t[p++];
            size_t e = t.find(q, p);
            if (e == std::string::npos) fail("unterminated attribute value");
            node.setStringProperty(key, unescape(t.substr(p, e - p)));
            p = e + 1;
        }
        for (;;) {
            skipProlog();
            if (p >= t.size()) fail("missing closing tag of " + node.getName());
            if (t.compare(p, 2, "</") == 0) {
                p += 2;
                if (name() != node.getName()) fail("mismatched closing tag");
                skipSpace();
                if (p >= t.size() || t[p] != '>') fail("expected '>'");
                p++;
                return;
            }
            if (t[p] != '<') { p++; continue; }          // text content is not part of the format: skipped
            element(node.createChildNode(""));
        }
    }
};
}
void XmlSerializer::write(const SerializationNode& node, std::ostream& stream) {
    stream << "<?xml version=\"1.0\" ?>\n";
    writeNode(node, stream, 0);
}
void XmlSerializer::read(std::istream& stream, SerializationNode& node) {
    const std::string text((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
    Parser(text).element(node);
}
void* XmlSerializer::deserializeStream(std::istream& stream) {
    SerializationNode node;
    read(stream, node);
    return SerializationProxy::getProxy(node.getStringProperty("type")).deserialize(node);
}
void XmlSerializer::throw_reserved() { throw OpenMMException("XmlSerializer: the property name 'type' is reserved"); }

} // namespace OpenMM
