// Stand-in for <yaml-cpp/yaml.h>: enough surface for PinholeStereoCamera's YAML constructor to compile; the oracle
// builds cameras through the explicit-parameter constructor, so loading throws.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
namespace YAML {
class Node {
public:
    Node operator[](const std::string &) const { return Node(); }
    bool IsDefined() const { return false; }
    template <typename T> T as() const { throw std::logic_error("[yaml-cpp stand-in] not available"); }
};
inline Node LoadFile(const std::string &) { throw std::logic_error("[yaml-cpp stand-in] not available"); }
} // namespace YAML
