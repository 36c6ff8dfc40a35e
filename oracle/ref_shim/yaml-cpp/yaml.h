// Stand-in for yaml-cpp, for oracle/_ref only (see oracle/ref_probe.cpp).  The reference's surface sources include
// <yaml-cpp/yaml.h> for their make_*() factory functions; yaml-cpp itself is fetched from the network by the reference's
// CMake and is not available offline.  This header lets those translation units compile; the factories are never called
// (oracle/ref_probe.cpp constructs the surfaces through their ordinary constructors), and every accessor throws.
#pragma once
#include <cstddef>
#include <array>
#include <map>
#include <set>
#include <memory>
#include <sstream>
#include <vector>
#include <stdexcept>
#include <string>
namespace YAML {
class Node {
 public:
  Node operator[](const std::string&) const { return Node(); }
  Node operator[](const char*) const { return Node(); }
  Node operator[](std::size_t) const { return Node(); }
  Node operator[](int) const { return Node(); }
  template <class T>
  T as() const { throw std::runtime_error("yaml-cpp stand-in: no document"); }
  bool IsScalar() const { return false; }
  bool IsSequence() const { return false; }
  bool IsMap() const { return false; }
  bool IsDefined() const { return false; }
  bool IsNull() const { return true; }
  std::size_t size() const { return 0; }
  explicit operator bool() const { return false; }
  bool operator!() const { return true; }
};
}  // namespace YAML
