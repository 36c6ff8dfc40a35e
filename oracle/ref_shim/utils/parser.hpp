// oracle/_ref only: takes the place of the reference's <utils/parser.hpp> on the include path.  The real header includes
// the whole simulation layer (boost, HighFive, NDArray: not available offline); the geometry sources compiled into
// oracle/_ref (src/cell.cpp, cell_universe.cpp, lattice.cpp, rect_lattice.cpp, hex_lattice.cpp ...) use it only inside
// their YAML factory functions, which are never called.  Declared here with the signatures of include/utils/parser.hpp:
// the id -> index maps and find_universe; oracle/ref_probe.cpp defines them.
#pragma once
#include <utils/settings.hpp>
#include <yaml-cpp/yaml.h>

#include <cstdint>
#include <map>
#include <memory>
#include <string>

extern std::map<uint32_t, size_t> surface_id_to_indx;
extern std::map<uint32_t, size_t> cell_id_to_indx;
extern std::map<uint32_t, size_t> universe_id_to_indx;
void make_universe(const YAML::Node& uni_node, const YAML::Node& input);
void find_universe(const YAML::Node& input, uint32_t id);
