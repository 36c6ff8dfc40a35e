// Stand-in for PapillonNDL (continuous-energy nuclear data library, fetched from the network by the reference's CMake),
// for oracle/_ref only.  Headers on the multigroup path (settings.hpp -> nd_directory.hpp -> ce_nuclide.hpp) name these
// types in declarations and one inline accessor; no continuous-energy translation unit is compiled into oracle/_ref and
// none of these members is ever called.
#pragma once
#include <vector>
namespace pndl {
struct ACE { enum class Type { ASCII, BINARY }; };
struct URRPTablesStandIn { const std::vector<double>& energy() const { static const std::vector<double> e; return e; } };
class STNeutron { public: const URRPTablesStandIn& urr_ptables() const { static const URRPTablesStandIn t; return t; } };
class STThermalScatteringLaw {};
}  // namespace pndl
